// Microtest (development aid): how exact is the fp32 accumulation of tcgen05.mma.kind::f16 over a long chain?
//   D = Z^T Z, Z = K x 64 fp16 (MN-major, 128B swizzle as in cholesky_tc.cu), M = N = 64, one MMA per 16 rows of Z,
// against the exact sum (fp64 of the same fp16 values), for K = 32 ... 4096 and for data with a non-zero mean (all
// products of one sign on the diagonal, where a truncating adder shows a bias linear in the chain length).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/umma_chain_test tools/umma_chain_test.cu
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one CTA; the tile buffer holds 64 rows (8 KB); the chain is processed 64 rows at a time (commit + wait in between)
__global__ void __launch_bounds__(128) chain_kernel(const __half *z, float *out, int k_rows) {
  extern __shared__ unsigned char raw[];
  const uint32_t r0 = smem_u32(raw);
  const uint32_t base = (r0 + 1023u) & ~1023u;
  unsigned char *g = raw + (base - r0);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
  uint32_t phase = 0;
  for (int k0 = 0; k0 < k_rows; k0 += 64) {
    for (int e = threadIdx.x; e < 64 * 8; e += 128) {
      const int r = e >> 3, c = e & 7;
      *reinterpret_cast<uint4 *>(g + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(z + (size_t)(k0 + r) * 64 + c * 8);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t d = (uint64_t)(((base + ks * 2048) >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t acc = (k0 > 0 || ks > 0);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(d), "l"(d), "r"(idesc), "r"(acc)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
            smem_u32(&bar)),
        "r"(phase)
        : "memory");
    phase ^= 1;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 64; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c)
                 : "memory");
    if (lane < 16)
      for (int j = 0; j < 8; ++j) out[(16 * warp + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

static double gauss() {
  double u = (rand() + 1.0) / (RAND_MAX + 2.0), v = (rand() + 1.0) / (RAND_MAX + 2.0);
  return sqrt(-2 * log(u)) * cos(6.283185307179586 * v);
}

int main() {
  const int KMAX = 4096;
  cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  __half *dz;
  float *dout;
  cudaMalloc(&dz, KMAX * 64 * 2);
  cudaMalloc(&dout, 64 * 64 * 4);
  for (int variant = 0; variant < 2; ++variant) {
    // variant 0: zero-mean entries (like centred factors); 1: mean 1 sigma (all-positive-ish: the first ALS iterations)
    srand(11);
    std::vector<__half> z(KMAX * 64);
    std::vector<double> f(KMAX * 64);
    for (int i = 0; i < KMAX * 64; ++i) {
      z[i] = __float2half((float)(800.0 * (gauss() + (variant ? 1.0 : 0.0))));
      f[i] = (double)__half2float(z[i]);
    }
    cudaMemcpy(dz, z.data(), KMAX * 64 * 2, cudaMemcpyHostToDevice);
    for (int K = 64; K <= KMAX; K *= 4) {
      chain_kernel<<<1, 128, 16384>>>(dz, dout, K);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("CUDA error %s\n", cudaGetErrorString(e));
        return 1;
      }
      std::vector<float> out(64 * 64);
      cudaMemcpy(out.data(), dout, 64 * 64 * 4, cudaMemcpyDeviceToHost);
      std::vector<double> ref(64 * 64, 0.0);
      std::vector<float> seq(64 * 64, 0.f);  // fp32, one addition per product, in order (what a scalar fp32 loop gives)
      for (int k = 0; k < K; ++k)
        for (int m = 0; m < 64; ++m)
          for (int n = 0; n < 64; ++n) {
            ref[m * 64 + n] += f[k * 64 + m] * f[k * 64 + n];
            seq[m * 64 + n] += (float)(f[k * 64 + m] * f[k * 64 + n]);
          }
      double dmax = 0, dmean = 0, smax = 0, diag_bias = 0, sdiag_bias = 0;
      for (int m = 0; m < 64; ++m)
        for (int n = 0; n < 64; ++n) {
          const double sc = sqrt(ref[m * 64 + m] * ref[n * 64 + n]);
          const double e1 = (out[m * 64 + n] - ref[m * 64 + n]) / sc, e2 = (seq[m * 64 + n] - ref[m * 64 + n]) / sc;
          dmax = fmax(dmax, fabs(e1));
          dmean += fabs(e1) / 4096;
          smax = fmax(smax, fabs(e2));
          if (m == n) {
            diag_bias += e1 / 64;
            sdiag_bias += e2 / 64;
          }
        }
      printf("%s K = %4d (%3d MMAs): tcgen05 error / sqrt(d_ii d_jj): max %.2e mean %.2e, mean signed error on the diagonal %+.2e | "
             "sequential fp32: max %.2e, diagonal %+.2e\n",
             variant ? "mean-1-sigma" : "zero-mean   ", K, K / 16, dmax, dmean, diag_bias, smax, sdiag_bias);
    }
  }
  return 0;
}
