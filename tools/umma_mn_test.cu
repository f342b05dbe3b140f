// Microtest (development aid, not part of the library): tcgen05.mma kind::f16 with MN-major, 128B-swizzled operand
// tiles whose rows are what a gather produces -- K rows (nonzeros) of 64 halves (128 bytes) each.
//   D[m][n] = sum_k Za[k][m] Zb[k][n],  M = N = 64, K = 32 (two MMAs), fp32 accumulate in TMEM.
// Prints the max error against the host for a few (LBO, SBO) encodings, and the TMEM lane map of an M = 64 tile.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_bin/umma_mn_test tools/umma_mn_test.cu
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// second test: N = 128, B = [Zb | Zc]: two 64-element atoms along N, the tiles `lbo` apart; out has 128 columns
__global__ void __launch_bounds__(128) test_kernel_n128(const __half *za, const __half *zb, const __half *zc, float *out, uint32_t lbo,
                                                         uint32_t sbo, int k_rows) {
  extern __shared__ unsigned char raw[];
  const uint32_t r0 = smem_u32(raw);
  const uint32_t base = (r0 + 1023u) & ~1023u;
  unsigned char *g = raw + (base - r0);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < k_rows * 8; e += 128) {
    const int r = e >> 3, c = e & 7;
    *reinterpret_cast<uint4 *>(g + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(za + r * 64 + c * 8);
    *reinterpret_cast<uint4 *>(g + 8192 + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(zb + r * 64 + c * 8);
    *reinterpret_cast<uint4 *>(g + 12288 + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4 *>(zc + r * 64 + c * 8);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((64u >> 4) << 24);
    for (int ks = 0; ks < k_rows / 16; ++ks) {
      const uint64_t ad = (uint64_t)(((base + ks * 2048) >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
      const uint64_t bd = (uint64_t)(((base + 8192 + ks * 2048) >> 4) & 0x3fffu) | ((uint64_t)lbo << 16) | ((uint64_t)sbo << 32) |
                          (1ull << 46) | (2ull << 61);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
          smem_u32(&bar)),
      "r"(0)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 128; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c)
                 : "memory");
    for (int j = 0; j < 8; ++j) out[(32 * warp + lane) * 128 + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

__global__ void __launch_bounds__(128) test_kernel(const __half *za, const __half *zb, float *out, uint32_t lbo, uint32_t sbo,
                                                    int k_rows) {
  extern __shared__ unsigned char raw[];
  const uint32_t r0 = smem_u32(raw);
  const uint32_t base = (r0 + 1023u) & ~1023u;
  unsigned char *g = raw + (base - r0);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tiles: A at 0, B at 8192; row r (nonzero) at r * 128, 16-byte chunk c at position c ^ (r & 7)
  for (int e = threadIdx.x; e < k_rows * 8; e += 128) {
    const int r = e >> 3, c = e & 7;
    const uint4 va = *reinterpret_cast<const uint4 *>(za + r * 64 + c * 8);
    const uint4 vb = *reinterpret_cast<const uint4 *>(zb + r * 64 + c * 8);
    *reinterpret_cast<uint4 *>(g + r * 128 + ((c ^ (r & 7)) << 4)) = va;
    *reinterpret_cast<uint4 *>(g + 8192 + r * 128 + ((c ^ (r & 7)) << 4)) = vb;
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    // idesc: fp32 accumulate, A/B fp16, both MN-major, N = 64, M = 64
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
    for (int ks = 0; ks < k_rows / 16; ++ks) {
      const uint64_t ad = (uint64_t)(((base + ks * 2048) >> 4) & 0x3fffu) | ((uint64_t)lbo << 16) | ((uint64_t)sbo << 32) |
                          (1ull << 46) | (2ull << 61);
      const uint64_t bd = (uint64_t)(((base + 8192 + ks * 2048) >> 4) & 0x3fffu) | ((uint64_t)lbo << 16) |
                          ((uint64_t)sbo << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
          smem_u32(&bar)),
      "r"(0)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // every lane of every quarter dumps its 64 columns: out[(32 warp + lane)][64]
  for (int c = 0; c < 64; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)c)
                 : "memory");
    for (int j = 0; j < 8; ++j) out[(32 * warp + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

int main() {
  const int K = 32;
  std::vector<__half> za(K * 64), zb(K * 64);
  std::vector<float> fa(K * 64), fb(K * 64);
  srand(3);
  for (int i = 0; i < K * 64; ++i) {
    za[i] = __float2half((rand() % 2001 - 1000) / 1000.f);
    zb[i] = __float2half((rand() % 2001 - 1000) / 1000.f);
    fa[i] = __half2float(za[i]);
    fb[i] = __half2float(zb[i]);
  }
  std::vector<double> ref(64 * 64, 0.0);
  for (int k = 0; k < K; ++k)
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < 64; ++n) ref[m * 64 + n] += (double)fa[k * 64 + m] * fb[k * 64 + n];
  __half *da, *db;
  float *dout;
  cudaMalloc(&da, K * 64 * 2);
  cudaMalloc(&db, K * 64 * 2);
  cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, za.data(), K * 64 * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, zb.data(), K * 64 * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
  const uint32_t cand[][2] = {{1, 64}, {64, 64}, {64, 1}, {128, 64}, {64, 128}, {0, 64}, {8, 64}};
  for (auto &c : cand) {
    cudaMemset(dout, 0, 128 * 64 * 4);
    test_kernel<<<1, 128, 20480>>>(da, db, dout, c[0], c[1], K);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("lbo %u sbo %u: CUDA error %s\n", c[0], c[1], cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> out(128 * 64);
    cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    // lane map hypothesis A: row m at TMEM lane 32 (m / 16) + m % 16;  hypothesis B: lane m (first 64 lanes)
    double ea = 0, eb = 0, eat = 0;
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < 64; ++n) {
        const int la = 32 * (m / 16) + m % 16;
        ea = fmax(ea, fabs(out[la * 64 + n] - ref[m * 64 + n]));
        eat = fmax(eat, fabs(out[la * 64 + n] - ref[n * 64 + m]));
        eb = fmax(eb, fabs(out[m * 64 + n] - ref[m * 64 + n]));
      }
    printf("lbo %3u sbo %3u: max err (lanes 0-15 of each quarter) %.3e  (transposed %.3e)  (lanes 0-63) %.3e   ref[0][1] %.4f got %.4f\n",
           c[0], c[1], ea, eat, eb, ref[1], out[1]);
  }
  // ---- N = 128: B = [Zb | Zc] ----
  std::vector<__half> zc(K * 64);
  std::vector<float> fc(K * 64);
  for (int i = 0; i < K * 64; ++i) {
    zc[i] = __float2half((rand() % 2001 - 1000) / 1000.f);
    fc[i] = __half2float(zc[i]);
  }
  std::vector<double> ref2(64 * 64, 0.0);
  for (int k = 0; k < K; ++k)
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < 64; ++n) ref2[m * 64 + n] += (double)fa[k * 64 + m] * fc[k * 64 + n];
  __half *dc;
  float *dout2;
  cudaMalloc(&dc, K * 64 * 2);
  cudaMalloc(&dout2, 128 * 128 * 4);
  cudaMemcpy(dc, zc.data(), K * 64 * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(test_kernel_n128, cudaFuncAttributeMaxDynamicSharedMemorySize, 24576);
  const uint32_t cand2[][2] = {{256, 64}, {64, 256}, {1, 64}, {512, 64}};
  for (auto &c : cand2) {
    cudaMemset(dout2, 0, 128 * 128 * 4);
    test_kernel_n128<<<1, 128, 24576>>>(da, db, dc, dout2, c[0], c[1], K);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("N=128 lbo %u sbo %u: CUDA error %s\n", c[0], c[1], cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> out(128 * 128);
    cudaMemcpy(out.data(), dout2, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    double e1 = 0, e2 = 0;
    for (int m = 0; m < 64; ++m)
      for (int n = 0; n < 64; ++n) {
        const int la = 32 * (m / 16) + m % 16;
        e1 = fmax(e1, fabs(out[la * 128 + n] - ref[m * 64 + n]));
        e2 = fmax(e2, fabs(out[la * 128 + 64 + n] - ref2[m * 64 + n]));
      }
    printf("N=128 lbo %3u sbo %3u: max err columns 0-63 (Za^T Zb) %.3e, columns 64-127 (Za^T Zc) %.3e\n", c[0], c[1], e1, e2);
  }
  return 0;
}
