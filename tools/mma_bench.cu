// Microbenchmark: issue rate of legacy warp-level mma.sync on sm_100a (TF32 m16n8k8, BF16 m16n8k16)
// and of plain FFMA, per SM.  Used to size the 3xTF32 accumulation in cholesky.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench tools/mma_bench.cu && tools/mma_bench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int ILP>
__global__ void tf32_kernel(float *out, int iters) {
  float d[ILP][4];
  for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void bf16_kernel(float *out, int iters) {
  float d[ILP][4];
  for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void ffma_kernel(float *out, int iters) {
  float d[ILP];
  for (int i = 0; i < ILP; ++i) d[i] = threadIdx.x * 1e-3f + i;
  float a = 1.0001f + threadIdx.x * 1e-7f, b = 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) d[i] = fmaf(d[i], a, b);
  }
  float s = 0;
  for (int i = 0; i < ILP; ++i) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename K>
double time_kernel(K launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("%s: %d SMs, max clock %.0f MHz\n", p.name, sms, clk_khz / 1e3);
  float *out;
  cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    const int threads = warps * 32;
    {
      double ms = time_kernel([&] { tf32_kernel<8><<<sms, threads>>>(out, iters); });
      double mmas = (double)sms * warps * iters * 8;
      printf("tf32 m16n8k8  warps/SM=%2d: %.3f ms  -> %.1f MMA/us/SM, %.0f MAC/clk/SM @max clock, %.1f TF32-TFLOP/s\n", warps, ms,
             mmas / sms / (ms * 1e3), mmas * 1024 / sms / (ms * 1e-3) / (clk_khz * 1e3), mmas * 2048 / (ms * 1e-3) / 1e12);
    }
    {
      double ms = time_kernel([&] { bf16_kernel<8><<<sms, threads>>>(out, iters); });
      double mmas = (double)sms * warps * iters * 8;
      printf("bf16 m16n8k16 warps/SM=%2d: %.3f ms  -> %.1f MMA/us/SM, %.0f MAC/clk/SM @max clock, %.1f BF16-TFLOP/s\n", warps, ms,
             mmas / sms / (ms * 1e3), mmas * 2048 / sms / (ms * 1e-3) / (clk_khz * 1e3), mmas * 4096 / (ms * 1e-3) / 1e12);
    }
    {
      double ms = time_kernel([&] { ffma_kernel<8><<<sms, threads>>>(out, iters); });
      double f = (double)sms * threads * iters * 8;
      printf("ffma          warps/SM=%2d: %.3f ms  -> %.0f FMA/clk/SM @max clock, %.1f TFLOP/s\n", warps, ms,
             f / sms / (ms * 1e-3) / (clk_khz * 1e3), f * 2 / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}
