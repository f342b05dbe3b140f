// R4: Gramian  G = Y^T Y  (reference: np.dot(Y.T, Y), implicit/cpu/_als.pyx:70,164,268;
// LeastSquaresSolver::calculate_yty, implicit/gpu/als.cu:122-152).
//
// Bandwidth-bound reduction over the rows of Y (f/2 flop per byte): every CTA streams a slice of
// rows through shared memory, keeps a (F/16 x F/16) register tile of the F x F result per thread,
// and writes one partial; a second tiny kernel sums the partials in a fixed order in fp64 so the
// result is deterministic and independent of the grid size rounding.
#include "common.h"

namespace als {

constexpr int kGramRows = 32;  // rows of Y staged per step

template <int T>  // F = 16 * T
__global__ void __launch_bounds__(256) gramian_partial_kernel(const float *__restrict__ Y, int64_t rows, int ld,
                                                              float *__restrict__ partials) {
  constexpr int F = 16 * T;
  __shared__ __align__(16) float tile[kGramRows][F];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[T][T];
#pragma unroll
  for (int a = 0; a < T; ++a)
#pragma unroll
    for (int b = 0; b < T; ++b) acc[a][b] = 0.f;

  const int64_t steps = (rows + kGramRows - 1) / kGramRows;
  for (int64_t s = blockIdx.x; s < steps; s += gridDim.x) {
    const int64_t r0 = s * kGramRows;
    // coalesced float4 staging (ld == F by construction)
    for (int e = threadIdx.x; e < kGramRows * F / 4; e += 256) {
      const int r = e / (F / 4), c4 = e % (F / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < rows) v = __ldg(reinterpret_cast<const float4 *>(Y + (r0 + r) * ld) + c4);
      reinterpret_cast<float4 *>(&tile[r][0])[c4] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kGramRows; ++r) {
      float a[T], b[T];
#pragma unroll
      for (int i = 0; i < T; ++i) {
        a[i] = tile[r][ty + 16 * i];  // row index of G: ty + 16 i  (broadcast within a half-warp)
        b[i] = tile[r][tx + 16 * i];  // col index of G: tx + 16 i  (conflict free)
      }
#pragma unroll
      for (int i = 0; i < T; ++i)
#pragma unroll
        for (int j = 0; j < T; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  float *out = partials + (size_t)blockIdx.x * F * F;
#pragma unroll
  for (int i = 0; i < T; ++i)
#pragma unroll
    for (int j = 0; j < T; ++j) out[(ty + 16 * i) * F + tx + 16 * j] = acc[i][j];
}

__global__ void gramian_reduce_kernel(const float *__restrict__ partials, int nparts, int n, float *__restrict__ G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double s = 0.0;
  for (int p = 0; p < nparts; ++p) s += (double)partials[(size_t)p * n + e];
  G[e] = (float)s;
}

// Greg = G + lambda I on the real dimensions, identity on the zero-padded ones (so that padded
// unknowns solve to exactly 0 even with lambda == 0).  Mirrors `YtY + regularization * np.eye(f)`
// (implicit/cpu/_als.pyx:85, :164): an fp32 add of fp32(lambda).
__global__ void regularize_kernel(const float *__restrict__ G, float *__restrict__ Greg, int f, int ld, float lambda) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ld * ld) return;
  const int i = e / ld, j = e % ld;
  float v = G[e];
  if (i == j) v = (i < f) ? v + lambda : 1.0f;
  Greg[e] = v;
}

template <int T>
static int run_gramian(als_ctx *ctx, const als_factors *Y, int grid) {
  gramian_partial_kernel<T><<<grid, 256, 0, ctx->stream>>>(Y->d, Y->rows, Y->ld, ctx->gram_partials);
  return ALS_OK;
}

int launch_gramian(als_ctx *ctx, const als_factors *Y) {
  const int F = Y->ld;
  if (F > 128) {
    set_error("gramian: factors=%d (padded %d) > 128 is not supported yet", Y->f, F);
    return ALS_E_UNSUPPORTED;
  }
  const int64_t steps = ceil_div(std::max<int64_t>(Y->rows, 1), kGramRows);
  const int grid = (int)std::min<int64_t>(steps, (int64_t)ctx->sm_count * 2);
  const int64_t need = (int64_t)grid * F * F;
  if (need > ctx->gram_partials_cap) {
    if (ctx->gram_partials) {
      ALS_CUDA(cudaStreamSynchronize(ctx->stream));
      ALS_CUDA(cudaFree(ctx->gram_partials));
      ctx->gram_partials = nullptr;
    }
    const int64_t cap = (int64_t)ctx->sm_count * 2 * 128 * 128;
    ALS_CUDA(cudaMalloc(&ctx->gram_partials, sizeof(float) * cap));
    ctx->gram_partials_cap = cap;
  }
  ProfScope prof(ctx, kProfGramian);
  switch (F / 16) {
    case 1: run_gramian<1>(ctx, Y, grid); break;
    case 2: run_gramian<2>(ctx, Y, grid); break;
    case 3: run_gramian<3>(ctx, Y, grid); break;
    case 4: run_gramian<4>(ctx, Y, grid); break;
    case 5: run_gramian<5>(ctx, Y, grid); break;
    case 6: run_gramian<6>(ctx, Y, grid); break;
    case 7: run_gramian<7>(ctx, Y, grid); break;
    case 8: run_gramian<8>(ctx, Y, grid); break;
    default: set_error("gramian: bad padded factors %d", F); return ALS_E_UNSUPPORTED;
  }
  ALS_CUDA(cudaGetLastError());
  gramian_reduce_kernel<<<(F * F + 255) / 256, 256, 0, ctx->stream>>>(ctx->gram_partials, grid, F * F, ctx->G);
  ALS_CUDA(cudaGetLastError());
  ctx->launches += 2;
  return ALS_OK;
}

int launch_regularize(als_ctx *ctx, int f, int ld, float lambda) {
  regularize_kernel<<<(ld * ld + 255) / 256, 256, 0, ctx->stream>>>(ctx->G, ctx->Greg, f, ld, lambda);
  ALS_CUDA(cudaGetLastError());
  ctx->launches += 1;
  return ALS_OK;
}

}  // namespace als
