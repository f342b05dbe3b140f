// R6: training loss (reference: _calculate_loss, implicit/cpu/_als.pyx:259-308; the reference's GPU
// version is calculate_loss_kernel, implicit/gpu/als.cu:199-251).
#include "common.h"

namespace als {

namespace {

template <int F>
struct CgCfg {
  static constexpr int V = F / 4;  // float4 slices per factor row
  static constexpr int L = V <= 4 ? 4 : V <= 8 ? 8 : V <= 16 ? 16 : 32;  // lanes per group
  static constexpr int NG = 32 / L;                                     // groups per warp
};

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int m = 1; m < L; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// ---- loss --------------------------------------------------------------------------------------
// loss numerator = sum_u x_u^T (Y^T Y) x_u + sum_k [(-2 c_k^+ + (|c_k| - 1) d_k) d_k + |c_k|],  d_k = y_k . x_u
// (expanding r.x in _als.pyx:282-300); the quadratic term is <Y^T Y, X^T X>_F and the norms are the
// traces of the two Gramians, so only the per-nonzero term needs the CSR.
template <int F>
__global__ void __launch_bounds__(256)
loss_nnz_kernel(const int32_t *__restrict__ indices, const float *__restrict__ data, const float *__restrict__ Y,
                const float *__restrict__ X, int64_t row_offset, const WorkItem *__restrict__ work, int n_work,
                int32_t *counter, double *out /* [0]=term sum, [1]=sum |c| */) {
  using C = CgCfg<F>;
  const int lane = threadIdx.x & 31, sub = lane % C::L, grp = lane / C::L;
  const bool active = sub < C::V;
  double term = 0.0, conf_sum = 0.0;
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(counter, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n_work) break;
    const int4 w = __ldg(reinterpret_cast<const int4 *>(work) + i);
    if (w.w == -2) continue;
    const float4 x = active ? __ldg(reinterpret_cast<const float4 *>(X + (row_offset + w.x) * F) + sub)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kb = w.y; kb < w.z; kb += C::NG) {  // uniform trip count: full-warp shuffles inside
      const int k = kb + grp;
      const bool valid = k < w.z;
      const int idx = valid ? __ldg(indices + k) : 0;
      const float c = valid ? __ldg(data + k) : 0.f;
      const float4 y = (active && valid) ? __ldg(reinterpret_cast<const float4 *>(Y + (int64_t)idx * F) + sub)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      const float d = group_sum<C::L>(dot4(y, x));
      const float conf = fabsf(c);
      const float temp = (c > 0.f ? -2.f * c : 0.f) + (conf - 1.f) * d;
      if (sub == 0 && valid) {
        term += (double)(temp * d) + (double)conf;
        conf_sum += (double)conf;
      }
    }
  }
  // warp reduce then one atomic per warp
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    term += __shfl_xor_sync(0xffffffffu, term, m);
    conf_sum += __shfl_xor_sync(0xffffffffu, conf_sum, m);
  }
  if (lane == 0) {
    atomicAdd(out + 0, term);
    atomicAdd(out + 1, conf_sum);
  }
}

// out[2] = <A, B>_F, out[3] = trace(A), out[4] = trace(B) over the f x f leading blocks
__global__ void frob_trace_kernel(const float *__restrict__ A, const float *__restrict__ B, int f, int ld, double *out) {
  __shared__ double sh[3][256];
  double s = 0.0, ta = 0.0, tb = 0.0;
  for (int e = threadIdx.x; e < f * f; e += blockDim.x) {
    const int i = e / f, j = e % f;
    s += (double)A[i * ld + j] * (double)B[i * ld + j];
    if (i == j) {
      ta += A[i * ld + j];
      tb += B[i * ld + j];
    }
  }
  sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = ta; sh[2][threadIdx.x] = tb;
  __syncthreads();
  for (int m = 128; m > 0; m >>= 1) {
    if (threadIdx.x < m)
      for (int q = 0; q < 3; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + m];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[2] = sh[0][0]; out[3] = sh[1][0]; out[4] = sh[2][0];
  }
}

// 128 < padded factors <= 1024 (multiples of 128): one warp per row, NV float4 words per lane
template <int NV>
__global__ void __launch_bounds__(256)
loss_nnz_wide_kernel(const int32_t *__restrict__ indices, const float *__restrict__ data, const float *__restrict__ Y,
                     const float *__restrict__ X, int64_t row_offset, const WorkItem *__restrict__ work, int n_work,
                     int32_t *counter, double *out) {
  constexpr int F = 128 * NV;
  const int lane = threadIdx.x & 31;
  double term = 0.0, conf_sum = 0.0;
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(counter, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n_work) break;
    const int4 w = __ldg(reinterpret_cast<const int4 *>(work) + i);
    if (w.w == -2) continue;
    float4 x[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) x[q] = __ldg(reinterpret_cast<const float4 *>(X + (row_offset + w.x) * F) + lane + 32 * q);
    for (int k = w.y; k < w.z; ++k) {
      const int idx = __ldg(indices + k);
      const float c = __ldg(data + k);
      float d = 0.f;
#pragma unroll
      for (int q = 0; q < NV; ++q) d += dot4(__ldg(reinterpret_cast<const float4 *>(Y + (int64_t)idx * F) + lane + 32 * q), x[q]);
      d = group_sum<32>(d);
      const float conf = fabsf(c);
      const float temp = (c > 0.f ? -2.f * c : 0.f) + (conf - 1.f) * d;
      if (lane == 0) {
        term += (double)(temp * d) + (double)conf;
        conf_sum += (double)conf;
      }
    }
  }
  if (lane == 0) {
    atomicAdd(out + 0, term);
    atomicAdd(out + 1, conf_sum);
  }
}

__global__ void zero_loss_scalars(int32_t *counters, double *d) {
  if (threadIdx.x < 16) counters[threadIdx.x] = 0;
  if (threadIdx.x < 8) d[threadIdx.x] = 0.0;
}

template <int F>
int run_loss_nnz(als_ctx *ctx, const als_csr *C, const als_factors *X, const als_factors *Y) {
  if (!C->n_work) return ALS_OK;
  const int64_t want = ceil_div(C->n_work, 8);
  const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 8);
  ProfScope prof(ctx, kProfLoss);
  loss_nnz_kernel<F><<<grid, 256, 0, ctx->stream>>>(C->indices, C->data, Y->d, X->d, C->row_offset, C->work,
                                                    (int)C->n_work, ctx->counters, ctx->dscalars);
  ALS_CUDA(cudaGetLastError());
  ctx->launches++;
  return ALS_OK;
}

template <int NV>
int run_loss_nnz_wide(als_ctx *ctx, const als_csr *C, const als_factors *X, const als_factors *Y) {
  if (!C->n_work) return ALS_OK;
  const int grid = (int)std::min<int64_t>(ceil_div(C->n_work, 8), (int64_t)ctx->sm_count * 8);
  ProfScope prof(ctx, kProfLoss);
  loss_nnz_wide_kernel<NV><<<grid, 256, 0, ctx->stream>>>(C->indices, C->data, Y->d, X->d, C->row_offset, C->work, (int)C->n_work,
                                                          ctx->counters, ctx->dscalars);
  ALS_CUDA(cudaGetLastError());
  ctx->launches++;
  return ALS_OK;
}

}  // namespace

#define ALS_DISPATCH_F(ld, CALL)                                                              \
  switch ((ld) / 16) {                                                                        \
    case 1: return CALL(16);                                                                  \
    case 2: return CALL(32);                                                                  \
    case 3: return CALL(48);                                                                  \
    case 4: return CALL(64);                                                                  \
    case 5: return CALL(80);                                                                  \
    case 6: return CALL(96);                                                                  \
    case 7: return CALL(112);                                                                 \
    case 8: return CALL(128);                                                                 \
    default:                                                                                  \
      set_error("loss: factors padded to %d are not supported (<= 128 in steps of 16, <= 1024 in steps of 128)", (ld));                    \
      return ALS_E_UNSUPPORTED;                                                               \
  }

static int loss_nnz_dispatch(als_ctx *ctx, const als_csr *C, const als_factors *X, const als_factors *Y) {
  if (Y->ld > 128 && Y->ld % 128 == 0) {
    switch (Y->ld / 128) {
      case 2: return run_loss_nnz_wide<2>(ctx, C, X, Y);
      case 3: return run_loss_nnz_wide<3>(ctx, C, X, Y);
      case 4: return run_loss_nnz_wide<4>(ctx, C, X, Y);
      case 5: return run_loss_nnz_wide<5>(ctx, C, X, Y);
      case 6: return run_loss_nnz_wide<6>(ctx, C, X, Y);
      case 7: return run_loss_nnz_wide<7>(ctx, C, X, Y);
      case 8: return run_loss_nnz_wide<8>(ctx, C, X, Y);
      default: break;
    }
  }
#define CALL(FF) run_loss_nnz<FF>(ctx, C, X, Y)
  ALS_DISPATCH_F(Y->ld, CALL)
#undef CALL
}

// Requires ctx->G == Y^T Y on entry (als_calculate_loss computes it first).
int launch_loss(als_ctx *ctx, const als_csr *C, const als_factors *X, const als_factors *Y, float reg, double *loss) {
  if (X->ld != Y->ld) {
    set_error("loss: X and Y strides differ (%d vs %d)", X->ld, Y->ld);
    return ALS_E_INVALID;
  }
  const int ld = Y->ld;
  // keep Y^T Y in Greg, then overwrite G with X^T X restricted to C's rows
  ALS_CUDA(cudaMemcpyAsync(ctx->Greg, ctx->G, sizeof(float) * ld * ld, cudaMemcpyDeviceToDevice, ctx->stream));
  als_factors Xs = *X;
  Xs.d = X->d + C->row_offset * (int64_t)ld;
  Xs.rows = C->rows;
  int rc = launch_gramian(ctx, &Xs);
  if (rc != ALS_OK) return rc;
  zero_loss_scalars<<<1, 32, 0, ctx->stream>>>(ctx->counters, ctx->dscalars);
  ALS_CUDA(cudaGetLastError());
  frob_trace_kernel<<<1, 256, 0, ctx->stream>>>(ctx->Greg, ctx->G, Y->f, ld, ctx->dscalars);
  ALS_CUDA(cudaGetLastError());
  ctx->launches += 2;
  rc = loss_nnz_dispatch(ctx, C, X, Y);
  if (rc != ALS_OK) return rc;
  double h[8];
  ALS_CUDA(cudaMemcpyAsync(h, ctx->dscalars, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  ALS_CUDA(cudaStreamSynchronize(ctx->stream));
  // h[0] nnz terms, h[1] sum|c|, h[2] <YtY, XtX>, h[3] tr(YtY) = ||Y||^2, h[4] tr(XtX) = ||X_C||^2
  // loss[0] numerator (without the division), loss[1] total confidence, so that shards can be summed
  // by the host: loss = (sum num) / (sum conf + U*I - nnz)       (_als.pyx:307-308)
  loss[0] = h[2] + h[0] + (double)reg * h[4];
  loss[1] = h[1];
  loss[2] = (double)reg * h[3];  // item-norm part: identical on every shard, add once
  return ALS_OK;
}

}  // namespace als
