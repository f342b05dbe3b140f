// R2: fused conjugate-gradient half-iteration (reference: _least_squares_cg, implicit/cpu/_als.pyx:154-248;
// the reference's own GPU kernel is least_squares_cg_kernel, implicit/gpu/als.cu:23-111).
// R6: training loss (reference: _calculate_loss, implicit/cpu/_als.pyx:259-308).
//
// Matrix-free CG, (1 + cg_steps) passes over the row's nonzeros.  A warp owns a row; inside the warp
// sub-groups of L = F/4 lanes each take one nonzero at a time: lane `sub` of a group holds the
// float4 slice [4 sub, 4 sub + 4) of every CG vector (x, r, p, Ap are replicated per group), so a
// factor row is read with one coalesced 16-byte load per lane and a dot product costs log2(L)
// shuffles.  Giant rows (more than kSplitNnz nonzeros) are handled by a whole CTA per row with a
// fixed-order cross-warp reduction, so a power-law hub cannot serialise the tail.
#include "common.h"

namespace als {

namespace {

constexpr int kCgWarps = 8;        // warps per CTA, warp-per-row kernel
constexpr int kCgGiantWarps = 16;  // warps per CTA, CTA-per-row kernel

template <int F>
struct CgCfg {
  static constexpr int V = F / 4;  // float4 slices per factor row
  static constexpr int L = V <= 4 ? 4 : V <= 8 ? 8 : V <= 16 ? 16 : 32;  // lanes per group
  static constexpr int NG = 32 / L;                                     // groups per warp
  static_assert(V <= 32, "CG kernel handles padded factors <= 128");
};

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void axpy4(float4 &y, float a, const float4 &x) {
  y.x = fmaf(a, x.x, y.x); y.y = fmaf(a, x.y, y.y); y.z = fmaf(a, x.z, y.z); y.w = fmaf(a, x.w, y.w);
}
__device__ __forceinline__ float4 shfl_xor4(const float4 &v, int m) {
  return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                     __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
__device__ __forceinline__ void add4(float4 &a, const float4 &b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// sum over the L lanes of a group (all lanes of the group get the same bits)
template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int m = 1; m < L; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
// sum a per-group float4 over the NG groups of the warp
template <int L>
__device__ __forceinline__ void across_groups(float4 &v) {
#pragma unroll
  for (int m = L; m < 32; m <<= 1) add4(v, shfl_xor4(v, m));
}

// Team = the warps that cooperate on one row (1 for the warp-per-row kernel, the CTA for giants).
template <int F, int TW>
struct Team {
  using C = CgCfg<F>;
  int lane, sub, grp, warp;  // warp = index inside the team
  bool active;
  float *xs;    // per-warp F floats: vector broadcast for the symv
  float *red;   // TW > 1: [TW][F] cross-warp reduction buffer

  // out = G v  (G symmetric F x F, row-major, L1/L2 resident).  v is replicated in every group.
  __device__ __forceinline__ float4 symv(const float *__restrict__ G, const float4 &v) const {
    if (grp == 0 && active) reinterpret_cast<float4 *>(xs)[sub] = v;
    __syncwarp();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
      for (int j = 4 * grp; j < F; j += 4 * C::NG) {
        const float4 xv = *reinterpret_cast<const float4 *>(xs + j);
        axpy4(acc, xv.x, __ldg(reinterpret_cast<const float4 *>(G + (j + 0) * F) + sub));
        axpy4(acc, xv.y, __ldg(reinterpret_cast<const float4 *>(G + (j + 1) * F) + sub));
        axpy4(acc, xv.z, __ldg(reinterpret_cast<const float4 *>(G + (j + 2) * F) + sub));
        axpy4(acc, xv.w, __ldg(reinterpret_cast<const float4 *>(G + (j + 3) * F) + sub));
      }
    }
    across_groups<C::L>(acc);
    __syncwarp();
    return acc;
  }

  // total over every group of every warp of the team; identical bits in all lanes of the team
  __device__ __forceinline__ void team_sum(float4 &v) const {
    across_groups<C::L>(v);
    if (TW > 1) {
      if (grp == 0 && active) reinterpret_cast<float4 *>(red + warp * F)[sub] = v;
      __syncthreads();
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active)
        for (int w = 0; w < TW; ++w) add4(s, reinterpret_cast<const float4 *>(red + w * F)[sub]);
      __syncthreads();
      v = s;
    }
  }
};

// One pass over the nonzeros [k0, k1):  acc += coef_k * y_k  with coef_k = pos_k - (|c_k| - 1) (y_k . v)
// where pos_k = c_k if (first pass and c_k > 0) else 0.   (_als.pyx:190-201 and :214-222)
template <int F, int TW, bool FIRST>
__device__ __forceinline__ float4 nnz_pass(const Team<F, TW> &tm, const int32_t *__restrict__ indices,
                                           const float *__restrict__ data, const float *__restrict__ Y, int k0, int k1,
                                           const float4 &v, float sign) {
  using C = CgCfg<F>;
  constexpr int TG = TW * C::NG;
  const int gid = tm.warp * C::NG + tm.grp;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int UN = 4;
  // The trip count must be uniform across the warp (the group reductions are full-warp shuffles):
  // every group walks the same kb and masks its own out-of-range nonzeros to y = 0.
  for (int kb = k0; kb < k1; kb += UN * TG) {
    float4 y[UN];
    float c[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int k = kb + u * TG + gid;
      const bool valid = k < k1;
      const int idx = valid ? __ldg(indices + k) : 0;
      c[u] = valid ? __ldg(data + k) : 0.f;
      y[u] = (tm.active && valid) ? __ldg(reinterpret_cast<const float4 *>(Y + (int64_t)idx * F) + tm.sub)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const float d = group_sum<C::L>(dot4(y[u], v));
      const float conf = fabsf(c[u]);
      const float pos = (FIRST && c[u] > 0.f) ? c[u] : 0.f;
      const float coef = pos - sign * (conf - 1.f) * d;
      axpy4(acc, coef, y[u]);  // masked nonzeros have y == 0
    }
  }
  return acc;
}

template <int F, int TW>
__device__ __forceinline__ void cg_row(const Team<F, TW> &tm, const int32_t *__restrict__ indices,
                                       const float *__restrict__ data, const float *__restrict__ Y,
                                       float *__restrict__ xrow, const float *__restrict__ Greg, int k0, int k1,
                                       int cg_steps) {
  using C = CgCfg<F>;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (k0 == k1) {  // no observations: zero the row (_als.pyx:182-184)
    if (tm.warp == 0 && tm.grp == 0 && tm.active) reinterpret_cast<float4 *>(xrow)[tm.sub] = zero4;
    return;
  }
  float4 x = tm.active ? reinterpret_cast<const float4 *>(xrow)[tm.sub] : zero4;  // warm start (:179)
  // r = -(YtY + lambda I) x + sum_k (c_k^+ - (|c_k| - 1) y_k.x) y_k      (:187-201)
  float4 r = tm.symv(Greg, x);
  r.x = -r.x; r.y = -r.y; r.z = -r.z; r.w = -r.w;
  {
    float4 a = nnz_pass<F, TW, true>(tm, indices, data, Y, k0, k1, x, 1.f);
    tm.team_sum(a);
    add4(r, a);
  }
  float4 p = r;
  float rsold = group_sum<C::L>(dot4(r, r));
  if (rsold < 1e-20f) return;  // :206-207 (x stays as it is)
  for (int it = 0; it < cg_steps; ++it) {
    // Ap = (YtY + lambda I) p + sum_k (|c_k| - 1) (y_k.p) y_k           (:212-222)
    float4 Ap = tm.symv(Greg, p);
    {
      float4 a = nnz_pass<F, TW, false>(tm, indices, data, Y, k0, k1, p, -1.f);
      tm.team_sum(a);
      add4(Ap, a);
    }
    const float alpha = rsold / group_sum<C::L>(dot4(p, Ap));  // :225
    axpy4(x, alpha, p);                                       // :228
    axpy4(r, -alpha, Ap);                                     // :231-232
    const float rsnew = group_sum<C::L>(dot4(r, r));          // :234
    if (rsnew < 1e-20f) break;                                // :235-236
    const float beta = rsnew / rsold;                         // :239-242
    p.x = fmaf(beta, p.x, r.x); p.y = fmaf(beta, p.y, r.y); p.z = fmaf(beta, p.z, r.z); p.w = fmaf(beta, p.w, r.w);
    rsold = rsnew;
  }
  if (tm.warp == 0 && tm.grp == 0 && tm.active) reinterpret_cast<float4 *>(xrow)[tm.sub] = x;
}

template <int F>
__global__ void __launch_bounds__(32 * kCgWarps)
cg_half_kernel(const int32_t *__restrict__ indices, const float *__restrict__ data, const float *__restrict__ Y,
               float *__restrict__ X, int64_t row_offset, const float *__restrict__ Greg,
               const WorkItem *__restrict__ work, int n_work, int32_t *counter, int cg_steps) {
  using C = CgCfg<F>;
  __shared__ __align__(16) float xs_all[kCgWarps][F];
  Team<F, 1> tm;
  tm.lane = threadIdx.x & 31;
  tm.sub = tm.lane % C::L;
  tm.grp = tm.lane / C::L;
  tm.warp = 0;
  tm.active = tm.sub < C::V;
  tm.xs = xs_all[threadIdx.x >> 5];
  tm.red = nullptr;
  for (;;) {
    int i = 0;
    if (tm.lane == 0) i = atomicAdd(counter, 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n_work) break;
    const int4 w = __ldg(reinterpret_cast<const int4 *>(work) + i);
    if (w.w != -1) continue;  // chunks of giant rows: the CTA-per-row kernel owns those rows
    cg_row<F, 1>(tm, indices, data, Y, X + (row_offset + w.x) * F, Greg, w.y, w.z, cg_steps);
  }
}

template <int F>
__global__ void __launch_bounds__(32 * kCgGiantWarps)
cg_giant_kernel(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                const float *__restrict__ Y, float *__restrict__ X, int64_t row_offset, const float *__restrict__ Greg,
                const WorkItem *__restrict__ finish, int n_finish, int cg_steps) {
  using C = CgCfg<F>;
  __shared__ __align__(16) float xs_all[kCgGiantWarps][F];
  __shared__ __align__(16) float red[kCgGiantWarps][F];
  Team<F, kCgGiantWarps> tm;
  tm.lane = threadIdx.x & 31;
  tm.sub = tm.lane % C::L;
  tm.grp = tm.lane / C::L;
  tm.warp = threadIdx.x >> 5;
  tm.active = tm.sub < C::V;
  tm.xs = xs_all[tm.warp];
  tm.red = &red[0][0];
  for (int i = blockIdx.x; i < n_finish; i += gridDim.x) {
    const int row = finish[i].row;
    cg_row<F, kCgGiantWarps>(tm, indices, data, Y, X + (row_offset + row) * F, Greg, indptr[row], indptr[row + 1],
                             cg_steps);
    __syncthreads();
  }
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

__global__ void zero_scalars(int32_t *counters, double *d) {
  if (threadIdx.x < 16) counters[threadIdx.x] = 0;
  if (threadIdx.x < 8) d[threadIdx.x] = 0.0;
}

template <int F>
int run_cg(als_ctx *ctx, const als_csr *C, als_factors *X, const als_factors *Y, int cg_steps) {
  zero_scalars<<<1, 32, 0, ctx->stream>>>(ctx->counters, ctx->dscalars);
  ALS_CUDA(cudaGetLastError());
  ctx->launches++;
  if (C->n_work) {
    const int64_t want = ceil_div(C->n_work, kCgWarps);
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * 8);
    ProfScope prof(ctx, kProfCg);
    cg_half_kernel<F><<<grid, 32 * kCgWarps, 0, ctx->stream>>>(C->indices, C->data, Y->d, X->d, C->row_offset, ctx->Greg,
                                                              C->work, (int)C->n_work, ctx->counters, cg_steps);
    ALS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (C->n_finish) {
    const int grid = (int)std::min<int64_t>(C->n_finish, (int64_t)ctx->sm_count * 2);
    ProfScope prof(ctx, kProfCgGiant);
    cg_giant_kernel<F><<<grid, 32 * kCgGiantWarps, 0, ctx->stream>>>(C->indptr, C->indices, C->data, Y->d, X->d,
                                                                    C->row_offset, ctx->Greg, C->finish,
                                                                    (int)C->n_finish, cg_steps);
    ALS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  return ALS_OK;
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
      set_error("factors padded to %d > 128 are not supported yet", (ld));                    \
      return ALS_E_UNSUPPORTED;                                                               \
  }

int launch_cg(als_ctx *ctx, const als_csr *C, als_factors *X, const als_factors *Y, int cg_steps) {
  if (X->ld != Y->ld) {
    set_error("cg: X and Y strides differ (%d vs %d)", X->ld, Y->ld);
    return ALS_E_INVALID;
  }
#define CALL(FF) run_cg<FF>(ctx, C, X, Y, cg_steps)
  ALS_DISPATCH_F(Y->ld, CALL)
#undef CALL
}

static int loss_nnz_dispatch(als_ctx *ctx, const als_csr *C, const als_factors *X, const als_factors *Y) {
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
  zero_scalars<<<1, 32, 0, ctx->stream>>>(ctx->counters, ctx->dscalars);
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
