// R3: fused scores + filter + top-k (reference: topk.topk / _topk_batch, implicit/cpu/topk.pyx:15-67, and
// select<T>, implicit/cpu/select.h:12-39; the reference's GPU path is the unfused
// cublasSgemm -> thrust filters -> raft select_k of implicit/gpu/knn.cu:131-265).
//
// A CTA owns a block of QB query rows for the whole call and streams every item tile past them:
//   scores  a register-tiled fp32 GEMM of the query block against the tile (operands staged
//           transposed in shared memory so each k-step is three 16-byte loads per 32 FMAs);
//   filter  the global filter_items mask and the per-row "liked" CSR columns are overwritten with
//           -FLT_MAX in the staged score tile (topk.pyx:51-56);
//   select  each warp walks its rows of the score tile IN COLUMN ORDER against the row's running
//           threshold and keeps a sorted k-list in shared memory, applying exactly the reference's
//           admission rule (`size < k || score > min.score`, evict the lexicographic (score, col)
//           minimum), so ties resolve bit-for-bit like select.h.  The score matrix is never
//           written to HBM.
#include <float.h>
#include <limits.h>
#include <string.h>

#include "common.h"

namespace als {
namespace {

constexpr int kTopkThreads = 256;

template <int F, int TQ>
struct TkCfg {
  static constexpr int QB = 16 * TQ;             // query rows per CTA
  static constexpr int IT = (F <= 64) ? 128 : 64;  // items per tile
  static constexpr int TI = IT / 16;             // items per thread
  static constexpr int SLD = IT + 4;             // score tile stride
  static constexpr int ROWS_PER_WARP = QB / 8;
  static int smem_floats(int k) { return F * QB + F * IT + QB * SLD + 2 * QB * k + 2 * QB; }
};

__device__ __forceinline__ bool pair_less(float s, int c, float s2, int c2) { return s < s2 || (s == s2 && c < c2); }

// Warp-cooperative insertion of (s, col) into the ascending list (ls, lc)[0..cnt) with capacity k.
// Precondition (checked by the caller): cnt < k or s > ls[0].
__device__ __forceinline__ void list_insert(float *ls, int *lc, int &cnt, int k, float s, int col, int lane) {
  int pos = 0;
  for (int base = 0; base < cnt; base += 32) {
    const int i = base + lane;
    const bool lt = i < cnt && pair_less(ls[i], lc[i], s, col);
    pos += __popc(__ballot_sync(0xffffffffu, lt));
  }
  if (cnt < k) {
    for (int hi = cnt; hi > pos; hi -= 32) {  // shift [pos, cnt) up by one, top chunk first
      const int i = hi - 1 - lane;
      float ts = 0.f;
      int tc = 0;
      if (i >= pos) { ts = ls[i]; tc = lc[i]; }
      __syncwarp();
      if (i >= pos) { ls[i + 1] = ts; lc[i + 1] = tc; }
      __syncwarp();
    }
    if (lane == 0) { ls[pos] = s; lc[pos] = col; }
    ++cnt;
  } else {
    for (int lo = 1; lo < pos; lo += 32) {  // drop entry 0, shift [1, pos) down by one
      const int i = lo + lane;
      float ts = 0.f;
      int tc = 0;
      if (i < pos) { ts = ls[i]; tc = lc[i]; }
      __syncwarp();
      if (i < pos) { ls[i - 1] = ts; lc[i - 1] = tc; }
      __syncwarp();
    }
    if (lane == 0) { ls[pos - 1] = s; lc[pos - 1] = col; }
  }
  __syncwarp();
}

template <int F, int TQ>
__global__ void __launch_bounds__(kTopkThreads)
topk_kernel(const float *__restrict__ items, int n_items, const float *__restrict__ queries,
            const int32_t *__restrict__ query_rows, int n_query, int k, const float *__restrict__ item_norms,
            const uint8_t *__restrict__ item_mask, const int32_t *__restrict__ liked_indptr,
            const int32_t *__restrict__ liked_indices, int32_t *__restrict__ out_ids, float *__restrict__ out_scores) {
  using C = TkCfg<F, TQ>;
  constexpr int QB = C::QB, IT = C::IT, TI = C::TI, SLD = C::SLD;
  extern __shared__ __align__(16) float smem[];
  float *Qs = smem;                   // [F][QB]
  float *Is = Qs + F * QB;            // [F][IT]
  float *Ss = Is + F * IT;            // [QB][SLD]
  float *Ls = Ss + QB * SLD;          // [QB][k]
  int *Lc = reinterpret_cast<int *>(Ls + QB * k);  // [QB][k]
  int *Cnt = Lc + QB * k;             // [QB]
  int *Cur = Cnt + QB;                // [QB] liked-list cursors

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & 15, ty = tid >> 4;
  const float neginf = -FLT_MAX;

  for (int q0 = blockIdx.x * QB; q0 < n_query; q0 += gridDim.x * QB) {
    __syncthreads();
    // stage the query block transposed; rows past n_query are zero
    for (int e = tid; e < QB * (F / 4); e += kTopkThreads) {
      const int q = e % QB, fc = e / QB;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q0 + q < n_query) {
        const int64_t row = query_rows ? query_rows[q0 + q] : (q0 + q);
        v = __ldg(reinterpret_cast<const float4 *>(queries + row * F) + fc);
      }
      Qs[(4 * fc + 0) * QB + q] = v.x; Qs[(4 * fc + 1) * QB + q] = v.y;
      Qs[(4 * fc + 2) * QB + q] = v.z; Qs[(4 * fc + 3) * QB + q] = v.w;
    }
    for (int q = tid; q < QB; q += kTopkThreads) {
      Cnt[q] = 0;
      Cur[q] = (liked_indptr && q0 + q < n_query) ? liked_indptr[q0 + q] : 0;
    }

    for (int i0 = 0; i0 < n_items; i0 += IT) {
      __syncthreads();  // previous tile fully consumed
      for (int e = tid; e < IT * (F / 4); e += kTopkThreads) {
        const int it = e % IT, fc = e / IT;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i0 + it < n_items) v = __ldg(reinterpret_cast<const float4 *>(items + (int64_t)(i0 + it) * F) + fc);
        Is[(4 * fc + 0) * IT + it] = v.x; Is[(4 * fc + 1) * IT + it] = v.y;
        Is[(4 * fc + 2) * IT + it] = v.z; Is[(4 * fc + 3) * IT + it] = v.w;
      }
      __syncthreads();
      // ---- scores: thread (ty, tx) owns queries ty*TQ.. and items tx*TI..
      float acc[TQ][TI];
#pragma unroll
      for (int a = 0; a < TQ; ++a)
#pragma unroll
        for (int b = 0; b < TI; ++b) acc[a][b] = 0.f;
#pragma unroll 4
      for (int f = 0; f < F; ++f) {
        float qa[TQ], ib[TI];
#pragma unroll
        for (int a = 0; a < TQ; ++a) qa[a] = Qs[f * QB + ty * TQ + a];
#pragma unroll
        for (int b4 = 0; b4 < TI / 4; ++b4) {
          const float4 v = *reinterpret_cast<const float4 *>(Is + f * IT + tx * TI + 4 * b4);
          ib[4 * b4 + 0] = v.x; ib[4 * b4 + 1] = v.y; ib[4 * b4 + 2] = v.z; ib[4 * b4 + 3] = v.w;
        }
#pragma unroll
        for (int a = 0; a < TQ; ++a)
#pragma unroll
          for (int b = 0; b < TI; ++b) acc[a][b] = fmaf(qa[a], ib[b], acc[a][b]);
      }
      // ---- norms, global mask, stage the tile
#pragma unroll
      for (int b = 0; b < TI; ++b) {
        const int col = i0 + tx * TI + b;
        const bool in = col < n_items;
        const float nrm = (item_norms && in) ? __ldg(item_norms + col) : 1.f;
        const bool masked = item_mask && in && item_mask[col];
#pragma unroll
        for (int a = 0; a < TQ; ++a) {
          float s = acc[a][b];
          if (item_norms) s = __fdiv_rn(s, nrm);  // topk.pyx:48-49
          if (masked) s = neginf;                 // topk.pyx:55-56
          Ss[(ty * TQ + a) * SLD + tx * TI + b] = s;
        }
      }
      __syncthreads();
      // ---- per-row liked filter + ordered selection; a warp owns ROWS_PER_WARP rows throughout
      for (int rr = 0; rr < C::ROWS_PER_WARP; ++rr) {
        const int row = warp * C::ROWS_PER_WARP + rr;
        if (q0 + row >= n_query) break;
        float *srow = Ss + row * SLD;
        if (liked_indptr) {  // topk.pyx:51-53 (row indices sorted ascending by the host)
          const int end = liked_indptr[q0 + row + 1];
          int cur = Cur[row];
          for (;;) {
            const int p = cur + lane;
            const int idx = p < end ? __ldg(liked_indices + p) : INT_MAX;
            const bool hit = idx < i0 + IT;
            if (hit && idx >= i0) srow[idx - i0] = neginf;
            const int n = __popc(__ballot_sync(0xffffffffu, hit));
            cur += n;
            if (n < 32) break;
          }
          __syncwarp();
          if (lane == 0) Cur[row] = cur;
        }
        float *ls = Ls + row * k;
        int *lc = Lc + row * k;
        int cnt = Cnt[row];
        float thr = (cnt == k) ? ls[0] : neginf;
        for (int c0 = 0; c0 < IT; c0 += 32) {
          const int col = i0 + c0 + lane;
          const float s = srow[c0 + lane];
          unsigned m = __ballot_sync(0xffffffffu, col < n_items && (cnt < k || s > thr));
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1;
            const float cs = __shfl_sync(0xffffffffu, s, b);
            if (cnt < k || cs > thr) {  // select.h:23, re-checked against the updated threshold
              list_insert(ls, lc, cnt, k, cs, i0 + c0 + b, lane);
              thr = (cnt == k) ? ls[0] : neginf;
            }
          }
        }
        if (lane == 0) Cnt[row] = cnt;
        __syncwarp();
      }
    }
    // ---- emit: descending by (score, col) (select.h:33); the tail past cnt stays zero (topk.pyx:20-21)
    __syncthreads();
    for (int rr = 0; rr < C::ROWS_PER_WARP; ++rr) {
      const int row = warp * C::ROWS_PER_WARP + rr;
      if (q0 + row >= n_query) break;
      const int cnt = Cnt[row];
      for (int j = lane; j < cnt; j += 32) {
        out_ids[(int64_t)(q0 + row) * k + j] = Lc[row * k + cnt - 1 - j];
        out_scores[(int64_t)(q0 + row) * k + j] = Ls[row * k + cnt - 1 - j];
      }
    }
  }
}

__global__ void mask_scatter_kernel(uint8_t *mask, const int32_t *idx, int64_t n, int n_items) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int v = idx[i];
    if (v >= 0 && v < n_items) mask[v] = 1;
  }
}

struct TopkArgs {
  const float *items;
  int n_items;
  const float *queries;
  const int32_t *query_rows;
  int n_query;
  int k;        // effective k (<= n_items)
  int k_out;    // row stride of the outputs is k (same here: we run with k_out and clamp admission)
  const float *norms;
  const uint8_t *mask;
  const int32_t *liked_indptr;
  const int32_t *liked_indices;
  int32_t *ids;
  float *scores;
};

template <int F, int TQ>
int run_topk(als_ctx *ctx, const TopkArgs &a) {
  using C = TkCfg<F, TQ>;
  const int smem = C::smem_floats(a.k) * (int)sizeof(float);
  if (smem > 227 * 1024) {
    set_error("topk: k=%d needs %d bytes of shared memory per CTA (limit 227 KB)", a.k, smem);
    return ALS_E_UNSUPPORTED;
  }
  auto kern = topk_kernel<F, TQ>;
  ALS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t blocks = ceil_div(a.n_query, C::QB);
  const int grid = (int)std::min<int64_t>(blocks, (int64_t)ctx->sm_count * 2);
  ProfScope prof(ctx, kProfTopk);
  kern<<<grid, kTopkThreads, smem, ctx->stream>>>(a.items, a.n_items, a.queries, a.query_rows, a.n_query, a.k, a.norms,
                                                  a.mask, a.liked_indptr, a.liked_indices, a.ids, a.scores);
  ALS_CUDA(cudaGetLastError());
  ctx->launches++;
  return ALS_OK;
}

template <int F>
int run_topk_f(als_ctx *ctx, const TopkArgs &a) {
  if (a.k <= 64) return run_topk<F, 4>(ctx, a);
  return run_topk<F, 1>(ctx, a);
}

}  // namespace

int launch_topk(als_ctx *ctx, const als_factors *items, const als_factors *queries, const int32_t *query_rows,
                int64_t n_query, int k, const float *item_norms_host, const als_csr *liked,
                const int32_t *filter_items, int64_t n_filter, int32_t *ids_host, float *scores_host) {
  if (items->ld != queries->ld) {
    set_error("topk: items and queries strides differ");
    return ALS_E_INVALID;
  }
  if (n_query == 0 || k == 0) return ALS_OK;
  memset(ids_host, 0, sizeof(int32_t) * n_query * k);      // topk.pyx:20-21
  memset(scores_host, 0, sizeof(float) * n_query * k);
  const int64_t I = items->rows;
  if (I == 0) return ALS_OK;
  if (n_query >= INT32_MAX || I >= INT32_MAX) {
    set_error("topk: too many rows");
    return ALS_E_UNSUPPORTED;
  }
  // The output row stride is k; admission is clamped to the number of items (the k-list can never
  // hold more than I entries), which leaves the tail zero exactly like the reference.
  const int k_eff = (int)std::min<int64_t>(k, I);
  // device staging: [ids | scores | query_rows | norms | mask | filter list]
  const int64_t out_elems = n_query * (int64_t)k_eff;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const int64_t o_ids = take(out_elems * 4), o_sc = take(out_elems * 4);
  const int64_t o_qr = take(query_rows ? n_query * 4 : 0);
  const int64_t o_nrm = take(item_norms_host ? I * 4 : 0);
  const int64_t o_mask = take(n_filter ? I : 0);
  const int64_t o_fl = take(n_filter * 4);
  int rc = ensure_scratch(ctx, off);
  if (rc != ALS_OK) return rc;
  char *base = (char *)ctx->scratch;
  ALS_CUDA(cudaMemsetAsync(base + o_ids, 0, out_elems * 4, ctx->stream));
  ALS_CUDA(cudaMemsetAsync(base + o_sc, 0, out_elems * 4, ctx->stream));
  if (query_rows) {
    for (int64_t q = 0; q < n_query; ++q)
      if (query_rows[q] < 0 || query_rows[q] >= queries->rows) {
        set_error("topk: query row %d out of range", query_rows[q]);
        return ALS_E_INVALID;
      }
    ALS_CUDA(cudaMemcpyAsync(base + o_qr, query_rows, n_query * 4, cudaMemcpyHostToDevice, ctx->stream));
  } else if (n_query > queries->rows) {
    set_error("topk: %lld queries requested from a matrix of %lld rows", (long long)n_query, (long long)queries->rows);
    return ALS_E_INVALID;
  }
  if (item_norms_host)
    ALS_CUDA(cudaMemcpyAsync(base + o_nrm, item_norms_host, I * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (n_filter) {
    ALS_CUDA(cudaMemsetAsync(base + o_mask, 0, I, ctx->stream));
    ALS_CUDA(cudaMemcpyAsync(base + o_fl, filter_items, n_filter * 4, cudaMemcpyHostToDevice, ctx->stream));
    mask_scatter_kernel<<<(unsigned)ceil_div(n_filter, 256), 256, 0, ctx->stream>>>(
        (uint8_t *)(base + o_mask), (const int32_t *)(base + o_fl), n_filter, (int)I);
    ALS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  TopkArgs a;
  a.items = items->d;
  a.n_items = (int)I;
  a.queries = queries->d;
  a.query_rows = query_rows ? (const int32_t *)(base + o_qr) : nullptr;
  a.n_query = (int)n_query;
  a.k = k_eff;
  a.k_out = k;
  a.norms = item_norms_host ? (const float *)(base + o_nrm) : nullptr;
  a.mask = n_filter ? (const uint8_t *)(base + o_mask) : nullptr;
  a.liked_indptr = liked ? liked->indptr : nullptr;
  a.liked_indices = liked ? liked->indices : nullptr;
  a.ids = (int32_t *)(base + o_ids);
  a.scores = (float *)(base + o_sc);
#define CALL(FF) run_topk_f<FF>(ctx, a)
  switch (items->ld / 16) {
    case 1: rc = CALL(16); break;
    case 2: rc = CALL(32); break;
    case 3: rc = CALL(48); break;
    case 4: rc = CALL(64); break;
    case 5: rc = CALL(80); break;
    case 6: rc = CALL(96); break;
    case 7: rc = CALL(112); break;
    case 8: rc = CALL(128); break;
    default:
      set_error("topk: factors padded to %d > 128 are not supported yet", items->ld);
      return ALS_E_UNSUPPORTED;
  }
#undef CALL
  if (rc != ALS_OK) return rc;
  // outputs: device [n_query, k_eff] -> host [n_query, k]
  ALS_CUDA(cudaMemcpy2DAsync(ids_host, sizeof(int32_t) * k, base + o_ids, sizeof(int32_t) * k_eff,
                             sizeof(int32_t) * k_eff, n_query, cudaMemcpyDeviceToHost, ctx->stream));
  ALS_CUDA(cudaMemcpy2DAsync(scores_host, sizeof(float) * k, base + o_sc, sizeof(float) * k_eff,
                             sizeof(float) * k_eff, n_query, cudaMemcpyDeviceToHost, ctx->stream));
  ALS_CUDA(cudaStreamSynchronize(ctx->stream));
  return ALS_OK;
}

}  // namespace als
