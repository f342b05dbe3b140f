// Input preparation on the device: CSR transpose (replaces the host `Ciu = Cui.T.tocsr()`,
// implicit/cpu/als.py:137, which costs seconds at 17M-500M nonzeros and would dominate an
// end-to-end fit once the solve itself takes milliseconds).
//
// Transposing a CSR is a STABLE sort of its entries by column: a stable LSD radix sort of
// (column, entry position) pairs yields, per column, the entries in increasing row order, i.e.
// exactly scipy's canonical result, deterministically.  The radix sort is CUB's (toolkit library
// code; this is one-time input preparation, not the per-iteration hot path).
#include <cub/device/device_radix_sort.cuh>

#include "common.h"

namespace als {
namespace {

__global__ void iota_kernel(int32_t *v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[i] = (int32_t)i;
}

// For each sorted entry j: the row of the original entry e = perm[j] (binary search in indptr) and its value.
__global__ void gather_transposed_kernel(const int32_t *__restrict__ perm, const int32_t *__restrict__ indptr,
                                         int rows, const float *__restrict__ data, int64_t nnz,
                                         int32_t *__restrict__ out_indices, float *__restrict__ out_data) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; j < nnz; j += stride) {
    const int32_t e = perm[j];
    int lo = 0, hi = rows;  // largest r with indptr[r] <= e
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (indptr[mid] <= e) lo = mid; else hi = mid;
    }
    out_indices[j] = lo;
    out_data[j] = data[e];
  }
}

// out_indptr[c] = first position j with sorted_cols[j] >= c
__global__ void indptr_from_sorted_kernel(const int32_t *__restrict__ sorted_cols, int64_t nnz, int cols,
                                          int32_t *__restrict__ out_indptr) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > cols) return;
  int64_t lo = 0, hi = nnz;  // first j in [0, nnz] with key >= c
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (sorted_cols[mid] < c) lo = mid + 1; else hi = mid;
  }
  out_indptr[c] = (int32_t)lo;
}

}  // namespace

int csr_transpose(als_ctx *ctx, const als_csr *in, als_csr **out) {
  *out = nullptr;
  if (in->row_offset != 0) {
    set_error("als_csr_transpose: row shards cannot be transposed");
    return ALS_E_INVALID;
  }
  ALS_CUDA(cudaSetDevice(ctx->device));
  const int64_t nnz = in->nnz;
  const int rows = (int)in->rows, cols = (int)in->cols;
  als_csr *t = new als_csr();
  t->ctx = ctx;
  t->rows = cols;
  t->cols = rows;
  t->nnz = nnz;
  int rc;
  if ((rc = dev_alloc(ctx, (void **)&t->indptr, sizeof(int32_t) * ((int64_t)cols + 1))) != ALS_OK ||
      (rc = dev_alloc(ctx, (void **)&t->indices, sizeof(int32_t) * std::max<int64_t>(nnz, 1))) != ALS_OK ||
      (rc = dev_alloc(ctx, (void **)&t->data, sizeof(float) * std::max<int64_t>(nnz, 1))) != ALS_OK) {
    als_csr_destroy(t);
    return rc;
  }
  // one pinned landing buffer per context: a transposed matrix whose schedule is still pending owns it
  if (ctx->sched_owner && (rc = ensure_schedule(ctx, ctx->sched_owner)) != ALS_OK) {
    als_csr_destroy(t);
    return rc;
  }
  const int64_t need = sizeof(int32_t) * ((int64_t)cols + 1);
  if (need > ctx->sched_pinned_cap) {
    if (ctx->sched_pinned) ALS_CUDA(cudaFreeHost(ctx->sched_pinned));
    ctx->sched_pinned = nullptr;
    ctx->sched_pinned_cap = 0;
    ALS_CUDA(cudaMallocHost((void **)&ctx->sched_pinned, (size_t)need));
    ctx->sched_pinned_cap = need;
  }
  if (nnz > 0) {
    int32_t *keys_out = nullptr, *vals_in = nullptr, *vals_out = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    int end_bit = 1;
    while (end_bit < 32 && (1ll << end_bit) < (long long)cols) ++end_bit;
    ALS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, in->indices, keys_out, vals_in, vals_out, (int)nnz, 0,
                                             end_bit, ctx->stream));
    if ((rc = dev_alloc(ctx, (void **)&keys_out, sizeof(int32_t) * nnz)) != ALS_OK ||
        (rc = dev_alloc(ctx, (void **)&vals_in, sizeof(int32_t) * nnz)) != ALS_OK ||
        (rc = dev_alloc(ctx, (void **)&vals_out, sizeof(int32_t) * nnz)) != ALS_OK ||
        (rc = dev_alloc(ctx, &tmp, (int64_t)tmp_bytes)) != ALS_OK) {
      als_csr_destroy(t);
      return rc;
    }
    iota_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(vals_in, nnz);
    ALS_CUDA(cudaGetLastError());
    ALS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, in->indices, keys_out, vals_in, vals_out, (int)nnz, 0,
                                             end_bit, ctx->stream));
    gather_transposed_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(vals_out, in->indptr, rows, in->data, nnz,
                                                                        t->indices, t->data);
    ALS_CUDA(cudaGetLastError());
    indptr_from_sorted_kernel<<<(cols + 1 + 255) / 256, 256, 0, ctx->stream>>>(keys_out, nnz, cols, t->indptr);
    ALS_CUDA(cudaGetLastError());
    ctx->launches += 3;
    dev_free(ctx, keys_out);  // stream ordered: after the kernels above
    dev_free(ctx, vals_in);
    dev_free(ctx, vals_out);
    dev_free(ctx, tmp);
  } else {
    ALS_CUDA(cudaMemsetAsync(t->indptr, 0, sizeof(int32_t) * ((int64_t)cols + 1), ctx->stream));
  }
  // The schedule needs the row lengths on the host.  The copy is left in flight and the schedule is built at the
  // first solve over this matrix (ensure_schedule): in a fit that is the item half, so the host-side sort of the
  // rows overlaps the user half that is already running instead of leaving the GPU idle.
  ALS_CUDA(cudaMemcpyAsync(ctx->sched_pinned, t->indptr, (size_t)need, cudaMemcpyDeviceToHost, ctx->stream));
  ALS_CUDA(cudaEventRecord(ctx->sched_ev, ctx->stream));
  t->sched_pending = true;
  ctx->sched_owner = t;
  *out = t;
  return ALS_OK;
}

}  // namespace als
