/*
 * oracle/als_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement ("port") of the reference's CPU ALS hot path, used exclusively as the
 * checker for the CUDA product (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference legs).  The product path (implicit_b200/) never links, imports or calls this.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (1) the reference's own compiled Cython (oracle/_ref, built by oracle/build_ref.py from
 *       /root/reference/implicit/cpu/{_als,topk}.pyx) when it is present, and
 *   (2) the committed golden vectors under tests/golden/ that were generated from that build
 *       (tests/golden/make_golden.py), and
 *   (3) the known-answer tests the reference holds for this path
 *       (tests/als_test.py:142-186 test_factorize, :304-324 test_calculate_loss_simple).
 *
 * The reference delegates its inner arithmetic to BLAS/LAPACK through scipy.linalg.cython_blas /
 * cython_lapack (saxpy, sdot, ssymv, sscal, sposv; scipy>=0.16 unpinned, OpenBLAS in this image).
 * Those are restated here as straight fp32 loops, so results agree with the reference to fp32
 * rounding (summation order inside OpenBLAS kernels differs), not bit for bit.
 *
 * Each function cites the reference file:line it follows (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ---- BLAS-1/2 restated (implicit/cpu/_als.pyx:19-57 wrappers) ------------------------------ */
static inline void s_axpy(int n, float a, const float *x, float *y) {
  for (int i = 0; i < n; ++i) y[i] += a * x[i];
}
/* sdot: OpenBLAS's x86 kernels keep 32 interleaved fp32 partial sums (4 AVX2 accumulators of 8 lanes)
 * and add them pairwise at the end; a single running sum would be measurably LESS accurate than the
 * reference on all-positive data (the first ALS half-iteration), so the restatement mirrors that. */
static inline float s_dot(int n, const float *x, const float *y) {
  float acc[32];
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  int i = 0;
  for (; i + 32 <= n; i += 32)
    for (int j = 0; j < 32; ++j) acc[j] += x[i + j] * y[i + j];
  for (int j = 0; i < n; ++i, ++j) acc[j] += x[i] * y[i];
  for (int w = 16; w >= 1; w >>= 1)
    for (int j = 0; j < w; ++j) acc[j] += acc[j + w];
  return acc[0];
}
/* y = alpha * A x with A symmetric (ssymv 'U', beta = 0); A is stored full so rows are used. */
static inline void s_symv(int n, float alpha, const float *A, const float *x, float *y) {
  for (int i = 0; i < n; ++i) y[i] = alpha * s_dot(n, A + (size_t)i * n, x);
}

/*
 * sposv('U', n, 1, A, n, b, n): Cholesky A = U^T U on the triangle LAPACK calls "upper" of the
 * column-major view.  The reference hands LAPACK a row-major symmetric buffer (implicit/cpu/_als.pyx:127),
 * so which triangle is read is immaterial mathematically; we factor the row-major lower triangle
 * (== column-major upper) in place, then forward/back substitute.  Returns 0 or the 1-based index
 * of the first non-positive pivot (LAPACK info).
 */
static int s_posv(int n, float *A, float *b) {
  for (int j = 0; j < n; ++j) {
    float *Aj = A + (size_t)j * n;
    float d = Aj[j];
    for (int k = 0; k < j; ++k) d -= Aj[k] * Aj[k];
    if (!(d > 0.f)) return j + 1;
    d = sqrtf(d);
    Aj[j] = d;
    for (int i = j + 1; i < n; ++i) {
      float *Ai = A + (size_t)i * n;
      float s = Ai[j];
      for (int k = 0; k < j; ++k) s -= Ai[k] * Aj[k];
      Ai[j] = s / d;
    }
  }
  /* L z = b */
  for (int i = 0; i < n; ++i) {
    const float *Ai = A + (size_t)i * n;
    float s = b[i];
    for (int k = 0; k < i; ++k) s -= Ai[k] * b[k];
    b[i] = s / Ai[i];
  }
  /* L^T x = z */
  for (int i = n - 1; i >= 0; --i) {
    float s = b[i];
    for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  return 0;
}

/* ---- R4: Gramian  YtY = Y^T Y  (implicit/cpu/_als.pyx:70, :164, :268 -- np.dot = sgemm) ------- */
ORACLE_API void oracle_gramian(const float *Y, int64_t rows, int f, float *G) {
  double *acc = (double *)calloc((size_t)f * f, sizeof(double));
  for (int64_t r = 0; r < rows; ++r) {
    const float *y = Y + r * f;
    for (int i = 0; i < f; ++i) {
      const double yi = y[i];
      for (int j = 0; j < f; ++j) acc[(size_t)i * f + j] += yi * (double)y[j];
    }
  }
  /* rounded once to fp32: the best fp32 representative of what sgemm approximates */
  for (int i = 0; i < f * f; ++i) G[i] = (float)acc[i];
  free(acc);
}

/* ---- R1: Cholesky half  (implicit/cpu/_als.pyx:76-142 _least_squares) ------------------------ */
/* YtY is WITHOUT lambda (added here, :85).  Returns -1 on success or the first failing row (:131-138). */
ORACLE_API int64_t oracle_least_squares(const float *YtY, const int32_t *indptr, const int32_t *indices,
                                        const float *data, float *X, const float *Y, int64_t users,
                                        int f, double regularization, int num_threads) {
  float *initialA = (float *)malloc(sizeof(float) * f * f);
  for (int i = 0; i < f; ++i)
    for (int j = 0; j < f; ++j)
      initialA[i * f + j] = YtY[i * f + j] + (i == j ? (float)regularization : 0.0f); /* :85 (fp32 add, numpy weak-scalar) */
  int64_t bad = -1;
#ifdef _OPENMP
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel num_threads(num_threads)
#endif
  {
    float *A = (float *)malloc(sizeof(float) * f * f); /* :93 */
    float *b = (float *)malloc(sizeof(float) * f);     /* :94 */
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8) /* :96 */
#endif
    for (int64_t u = 0; u < users; ++u) {
      if (indptr[u] == indptr[u + 1]) { /* :98-100 empty row -> zeros */
        memset(X + u * f, 0, sizeof(float) * f);
        continue;
      }
      memcpy(A, initialA, sizeof(float) * f * f); /* :106 */
      memset(b, 0, sizeof(float) * f);            /* :107 */
      for (int32_t idx = indptr[u]; idx < indptr[u + 1]; ++idx) {
        const float *yi = Y + (int64_t)indices[idx] * f;
        float confidence = data[idx];
        if (confidence > 0) s_axpy(f, confidence, yi, b); /* :115-116 */
        else confidence = -1 * confidence;               /* :117-118 */
        for (int j = 0; j < f; ++j) {                     /* :122-124 */
          float temp = (confidence - 1) * yi[j];
          s_axpy(f, temp, yi, A + (size_t)j * f);
        }
      }
      int err = s_posv(f, A, b); /* :127 */
      if (!err) memcpy(X + u * f, b, sizeof(float) * f); /* :129-130 */
      else {
#ifdef _OPENMP
#pragma omp critical
#endif
        { if (bad < 0 || u < bad) bad = u; }
      }
    }
    free(A);
    free(b);
  }
  free(initialA);
  return bad;
}

/* ---- R2: CG half  (implicit/cpu/_als.pyx:154-248 _least_squares_cg) -------------------------- */
/* YtY_reg = Y^T Y + lambda I as float (:155 `float regularization`, :164).  X updated in place. */
ORACLE_API void oracle_least_squares_cg(const float *YtY_reg, const int32_t *indptr, const int32_t *indices,
                                        const float *data, float *X, const float *Y, int64_t users, int N,
                                        int cg_steps, int num_threads) {
#ifdef _OPENMP
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel num_threads(num_threads)
#endif
  {
    float *Ap = (float *)malloc(sizeof(float) * N);
    float *p = (float *)malloc(sizeof(float) * N);
    float *r = (float *)malloc(sizeof(float) * N);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8) /* :177 */
#endif
    for (int64_t u = 0; u < users; ++u) {
      float *x = X + u * N; /* :179 warm start, in place */
      if (indptr[u] == indptr[u + 1]) { /* :182-184 */
        memset(x, 0, sizeof(float) * N);
        continue;
      }
      s_symv(N, -1.0f, YtY_reg, x, r); /* :187-188 r = -YtY x */
      for (int32_t idx = indptr[u]; idx < indptr[u + 1]; ++idx) { /* :190-201 */
        const float *yi = Y + (int64_t)indices[idx] * N;
        float confidence = data[idx], temp;
        if (confidence > 0) temp = confidence;
        else { temp = 0; confidence = -1 * confidence; }
        temp = temp - (confidence - 1) * s_dot(N, yi, x);
        s_axpy(N, temp, yi, r);
      }
      memcpy(p, r, sizeof(float) * N); /* :203 */
      float rsold = s_dot(N, r, r);    /* :204 */
      if (rsold < 1e-20f) continue;    /* :206-207 */
      for (int it = 0; it < cg_steps; ++it) { /* :209 */
        s_symv(N, 1.0f, YtY_reg, p, Ap);      /* :212 */
        for (int32_t idx = indptr[u]; idx < indptr[u + 1]; ++idx) { /* :214-222 */
          const float *yi = Y + (int64_t)indices[idx] * N;
          float confidence = data[idx];
          if (confidence < 0) confidence = -1 * confidence;
          float temp = (confidence - 1) * s_dot(N, yi, p);
          s_axpy(N, temp, yi, Ap);
        }
        float alpha = rsold / s_dot(N, p, Ap); /* :225 */
        s_axpy(N, alpha, p, x);                /* :228 */
        s_axpy(N, -alpha, Ap, r);              /* :231-232 */
        float rsnew = s_dot(N, r, r);          /* :234 */
        if (rsnew < 1e-20f) break;             /* :235-236 */
        float beta = rsnew / rsold;            /* :239-242 p = r + beta p */
        for (int i = 0; i < N; ++i) p[i] = beta * p[i];
        s_axpy(N, 1.0f, r, p);
        rsold = rsnew;                         /* :244 */
      }
    }
    free(p);
    free(r);
    free(Ap);
  }
}

/* ---- R6: training loss  (implicit/cpu/_als.pyx:259-308 _calculate_loss) ---------------------- */
ORACLE_API double oracle_calculate_loss(const float *YtY, const int32_t *indptr, const int32_t *indices,
                                        const float *data, const float *X, const float *Y, int64_t users,
                                        int64_t items, int N, float regularization, int num_threads) {
  double loss = 0, total_confidence = 0, item_norm = 0, user_norm = 0;
  (void)num_threads;
  float *r = (float *)malloc(sizeof(float) * N);
  for (int64_t u = 0; u < users; ++u) {
    const float *xu = X + u * N;
    s_symv(N, 1.0f, YtY, xu, r); /* :282 */
    for (int32_t idx = indptr[u]; idx < indptr[u + 1]; ++idx) { /* :284-298 */
      const float *yi = Y + (int64_t)indices[idx] * N;
      float confidence = data[idx], temp;
      if (confidence > 0) temp = -2 * confidence;
      else { temp = 0; confidence = -1 * confidence; }
      temp = temp + (confidence - 1) * s_dot(N, yi, xu);
      s_axpy(N, temp, yi, r);
      total_confidence += confidence;
      loss += confidence;
    }
    loss += s_dot(N, r, xu);       /* :300 */
    user_norm += s_dot(N, xu, xu); /* :301 */
  }
  for (int64_t i = 0; i < items; ++i) item_norm += s_dot(N, Y + i * N, Y + i * N); /* :303-304 */
  free(r);
  loss += regularization * (item_norm + user_norm); /* :307 */
  const int64_t nnz = indptr[users];
  return loss / (total_confidence + (double)users * (double)items - (double)nnz); /* :308 */
}

/* ---- R3: top-k select  (implicit/cpu/select.h:12-39) ----------------------------------------- */
typedef struct { float score; int32_t col; } pair_t;
/* std::greater<std::pair<T,int>> heap order == min-heap on (score, col) lexicographic */
static inline int pair_less(pair_t a, pair_t b) {
  return a.score < b.score || (!(b.score < a.score) && a.col < b.col);
}
static void heap_sift_down(pair_t *h, int n, int i) {
  for (;;) {
    int l = 2 * i + 1, r = l + 1, m = i;
    if (l < n && pair_less(h[l], h[m])) m = l;
    if (r < n && pair_less(h[r], h[m])) m = r;
    if (m == i) return;
    pair_t t = h[i]; h[i] = h[m]; h[m] = t;
    i = m;
  }
}
static void heap_sift_up(pair_t *h, int i) {
  while (i > 0) {
    int p = (i - 1) / 2;
    if (!pair_less(h[i], h[p])) return;
    pair_t t = h[i]; h[i] = h[p]; h[p] = t;
    i = p;
  }
}
/* ids/distances rows keep their incoming contents past results.size() (reference: zero-initialised
 * by topk.pyx:20-21). */
ORACLE_API void oracle_select(const float *batch, int rows, int cols, int k, int32_t *ids, float *distances) {
  pair_t *h = (pair_t *)malloc(sizeof(pair_t) * (k > 0 ? k : 1));
  for (int row = 0; row < rows; ++row) {
    int n = 0;
    for (int col = 0; col < cols; ++col) {
      float score = batch[(size_t)row * cols + col];
      if (n < k || score > h[0].score) { /* select.h:23: strict on score only */
        if (n >= k) { h[0] = h[n - 1]; --n; heap_sift_down(h, n, 0); } /* pop min pair */
        h[n].score = score; h[n].col = col; ++n;
        heap_sift_up(h, n - 1);
      }
    }
    /* sort_heap with greater<> => descending by (score, col): repeatedly extract the min to the back */
    for (int m = n; m > 1; --m) {
      pair_t t = h[0]; h[0] = h[m - 1]; h[m - 1] = t;
      heap_sift_down(h, m - 1, 0);
    }
    for (int i = 0; i < n; ++i) {
      ids[(size_t)row * k + i] = h[i].col;
      distances[(size_t)row * k + i] = h[i].score;
    }
  }
  free(h);
}

/* ---- R3: topk  (implicit/cpu/topk.pyx:15-67) ------------------------------------------------- */
/* scores = query . items^T (fp32), optional /item_norms (:48-49), liked columns (CSR over the query
 * rows, :51-53) and global filter_items (:55-56) set to -FLT_MAX, then select.  Outputs must be
 * zero-initialised by the caller (:20-21). */
ORACLE_API void oracle_topk(const float *items, int64_t n_items, const float *query, int64_t n_query, int f,
                            int k, const float *item_norms, const int32_t *filt_indptr,
                            const int32_t *filt_indices, const int32_t *filter_items, int64_t n_filter,
                            int32_t *ids, float *distances, int num_threads) {
  const float neginf = -3.402823466e+38f;
#ifdef _OPENMP
  if (num_threads <= 0) num_threads = omp_get_max_threads();
#pragma omp parallel num_threads(num_threads)
#endif
  {
    float *scores = (float *)malloc(sizeof(float) * (size_t)n_items);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
    for (int64_t q = 0; q < n_query; ++q) {
      const float *qv = query + q * f;
      for (int64_t i = 0; i < n_items; ++i) {
        float s = s_dot(f, qv, items + i * f);
        scores[i] = item_norms ? s / item_norms[i] : s;
      }
      if (filt_indptr)
        for (int32_t j = filt_indptr[q]; j < filt_indptr[q + 1]; ++j) scores[filt_indices[j]] = neginf;
      for (int64_t j = 0; j < n_filter; ++j) scores[filter_items[j]] = neginf;
      oracle_select(scores, 1, (int)n_items, k, ids + q * k, distances + q * k);
    }
    free(scores);
  }
}
