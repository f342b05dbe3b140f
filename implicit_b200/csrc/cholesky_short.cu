// R1, short rows: the same solve as cholesky.cu (reference: _least_squares, implicit/cpu/_als.pyx:76-142) for a
// row with n nonzeros, n well below the factor count F, through the n x n "push-through" system instead of the
// F x F normal equations.
//
// With G = Y^T Y + lambda I = R^T R (shared by every row of the half), P = R^-1 and the whitened factors W = Y P:
//     A_u = G + V^T D V,  b_u = V^T c+          V = the n gathered rows of Y,  D = diag(|c| - 1),  c+ = max(c, 0)
//     x_u = A_u^-1 b_u = G^-1 V^T (I + D K)^-1 c+,   K = V G^-1 V^T = W_u W_u^T
//         = P W_u^T E (I + E K E)^-1 E^-1 c+,        E = sqrt(D)
// so a row costs an n x n Gram matrix of whitened rows (mma.sync 3xTF32, as in cholesky.cu with the roles of
// "nonzero" and "factor" swapped), an n x n Cholesky of M = I + E K E (eigenvalues >= 1: always well conditioned),
// one pass r = W_u^T (E s) and the product x = P r with the triangular P held in shared memory.  Rows with n <= 16
// / 32 / 48 use a 16 / 32 / 48-wide system; cost drops from O(n F^2 + F^3) to O(n^2 F + n^3 + F^2).
// fp32 accuracy is on par with the F x F path (DESIGN.md section 4.1b has the comparison against an fp64 solve).
//
// Not every short item qualifies: a negative weight |c| - 1 < 0 (|c| < 1, or an explicit zero) makes M indefinite,
// chunks of giant rows are not rows, and a G that is not positive definite has no R.  Such items are appended to
// a deferred list that the full-size kernel of cholesky.cu processes right after; results never depend on which
// path took a row beyond fp32 rounding.
#include "cholesky_device.cuh"

namespace als {

namespace {

constexpr int kShortWarps = 8;

// ---- P = R^-1 in fp64 ---------------------------------------------------------------------------
// One CTA of 32 x 32 threads on the augmented matrix [G | I] (F x 2F doubles in shared memory).  Gaussian
// elimination without pivoting (G is SPD) turns it into [D U | L1^-1] with G = U^T D U, U unit upper triangular,
// L1 = U^T; then R = D^1/2 U and P = R^-1 = (D^-1/2 L1^-1)^T.  One barrier per pivot: step k only reads row k.
__global__ void __launch_bounds__(1024) whiten_factor_kernel(const float *__restrict__ Greg, int F, float *__restrict__ P,
                                                             int32_t *ok) {
  extern __shared__ __align__(16) unsigned char whiten_smem[];
  double *a = reinterpret_cast<double *>(whiten_smem);  // [F][2F + 1]
  const int ld = 2 * F + 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < F; i += 32)
    for (int j = tx; j < 2 * F; j += 32) a[i * ld + j] = j < F ? (double)Greg[i * F + j] : (j - F == i ? 1.0 : 0.0);
  __syncthreads();
  for (int k = 0; k < F; ++k) {
    const double d = a[k * ld + k];
    // same for every thread: G is not positive definite, not finite, or out of the range the fp32 seed covers
    if (!(d > 1e-30) || !(d < 1e30)) {
      if (threadIdx.x == 0) *ok = 0;
      return;
    }
    double rinv = (double)(1.f / (float)d);  // seed + two Newton steps: full double precision
    rinv = rinv * (2.0 - d * rinv);
    rinv = rinv * (2.0 - d * rinv);
    for (int i = k + 1 + ty; i < F; i += 32) {
      const double m = a[k * ld + i] * rinv;  // G is symmetric: the multiplier of row i is (D U)[k][i] / d_k
      for (int j = tx; j < 2 * F; j += 32)
        if ((j >= i && j < F) || (j >= F && j <= F + k)) a[i * ld + j] -= m * a[k * ld + j];
    }
    __syncthreads();  // row k + 1 (the next pivot row) is complete
  }
  for (int j = ty; j < F; j += 32) {  // row j of L1^-1 scaled by d_j^-1/2 is column j of P
    const double d = a[j * ld + j];
    double s = (double)rsqrtf((float)d);
    s = s * (1.5 - 0.5 * d * s * s);
    s = s * (1.5 - 0.5 * d * s * s);
    for (int i = tx; i < F; i += 32) P[i * F + j] = i <= j ? (float)(a[j * ld + F + i] * s) : 0.f;
  }
  if (threadIdx.x == 0) *ok = 1;
}

// ---- W = Y P ------------------------------------------------------------------------------------
// 128 rows per CTA pass; thread (ty, tx) owns rows 4 ty .. 4 ty + 3 and, in every 32-column half, columns
// 4 tx .. 4 tx + 3 (so both the P reads and the W writes of a warp are contiguous and conflict free).
template <int NB>
struct WhitenCfg {
  static constexpr int F = 16 * NB, LDY = F + 1, RT = 128;
  static constexpr int NH = (F + 31) / 32;  // 32-column halves
  static constexpr int SMEM_FLOATS = F * F + RT * LDY;
};

template <int NB>
__global__ void __launch_bounds__(256) whiten_rows_kernel(const float *__restrict__ Y, const float *__restrict__ P,
                                                          float *__restrict__ W, int64_t rows) {
  using C = WhitenCfg<NB>;
  constexpr int F = C::F, LDY = C::LDY, RT = C::RT, NH = C::NH;
  extern __shared__ __align__(16) unsigned char whiten_rows_smem[];
  float *Ps = reinterpret_cast<float *>(whiten_rows_smem);
  float *Ys = Ps + F * F;
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  for (int e = tid; e < F * F; e += 256) Ps[e] = P[e];
  for (int64_t r0 = (int64_t)blockIdx.x * RT; r0 < rows; r0 += (int64_t)gridDim.x * RT) {
    __syncthreads();
    for (int e = tid; e < RT * (F / 4); e += 256) {
      const int r = e / (F / 4), c4 = e % (F / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < rows) v = __ldg(reinterpret_cast<const float4 *>(Y + (r0 + r) * F) + c4);
      float *dst = Ys + r * LDY + 4 * c4;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    float acc[4][NH][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int h = 0; h < NH; ++h) acc[i][h][0] = acc[i][h][1] = acc[i][h][2] = acc[i][h][3] = 0.f;
    // P is upper triangular: row k only reaches the 32-column halves h >= k / 32
#pragma unroll
    for (int kb = 0; kb < NH; ++kb) {
#pragma unroll 4
      for (int k = 32 * kb; k < (32 * kb + 32 < F ? 32 * kb + 32 : F); ++k) {
        float y[4];
        float4 p[NH];
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = Ys[(4 * ty + i) * LDY + k];
#pragma unroll
        for (int h = kb; h < NH; ++h) {
          const int c = 32 * h + 4 * tx;
          p[h] = c < F ? *reinterpret_cast<const float4 *>(Ps + k * F + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int h = kb; h < NH; ++h) {
            acc[i][h][0] = fmaf(y[i], p[h].x, acc[i][h][0]);
            acc[i][h][1] = fmaf(y[i], p[h].y, acc[i][h][1]);
            acc[i][h][2] = fmaf(y[i], p[h].z, acc[i][h][2]);
            acc[i][h][3] = fmaf(y[i], p[h].w, acc[i][h][3]);
          }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + 4 * ty + i;
      if (r < rows) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int c = 32 * h + 4 * tx;
          if (c < F)
            *reinterpret_cast<float4 *>(W + r * F + c) = make_float4(acc[i][h][0], acc[i][h][1], acc[i][h][2], acc[i][h][3]);
        }
      }
    }
  }
}

// ---- W = Y P on the tensor cores ------------------------------------------------------------------
// Warp per 16-row tile: A = the tile of Y (cp.async, double buffered), B = P pre-split into TF32 hi / lo parts in
// shared memory, 3xTF32 mma.sync, triangular P: k-step s only reaches the 8-column tiles j >= s.  Memory-bound.
constexpr int kWhitenWarps = 8;

template <int NB>
struct WhitenMmaCfg {
  static constexpr int F = 16 * NB, NT8 = 2 * NB;
  static constexpr int LDY = F + 4;   // A-fragment reads conflict free
  static constexpr int LDP = F + 8;   // B-fragment reads conflict free
  static constexpr int TILE = 16 * LDY;
  static constexpr int SMEM_FLOATS = 2 * F * LDP + kWhitenWarps * 2 * TILE;
};

template <int NB>
__global__ void __launch_bounds__(32 * kWhitenWarps, 2)
whiten_rows_mma_kernel(const float *__restrict__ Y, const float *__restrict__ P, float *__restrict__ W, int64_t rows) {
  using C = WhitenMmaCfg<NB>;
  constexpr int F = C::F, NT8 = C::NT8, LDY = C::LDY, LDP = C::LDP;
  extern __shared__ __align__(16) unsigned char whiten_mma_smem[];
  uint32_t *Ph = reinterpret_cast<uint32_t *>(whiten_mma_smem);
  uint32_t *Pl = Ph + F * LDP;
  float *tiles = reinterpret_cast<float *>(Pl + F * LDP);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int e = threadIdx.x; e < F * F; e += blockDim.x) {
    uint32_t hi, lo;
    split_tf32(__ldg(P + e), hi, lo);
    Ph[(e / F) * LDP + e % F] = hi;
    Pl[(e / F) * LDP + e % F] = lo;
  }
  __syncthreads();
  float *mine = tiles + warp * 2 * C::TILE;
  const int64_t ntiles = (rows + 15) >> 4;
  const int64_t stride = (int64_t)gridDim.x * kWhitenWarps;
  auto issue = [&](int64_t tile, int buf) {
    if (tile < ntiles) {
      float *st = mine + buf * C::TILE;
      constexpr int CH = F / 4;
#pragma unroll
      for (int q = 0; q < 2 * NB; ++q) {  // 16 rows x F/4 chunks = 64 NB chunks
        const int id = q * 32 + lane, row = id / CH, ch = id % CH;
        int64_t r = tile * 16 + row;
        if (r >= rows) r = rows - 1;  // clamp: the duplicate rows are never stored
        cp_async16(st + row * LDY + ch * 4, Y + r * F + ch * 4);
      }
    }
    cp_async_commit();
  };
  int64_t tile = (int64_t)blockIdx.x * kWhitenWarps + warp;
  issue(tile, 0);
  int buf = 0;
  for (; tile < ntiles; tile += stride) {
    issue(tile + stride, buf ^ 1);
    cp_async_wait<1>();
    __syncwarp();
    const float *st = mine + buf * C::TILE;
    float acc[NT8][4];
#pragma unroll
    for (int j = 0; j < NT8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int s = 0; s < NT8; ++s) {  // k-step: factor dimensions 8s .. 8s+7
      uint32_t ah[4], al[4];
      split_tf32(st[g * LDY + 8 * s + t], ah[0], al[0]);
      split_tf32(st[(g + 8) * LDY + 8 * s + t], ah[1], al[1]);
      split_tf32(st[g * LDY + 8 * s + t + 4], ah[2], al[2]);
      split_tf32(st[(g + 8) * LDY + 8 * s + t + 4], ah[3], al[3]);
#pragma unroll
      for (int j = s; j < NT8; ++j) {
        const uint32_t bh0 = Ph[(8 * s + t) * LDP + 8 * j + g], bh1 = Ph[(8 * s + t + 4) * LDP + 8 * j + g];
        const uint32_t bl0 = Pl[(8 * s + t) * LDP + 8 * j + g], bl1 = Pl[(8 * s + t + 4) * LDP + 8 * j + g];
        mma_tf32(acc[j], al[0], al[1], al[2], al[3], bh0, bh1);
        mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], bl0, bl1);
        mma_tf32(acc[j], ah[0], ah[1], ah[2], ah[3], bh0, bh1);
      }
    }
    const int64_t r0 = tile * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
      if (r0 < rows) *reinterpret_cast<float2 *>(W + r0 * F + 8 * j + 2 * t) = make_float2(acc[j][0], acc[j][1]);
      if (r1 < rows) *reinterpret_cast<float2 *>(W + r1 * F + 8 * j + 2 * t) = make_float2(acc[j][2], acc[j][3]);
    }
    __syncwarp();  // everyone is done with this buffer before the next iteration refills it
    buf ^= 1;
  }
  cp_async_wait<0>();
}

template <int NB>
int run_whiten_rows(als_ctx *ctx, const als_factors *Y, cudaStream_t stream) {
  // Default: fp32 FMA (round-to-nearest accumulation).  The mma.sync version below is ~40 us faster per half on C2
  // but the tensor core truncates its accumulator on every add: measured 3x the error on short rows (7e-6 vs 2e-6
  // against an fp64 solve), so it stays behind ALS_B200_WHITEN_MMA.
  if (!getenv("ALS_B200_WHITEN_MMA")) {
    using C = WhitenCfg<NB>;
    const int smem = C::SMEM_FLOATS * (int)sizeof(float);
    auto kern = whiten_rows_kernel<NB>;
    ALS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(Y->rows, 1), C::RT), (int64_t)ctx->sm_count * 4);
    kern<<<grid, 256, smem, stream>>>(Y->d, ctx->Pinv, ctx->whitened, Y->rows);
  } else {
    using C = WhitenMmaCfg<NB>;
    const int smem = C::SMEM_FLOATS * (int)sizeof(float);
    auto kern = whiten_rows_mma_kernel<NB>;
    ALS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(Y->rows, 1), 16 * kWhitenWarps),
                                            (int64_t)ctx->sm_count * 2);
    kern<<<grid, 32 * kWhitenWarps, smem, stream>>>(Y->d, ctx->Pinv, ctx->whitened, Y->rows);
  }
  ALS_CUDA(cudaGetLastError());
  ctx->launches++;
  return ALS_OK;
}

// ---- the short-row solver -----------------------------------------------------------------------
template <int NB, int NBs>
struct ShortCfg {
  using S = Cfg<NBs>;                      // the n x n system, padded to NS
  static constexpr int F = 16 * NB;
  static constexpr int NS = 16 * NBs;
  static constexpr int NL = (NS + 31) / 32;  // nonzeros held per lane
  static constexpr int NG = NS / 8;          // 8-row groups of W_u
  static constexpr int KP = 16;              // factor dimensions staged per phase
  static constexpr int NPH = F / KP;
  static constexpr int LDW = KP + 4;         // conflict-free fragment reads, 16-byte aligned rows
  static constexpr int STAGE = NS * LDW;
  static constexpr int WORK = 2 * STAGE > S::U_FLOATS ? 2 * STAGE : S::U_FLOATS;  // stages, then U, then r
  static constexpr int WARP_FLOATS = WORK + NS;
  static constexpr int LDP = F + 4;
  static constexpr int SMEM_FLOATS = F * LDP + kShortWarps * WARP_FLOATS;
  static_assert(WORK >= F, "r does not fit");
};

// phase ph of the gather: 16 factor dimensions of every live 8-row group of W_u -> stage ph & 1
template <class C>
__device__ __forceinline__ void short_issue(float *wsm, const float *const (&src)[C::NG], int ph, int n, int g, int t) {
  float *st = wsm + (ph & 1) * C::STAGE;
#pragma unroll
  for (int q = 0; q < C::NG; ++q)
    if (8 * q < n) cp_async16(st + (8 * q + g) * C::LDW + 4 * t, src[q] + C::KP * ph);
  cp_async_commit();
}

template <int NB, int NBs>
__global__ void __launch_bounds__(32 * kShortWarps, NBs == 3 ? 2 : NBs == 2 ? 3 : 4)
short_rows_kernel(const int32_t *__restrict__ indices, const float *__restrict__ data, const float *__restrict__ W,
                  const float *__restrict__ P, float *__restrict__ X, int64_t row_offset,
                  const WorkItem *__restrict__ work, int n_work, int32_t *counter, WorkItem *deferred,
                  int32_t *n_deferred, const int32_t *whiten_ok, float *const *peers, int n_peers) {
  using C = ShortCfg<NB, NBs>;
  using S = typename C::S;
  constexpr int F = C::F, NS = C::NS, NL = C::NL, NG = C::NG, LDW = C::LDW, LDP = C::LDP;
  extern __shared__ __align__(16) unsigned char short_smem[];
  float *smem = reinterpret_cast<float *>(short_smem);
  float *Ps = smem;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float *wsm = smem + F * LDP + warp * C::WARP_FLOATS;
  float *zb = wsm + C::WORK;
  for (int e = threadIdx.x; e < F * F; e += blockDim.x) Ps[(e / F) * LDP + e % F] = __ldg(P + e);
  __syncthreads();
  const bool usable = *whiten_ok != 0;

  for (;;) {
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    v = __shfl_sync(0xffffffffu, v, 0);
    if (v >= n_work) break;
    const int4 raw = __ldg(reinterpret_cast<const int4 *>(work) + v);
    const WorkItem wi{raw.x, raw.y, raw.z, raw.w};
    const int n = wi.k1 - wi.k0;
    const int64_t xoff = (row_offset + wi.row) * F;
    auto defer = [&]() {
      if (lane == 0) deferred[atomicAdd(n_deferred, 1)] = wi;
    };
    if (wi.slot != -1 || !usable || n > NS) {
      defer();
      continue;
    }
    if (n == 0) {  // no observations: the reference zeroes the row (_als.pyx:98-100)
      for (int m = lane; m < F; m += 32) {
        X[xoff + m] = 0.f;
        for (int pi = 0; pi < n_peers; ++pi) peers[pi][xoff + m] = 0.f;
      }
      continue;
    }
    // ---- the row's nonzeros: nonzero i lives in slot i / 32 of lane i % 32
    int idx[NL];
    float ew[NL], rhs[NL];
    bool bad = false;
#pragma unroll
    for (int sl = 0; sl < NL; ++sl) {
      const int k = wi.k0 + 32 * sl + lane;
      const bool valid = k < wi.k1;
      idx[sl] = valid ? __ldg(indices + k) : -1;
      const float c = valid ? __ldg(data + k) : 1.f;
      const float w = fabsf(c) - 1.f;  // _als.pyx:115-118
      bad = bad || !(w >= 0.f);        // negative weight or NaN: not for this path
      ew[sl] = valid ? sqrtf(fmaxf(w, 1e-10f)) : 0.f;
      rhs[sl] = (valid && c > 0.f) ? c / ew[sl] : 0.f;  // E^-1 c+   (_als.pyx:119-121: only c > 0 feeds b)
    }
    if (__any_sync(0xffffffffu, bad)) {
      defer();
      continue;
    }
    const int first = __shfl_sync(0xffffffffu, idx[0], 0);
#pragma unroll
    for (int sl = 0; sl < NL; ++sl)
      if (idx[sl] < 0) idx[sl] = first;  // padding repeats a real row; its weight is 0
    // gather pointers: group q covers W_u rows 8q .. 8q+7, this lane copies chunk (lane & 3) of row 8q + (lane >> 2)
    const float *src[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) {
      const int ri = __shfl_sync(0xffffffffu, idx[(8 * q) >> 5], (8 * q + g) & 31);
      src[q] = W + (int64_t)ri * F + 4 * t;
    }
    RowState<NBs> st;
#pragma unroll
    for (int e = 0; e < S::NTILES; ++e) st.acc[e][0] = st.acc[e][1] = st.acc[e][2] = st.acc[e][3] = 0.f;

    // ---- K = W_u W_u^T, 16 factor dimensions per phase, double buffered
    short_issue<C>(wsm, src, 0, n, g, t);
#pragma unroll  // (cicc 12.9 crashes on this loop when it is kept rolled)
    for (int ph = 0; ph < C::NPH; ++ph) {
      if (ph + 1 < C::NPH) {
        short_issue<C>(wsm, src, ph + 1, n, g, t);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const float *sg = wsm + (ph & 1) * C::STAGE;
#pragma unroll
      for (int kk = 0; kk < C::KP; kk += 8) {
        uint32_t vh0[S::NT8], vl0[S::NT8], vh1[S::NT8], vl1[S::NT8];
#pragma unroll
        for (int c = 0; c < S::NT8; ++c) {
          split_tf32(sg[(8 * c + g) * LDW + kk + t], vh0[c], vl0[c]);
          split_tf32(sg[(8 * c + g) * LDW + kk + t + 4], vh1[c], vl1[c]);
        }
#pragma unroll
        for (int term = 0; term < 3; ++term) {
#pragma unroll
          for (int i = 0; i < NBs; ++i) {
            const uint32_t a0 = term == 0 ? vl0[2 * i] : vh0[2 * i];
            const uint32_t a1 = term == 0 ? vl0[2 * i + 1] : vh0[2 * i + 1];
            const uint32_t a2 = term == 0 ? vl1[2 * i] : vh1[2 * i];
            const uint32_t a3 = term == 0 ? vl1[2 * i + 1] : vh1[2 * i + 1];
#pragma unroll
            for (int j = 2 * i; j < S::NT8; ++j) {
              float(&d)[4] = st.acc[S::tidx(i, j)];
              if (term == 1) mma_tf32(d, a0, a1, a2, a3, vl0[j], vl1[j]);
              else mma_tf32(d, a0, a1, a2, a3, vh0[j], vh1[j]);
            }
          }
        }
      }
      __syncwarp();  // the stage is free for phase ph + 2
    }

    // ---- M = I + E K E on the n x n leading block, identity on the padding; rhs = E^-1 c+
#pragma unroll
    for (int i = 0; i < NBs; ++i) {
      const int r0 = 16 * i + g, r1 = r0 + 8;
      const float er0 = __shfl_sync(0xffffffffu, ew[(16 * i) >> 5], r0 & 31);
      const float er1 = __shfl_sync(0xffffffffu, ew[(16 * i) >> 5], r1 & 31);
#pragma unroll
      for (int j = 2 * i; j < S::NT8; ++j) {
        const int c0 = 8 * j + 2 * t, c1 = c0 + 1;
        const float ec0 = __shfl_sync(0xffffffffu, ew[(8 * j) >> 5], c0 & 31);
        const float ec1 = __shfl_sync(0xffffffffu, ew[(8 * j) >> 5], c1 & 31);
        float(&d)[4] = st.acc[S::tidx(i, j)];
        d[0] = (r0 < n && c0 < n ? er0 * ec0 * d[0] : 0.f) + (r0 == c0 ? 1.f : 0.f);
        d[1] = (r0 < n && c1 < n ? er0 * ec1 * d[1] : 0.f) + (r0 == c1 ? 1.f : 0.f);
        d[2] = (r1 < n && c0 < n ? er1 * ec0 * d[2] : 0.f) + (r1 == c0 ? 1.f : 0.f);
        d[3] = (r1 < n && c1 < n ? er1 * ec1 * d[3] : 0.f) + (r1 == c1 ? 1.f : 0.f);
      }
    }
#pragma unroll
    for (int c = 0; c < S::NT8; ++c) {
      const float bv = __shfl_sync(0xffffffffu, rhs[(8 * c) >> 5], (8 * c + g) & 31);
      st.bp[c] = t == 0 ? bv : 0.f;  // factor_solve sums bp over the 4 lanes of a group
    }

    // ---- M s = rhs
    bool ok = true;
    float s[NL];
    factor_solve<NBs>(st, wsm, zb, lane, ok, 0, s);
    if (!ok) {  // cannot happen for finite inputs (M >= I); let the full-size path decide
      defer();
      __syncwarp();
      continue;
    }

    // ---- r = W_u^T (E s): lane owns columns 2 lane, 2 lane + 1
    float r0 = 0.f, r1 = 0.f;
    const bool owns = 2 * lane < F;
#pragma unroll
    for (int sl = 0; sl < NL; ++sl) {
      const float tc = ew[sl] * s[sl];
      const int cnt = min(32, n - 32 * sl);
#pragma unroll 8
      for (int i = 0; i < cnt; ++i) {
        const float ti = __shfl_sync(0xffffffffu, tc, i);
        const int ri = __shfl_sync(0xffffffffu, idx[sl], i);
        if (owns) {
          const float2 w2 = __ldcg(reinterpret_cast<const float2 *>(W + (int64_t)ri * F) + lane);
          r0 = fmaf(ti, w2.x, r0);
          r1 = fmaf(ti, w2.y, r1);
        }
      }
    }
    __syncwarp();
    if (owns) *reinterpret_cast<float2 *>(wsm + 2 * lane) = make_float2(r0, r1);
    __syncwarp();
    // ---- x = P r, P upper triangular: row m = lane + 32 q needs columns k >= m (>= 32 q for the whole slot)
    constexpr int QF = (F + 31) / 32;
    float xx[QF];
#pragma unroll
    for (int q = 0; q < QF; ++q) {
      const int m = lane + 32 * q;
      const float *prow = Ps + (m < F ? m : 0) * LDP;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int k = 32 * q; k < F; k += 8) {
        const float4 p0 = *reinterpret_cast<const float4 *>(prow + k);
        const float4 p1 = *reinterpret_cast<const float4 *>(prow + k + 4);
        const float4 q0 = *reinterpret_cast<const float4 *>(wsm + k);
        const float4 q1 = *reinterpret_cast<const float4 *>(wsm + k + 4);
        a0 = fmaf(p0.x, q0.x, fmaf(p0.y, q0.y, fmaf(p0.z, q0.z, fmaf(p0.w, q0.w, a0))));
        a1 = fmaf(p1.x, q1.x, fmaf(p1.y, q1.y, fmaf(p1.z, q1.z, fmaf(p1.w, q1.w, a1))));
      }
      xx[q] = a0 + a1;
    }
    store_solution<F>(xx, X + xoff, lane, peers, n_peers, xoff);
    __syncwarp();  // r is dead; the next row may overwrite the work area
  }
}

template <int NB, int NBs>
int run_short(als_ctx *ctx, const als_csr *Cm, als_factors *X, int64_t begin, int64_t count, int slot, cudaStream_t stream) {
  using C = ShortCfg<NB, NBs>;
  if (count <= 0) return ALS_OK;
  const int smem = C::SMEM_FLOATS * (int)sizeof(float);
  auto kern = short_rows_kernel<NB, NBs>;
  ALS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int ctas_per_sm = 0;
  ALS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 32 * kShortWarps, smem));
  if (ctas_per_sm < 1) {
    set_error("short rows: kernel does not fit on an SM (smem %d bytes)", smem);
    return ALS_E_CUDA;
  }
  const int grid = (int)std::min<int64_t>(ceil_div(count, kShortWarps), (int64_t)ctx->sm_count * ctas_per_sm);
  kern<<<grid, 32 * kShortWarps, smem, stream>>>(Cm->indices, Cm->data, ctx->whitened, ctx->Pinv, X->d, Cm->row_offset,
                                                       Cm->work + begin, (int)count, ctx->counters + kCtrShort + slot,
                                                       ctx->deferred, ctx->counters + kCtrDeferredCount,
                                                       ctx->counters + kCtrWhitenOk, X->peers_dev, X->n_peers);
  ALS_CUDA(cudaGetLastError());
  ctx->launches++;
  return ALS_OK;
}

// items [begin, n_work) split into the classes (32, 48], (16, 32], [0, 16] by the schedule's suffix offsets
template <int NB>
int run_short_classes(als_ctx *ctx, const als_csr *Cm, als_factors *X, int64_t begin, int max_len, cudaStream_t stream) {
  const int64_t b48 = std::max(begin, Cm->le_begin[0]), b32 = std::max(begin, Cm->le_begin[1]),
                b16 = std::max(begin, Cm->le_begin[2]);
  int rc = ALS_OK;
  if (begin < b48) {
    set_error("short rows: items longer than 48 nonzeros");
    return ALS_E_INVALID;
  }
  if constexpr (NB >= 4) {
    if (max_len > 32) rc = run_short<NB, 3>(ctx, Cm, X, b48, b32 - b48, 0, stream);
    if (rc != ALS_OK) return rc;
  }
  if constexpr (NB >= 3) {
    if (max_len > 16) rc = run_short<NB, 2>(ctx, Cm, X, b32, b16 - b32, 1, stream);
    if (rc != ALS_OK) return rc;
  }
  return run_short<NB, 1>(ctx, Cm, X, b16, Cm->n_work - b16, 2, stream);
}

}  // namespace

int short_rows_prepare(als_ctx *ctx, const als_factors *Y, cudaStream_t stream) {
  const int F = Y->ld;
  if (F > 64) {
    set_error("short rows: factors beyond 64 are not supported");
    return ALS_E_UNSUPPORTED;
  }
  int rc = ensure_device_buffer(ctx, (void **)&ctx->whitened, &ctx->whitened_bytes,
                                std::max<int64_t>(Y->rows, 1) * F * (int64_t)sizeof(float));
  if (rc != ALS_OK) return rc;
  const int smem = F * (2 * F + 1) * (int)sizeof(double);
  ALS_CUDA(cudaFuncSetAttribute(whiten_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  whiten_factor_kernel<<<1, 1024, smem, stream>>>(ctx->Greg, F, ctx->Pinv, ctx->counters + kCtrWhitenOk);
  ALS_CUDA(cudaGetLastError());
  ctx->launches++;
  switch (F / 16) {
    case 2: return run_whiten_rows<2>(ctx, Y, stream);
    case 3: return run_whiten_rows<3>(ctx, Y, stream);
    case 4: return run_whiten_rows<4>(ctx, Y, stream);
    default:
      set_error("short rows: padded factors %d not supported", F);
      return ALS_E_UNSUPPORTED;
  }
}

int short_rows_launch(als_ctx *ctx, const als_csr *C, als_factors *X, const als_factors *Y, int64_t begin, int max_len,
                      cudaStream_t stream) {
  const int64_t count = C->n_work - begin;
  if (count <= 0) return ALS_OK;
  int rc = ensure_device_buffer(ctx, (void **)&ctx->deferred, &ctx->deferred_cap, count * (int64_t)sizeof(WorkItem));
  if (rc != ALS_OK) return rc;
  switch (Y->ld / 16) {
    case 2: return run_short_classes<2>(ctx, C, X, begin, max_len, stream);
    case 3: return run_short_classes<3>(ctx, C, X, begin, max_len, stream);
    case 4: return run_short_classes<4>(ctx, C, X, begin, max_len, stream);
    default:
      set_error("short rows: padded factors %d not supported", Y->ld);
      return ALS_E_UNSUPPORTED;
  }
}

}  // namespace als
