// R1: fused Cholesky half-iteration (reference: _least_squares, implicit/cpu/_als.pyx:76-142).
//
// One warp owns one row u of the CSR at a time (persistent warps pull rows, longest first, from an
// atomic work counter):
//   gather   the factor rows Y[i] of the row's nonzeros are staged 8 at a time into shared memory
//            with 16-byte cp.async copies (3-deep ring per warp, prefetched across row boundaries);
//   A, b     A_u = (Y^T Y + lambda I) + sum_k (|c_k| - 1) y_k y_k^T is accumulated in REGISTERS as the
//            upper-triangular set of 16x8 mma.sync.m16n8k8 TF32 tiles, with the 3xTF32 split
//            (hi*hi + hi*lo + lo*hi) so the result is fp32-faithful (plain TF32 would miss the 1e-4
//            parity bar); b_u = sum_{c_k > 0} c_k y_k rides along in fp32 FMAs;
//   solve    a right-looking blocked Cholesky with 8-row panels: each panel is spilled to shared
//            memory, factored (8x8 diagonal block redundantly per lane, panel columns one per lane),
//            and the trailing matrix is updated IN REGISTERS by the same 3xTF32 mma tiles; the
//            forward substitution rides along as one more column, the back substitution runs on the
//            packed U left in shared memory.
// Giant rows are split into chunks whose partial (A, b) go to global scratch and are summed in a
// fixed order by a second "finish" launch, so results do not depend on scheduling.
#include <limits.h>

#include "common.h"

namespace als {

namespace {

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gmem_src) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

constexpr uint32_t kTf32Mask = 0xffffe000u;  // keep sign, exponent and the 10 TF32 mantissa bits
constexpr uint32_t kSignBit = 0x80000000u;

// hi = x rounded to nearest TF32 (add half an ulp of the 10-bit mantissa, then clear the low 13 bits),
// lo = (x - hi) -- exact in fp32 -- rounded the same way.  Rounding (instead of letting the tensor core
// truncate) halves the error of each term and, more importantly, removes its bias: on all-positive
// data (the first ALS half-iteration) truncation errors add up linearly instead of as a random walk.
__device__ __forceinline__ uint32_t rn_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & kTf32Mask; }
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  hi = rn_tf32(x);
  lo = rn_tf32(x - __uint_as_float(hi));
}

template <int NB>
struct Cfg {
  static constexpr int F = 16 * NB;         // padded factors
  static constexpr int NT8 = 2 * NB;        // 8-wide column tiles == 8-row panels
  static constexpr int NTILES = NB * (NB + 1);
  static constexpr int LDS = F + 8;         // staged-row stride: conflict-free fragment reads
  static constexpr int NSTAGE = 3;
  static constexpr int STAGE_FLOATS = 8 * LDS + 16;  // 8 rows + w[8] + cpos[8]
  // packed U: panel p holds rows 8p..8p+7, columns 8p..F-1; stride == 8 or 24 (mod 32)
  __host__ __device__ static constexpr int pstride(int p) { return F - 8 * p + ((p & 1) ? 0 : 8); }
  __host__ __device__ static constexpr int poff(int p) {
    int o = 0;
    for (int q = 0; q < p; ++q) o += 8 * pstride(q);
    return o;
  }
  static constexpr int U_FLOATS = poff(NT8);
  static constexpr int WARP_FLOATS = NSTAGE * STAGE_FLOATS + U_FLOATS + F /* z */;
  // index of tile (i, j), j >= 2i, in the upper-triangular tile list
  __host__ __device__ static constexpr int tidx(int i, int j) { return i * NT8 - i * (i - 1) + (j - 2 * i); }
  static constexpr int SLOT_FLOATS = 32 * (NTILES * 4 + NT8);
};

constexpr int kWarpsPerCta = 4;

template <int NB>
struct RowState {
  float acc[Cfg<NB>::NTILES][4];
  float bp[Cfg<NB>::NT8];  // b partials: b[8c + g] = sum over the 4 lanes of group g of bp[c]
};

// ---- gather ------------------------------------------------------------------------------------
template <int NB>
__device__ __forceinline__ void issue_kstep(float *stage, const WorkItem &wi, int ks, const int32_t *__restrict__ indices,
                                            const float *__restrict__ data, const float *__restrict__ Y, int lane) {
  using C = Cfg<NB>;
  const int kbase = wi.k0 + 8 * ks;
  if (kbase < wi.k1) {
    const int k = kbase + (lane & 7);
    const bool valid = k < wi.k1;
    const int idx = __ldg(indices + (valid ? k : wi.k0));
    const float c = valid ? __ldg(data + k) : 0.f;
    if (lane < 8) {
      // confidence > 0: b += c y, A += (c - 1) y y^T;  else: A += (-c - 1) y y^T   (_als.pyx:115-124)
      stage[8 * C::LDS + lane] = valid ? (fabsf(c) - 1.f) : 0.f;
      stage[8 * C::LDS + 8 + lane] = c > 0.f ? c : 0.f;
    }
    constexpr int CH = C::F / 4;  // 16-byte chunks per factor row
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int id = q * 32 + lane;
      const int row = id / CH, ch = id % CH;
      const int ridx = __shfl_sync(0xffffffffu, idx, row);
      cp_async16(stage + row * C::LDS + ch * 4, Y + (int64_t)ridx * C::F + ch * 4);
    }
  }
  cp_async_commit();
}

// ---- accumulate one k-step (8 nonzeros) --------------------------------------------------------
template <int NB>
__device__ __forceinline__ void consume_kstep(RowState<NB> &st, const float *stage, int g, int t) {
  using C = Cfg<NB>;
  const float w0 = stage[8 * C::LDS + t], w1 = stage[8 * C::LDS + t + 4];
  const float c0 = stage[8 * C::LDS + 8 + t], c1 = stage[8 * C::LDS + 8 + t + 4];
  uint32_t yh0[C::NT8], yl0[C::NT8], yh1[C::NT8], yl1[C::NT8];
  float y0[C::NT8], y1[C::NT8];
#pragma unroll
  for (int c = 0; c < C::NT8; ++c) {
    y0[c] = stage[t * C::LDS + 8 * c + g];
    y1[c] = stage[(t + 4) * C::LDS + 8 * c + g];
  }
#pragma unroll
  for (int c = 0; c < C::NT8; ++c) {
    st.bp[c] = fmaf(c0, y0[c], st.bp[c]);
    st.bp[c] = fmaf(c1, y1[c], st.bp[c]);
    split_tf32(y0[c], yh0[c], yl0[c]);
    split_tf32(y1[c], yh1[c], yl1[c]);
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    // A fragment rows: z = w * y for factor indices 16i + g (a0,a2) and 16i + 8 + g (a1,a3)
    uint32_t ah[4], al[4];
    split_tf32(w0 * y0[2 * i], ah[0], al[0]);
    split_tf32(w0 * y0[2 * i + 1], ah[1], al[1]);
    split_tf32(w1 * y1[2 * i], ah[2], al[2]);
    split_tf32(w1 * y1[2 * i + 1], ah[3], al[3]);
#pragma unroll
    for (int j = 2 * i; j < C::NT8; ++j) {
      float(&d)[4] = st.acc[C::tidx(i, j)];
      mma_tf32(d, al[0], al[1], al[2], al[3], yh0[j], yh1[j]);
      mma_tf32(d, ah[0], ah[1], ah[2], ah[3], yl0[j], yl1[j]);
      mma_tf32(d, ah[0], ah[1], ah[2], ah[3], yh0[j], yh1[j]);
    }
  }
}

// ---- blocked Cholesky + solves -----------------------------------------------------------------
// Returns false when a pivot is not positive (LAPACK posv info != 0, _als.pyx:131-138).
template <int NB>
__device__ __forceinline__ bool factor_solve(RowState<NB> &st, float *U, float *zb, float *__restrict__ xout, int lane) {
  using C = Cfg<NB>;
  constexpr int F = C::F;
  const int g = lane >> 2, t = lane & 3;
  bool ok = true;

#pragma unroll
  for (int p = 0; p < C::NT8; ++p) {
    const int i = p >> 1, h = p & 1;
    float *Up = U + C::poff(p);
    constexpr int dummy = 0;
    (void)dummy;
    const int sp = C::pstride(p);
    const int Wp = F - 8 * p;
    // 1. spill panel rows 8p..8p+7 (columns 8p..F-1) and the matching slice of b
#pragma unroll
    for (int j = p; j < C::NT8; ++j) {
      const float2 v = make_float2(st.acc[C::tidx(i, j)][2 * h], st.acc[C::tidx(i, j)][2 * h + 1]);
      *reinterpret_cast<float2 *>(Up + g * sp + 8 * (j - p) + 2 * t) = v;
    }
    {
      float bq = st.bp[p];
      bq += __shfl_xor_sync(0xffffffffu, bq, 1);
      bq += __shfl_xor_sync(0xffffffffu, bq, 2);
      if (t == 0) zb[8 * p + g] = bq;
    }
    __syncwarp();
    // 2. 8x8 diagonal block, factored redundantly by every lane: U_d^T U_d = D, inv[r] = 1 / U_d[r][r]
    float D[8][8], inv[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float4 lo = *reinterpret_cast<const float4 *>(Up + r * sp);
      const float4 hi = *reinterpret_cast<const float4 *>(Up + r * sp + 4);
      D[r][0] = lo.x; D[r][1] = lo.y; D[r][2] = lo.z; D[r][3] = lo.w;
      D[r][4] = hi.x; D[r][5] = hi.y; D[r][6] = hi.z; D[r][7] = hi.w;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float d = D[r][r];
#pragma unroll
      for (int q = 0; q < r; ++q) d = fmaf(-D[q][r], D[q][r], d);
      ok = ok && (d > 0.f);
      float s = rsqrtf(d);
      s = s * fmaf(-0.5f * d * s, s, 1.5f);  // one Newton step: full fp32 accuracy
      inv[r] = s;
#pragma unroll
      for (int c = r + 1; c < 8; ++c) {
        float v = D[r][c];
#pragma unroll
        for (int q = 0; q < r; ++q) v = fmaf(-D[q][r], D[q][c], v);
        D[r][c] = v * s;
      }
    }
    // 3. panel columns, one per lane: v <- U_d^-T v (local column Wp is the rhs slice in zb)
    for (int c = lane; c <= Wp; c += 32) {
      float *colp = (c < Wp) ? (Up + c) : (zb + 8 * p);
      const int rs = (c < Wp) ? sp : 1;
      float v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) v[r] = colp[r * rs];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float a = v[r];
#pragma unroll
        for (int q = 0; q < r; ++q) a = fmaf(-D[q][r], v[q], a);
        v[r] = (c == r) ? inv[r] : a * inv[r];  // the diagonal stores the reciprocal pivot
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) colp[r * rs] = v[r];
    }
    __syncwarp();
    if (!ok) return false;
    // 4. trailing update in registers: A[m][n] -= sum_r U[r][m] U[r][n]; b[m] -= sum_r U[r][m] z[r]
    if (p + 1 < C::NT8) {
      const float z0 = zb[8 * p + t], z1 = zb[8 * p + t + 4];
      uint32_t uh0[C::NT8], ul0[C::NT8], uh1[C::NT8], ul1[C::NT8];
#pragma unroll
      for (int j = p + 1; j < C::NT8; ++j) {
        const float u0 = Up[t * sp + 8 * (j - p) + g];
        const float u1 = Up[(t + 4) * sp + 8 * (j - p) + g];
        st.bp[j] = fmaf(-u0, z0, st.bp[j]);
        st.bp[j] = fmaf(-u1, z1, st.bp[j]);
        split_tf32(u0, uh0[j], ul0[j]);
        split_tf32(u1, uh1[j], ul1[j]);
      }
#pragma unroll
      for (int ib = (p + 1) >> 1; ib < NB; ++ib) {
        // rows 16 ib + g (a0, a2) are still live only if 2 ib > p
        const bool top = (2 * ib > p);
        const uint32_t ah0 = top ? (uh0[top ? 2 * ib : p + 1] ^ kSignBit) : 0u;
        const uint32_t al0 = top ? (ul0[top ? 2 * ib : p + 1] ^ kSignBit) : 0u;
        const uint32_t ah2 = top ? (uh1[top ? 2 * ib : p + 1] ^ kSignBit) : 0u;
        const uint32_t al2 = top ? (ul1[top ? 2 * ib : p + 1] ^ kSignBit) : 0u;
        const uint32_t ah1 = uh0[2 * ib + 1] ^ kSignBit, al1 = ul0[2 * ib + 1] ^ kSignBit;
        const uint32_t ah3 = uh1[2 * ib + 1] ^ kSignBit, al3 = ul1[2 * ib + 1] ^ kSignBit;
#pragma unroll
        for (int j = (2 * ib > p + 1 ? 2 * ib : p + 1); j < C::NT8; ++j) {
          float(&d)[4] = st.acc[C::tidx(ib, j)];
          mma_tf32(d, al0, al1, al2, al3, uh0[j], uh1[j]);
          mma_tf32(d, ah0, ah1, ah2, ah3, ul0[j], ul1[j]);
          mma_tf32(d, ah0, ah1, ah2, ah3, uh0[j], uh1[j]);
        }
      }
    }
  }

  // 5. back substitution U x = z on the packed panels (column oriented: no reductions)
  constexpr int Q = (F + 31) / 32;
  float zz[Q], xx[Q];
  int rowoff[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int m = lane + 32 * q;
    const int mm = m < F ? m : 0;
    const int pm = mm >> 3;
    // poff(pm) in closed form: 8 * sum_{s<pm} (F - 8 s + 8 [s even])
    const int po = 8 * (pm * F - 4 * pm * (pm - 1) + 8 * ((pm + 1) >> 1));
    const int ps = F - 8 * pm + ((pm & 1) ? 0 : 8);
    rowoff[q] = po + (mm & 7) * ps - 8 * pm;
    zz[q] = zb[mm];
    xx[q] = 0.f;
  }
#pragma unroll
  for (int r = F - 1; r >= 0; --r) {
    constexpr int unused = 0;
    (void)unused;
    const int pr = r >> 3;
    const float invr = U[C::poff(pr) + (r & 7) * C::pstride(pr) + (r - 8 * pr)];
    const float xr = __shfl_sync(0xffffffffu, zz[r >> 5], r & 31) * invr;
    if (lane == (r & 31)) xx[r >> 5] = xr;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int m = lane + 32 * q;
      if (m < r && 32 * q < r) zz[q] = fmaf(-U[rowoff[q] + r], xr, zz[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int m = lane + 32 * q;
    if (m < F) xout[m] = xx[q];
  }
  return true;
}

// ---- kernel ------------------------------------------------------------------------------------
// pass 0: whole rows and chunks of giant rows;  pass 1: finish giant rows from their chunk slots.
template <int NB>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 3)
cholesky_half_kernel(const int32_t *__restrict__ indices, const float *__restrict__ data, const float *__restrict__ Y,
                     float *__restrict__ X, int64_t row_offset, const float *__restrict__ Greg,
                     const WorkItem *__restrict__ work, int n_work, int32_t *counter, float *slots,
                     long long *bad_row, int pass) {
  using C = Cfg<NB>;
  constexpr int F = C::F;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float *wsm = smem + warp * C::WARP_FLOATS;
  float *stages = wsm;
  float *U = wsm + C::NSTAGE * C::STAGE_FLOATS;
  float *zb = U + C::U_FLOATS;

  auto fetch = [&]() -> int {
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    v = __shfl_sync(0xffffffffu, v, 0);
    return v < n_work ? v : -1;
  };
  auto load_item = [&](int i) -> WorkItem {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(work) + i);
    return WorkItem{v.x, v.y, v.z, v.w};
  };

  int cur = fetch();
  WorkItem wi{0, 0, 0, -1};
  int stage = 0;
  if (cur >= 0) {
    wi = load_item(cur);
    if (pass == 0) {
      issue_kstep<NB>(stages + 0 * C::STAGE_FLOATS, wi, 0, indices, data, Y, lane);
      issue_kstep<NB>(stages + 1 * C::STAGE_FLOATS, wi, 1, indices, data, Y, lane);
    }
  }

  RowState<NB> st;
  while (cur >= 0) {
    const bool whole = (wi.slot == -1), chunk = (wi.slot >= 0), finish = (wi.slot == -2);
    // ---- initialise the accumulators: Y^T Y + lambda I for a row that will be solved, 0 for a chunk
#pragma unroll
    for (int c = 0; c < C::NT8; ++c) st.bp[c] = 0.f;
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
      for (int j = 2 * i; j < C::NT8; ++j) {
        float(&d)[4] = st.acc[C::tidx(i, j)];
        if (chunk) {
          d[0] = d[1] = d[2] = d[3] = 0.f;
        } else {
          const float2 top = __ldg(reinterpret_cast<const float2 *>(Greg + (16 * i + g) * F + 8 * j + 2 * t));
          const float2 bot = __ldg(reinterpret_cast<const float2 *>(Greg + (16 * i + g + 8) * F + 8 * j + 2 * t));
          d[0] = top.x; d[1] = top.y; d[2] = bot.x; d[3] = bot.y;
        }
      }

    int nxt = -1;
    WorkItem wn{0, 0, 0, -1};
    if (pass == 0) {
      const int nks = (wi.k1 - wi.k0 + 7) >> 3;
      for (int ks = 0; ks < nks; ++ks) {
        cp_async_wait<1>();
        __syncwarp();
        int s2 = stage + 2;
        if (s2 >= C::NSTAGE) s2 -= C::NSTAGE;
        issue_kstep<NB>(stages + s2 * C::STAGE_FLOATS, wi, ks + 2, indices, data, Y, lane);
        consume_kstep<NB>(st, stages + stage * C::STAGE_FLOATS, g, t);
        if (++stage == C::NSTAGE) stage = 0;
      }
      // prefetch the first two k-steps of the next item; they land while this row is factored
      nxt = fetch();
      __syncwarp();
      if (nxt >= 0) {
        wn = load_item(nxt);
        int s1 = stage + 1;
        if (s1 >= C::NSTAGE) s1 -= C::NSTAGE;
        issue_kstep<NB>(stages + stage * C::STAGE_FLOATS, wn, 0, indices, data, Y, lane);
        issue_kstep<NB>(stages + s1 * C::STAGE_FLOATS, wn, 1, indices, data, Y, lane);
      }
    } else {
      // finish: add the chunk partials in slot order
      for (int s = 0; s < wi.k1; ++s) {
        const float *sl = slots + (int64_t)(wi.k0 + s) * C::SLOT_FLOATS;
#pragma unroll
        for (int e = 0; e < C::NTILES; ++e)
#pragma unroll
          for (int v = 0; v < 4; ++v) st.acc[e][v] += sl[(e * 4 + v) * 32 + lane];
#pragma unroll
        for (int c = 0; c < C::NT8; ++c) st.bp[c] += sl[(C::NTILES * 4 + c) * 32 + lane];
      }
      nxt = fetch();
      if (nxt >= 0) wn = load_item(nxt);
    }

    if (chunk) {
      float *sl = slots + (int64_t)wi.slot * C::SLOT_FLOATS;
#pragma unroll
      for (int e = 0; e < C::NTILES; ++e)
#pragma unroll
        for (int v = 0; v < 4; ++v) sl[(e * 4 + v) * 32 + lane] = st.acc[e][v];
#pragma unroll
      for (int c = 0; c < C::NT8; ++c) sl[(C::NTILES * 4 + c) * 32 + lane] = st.bp[c];
    } else {
      float *xout = X + (row_offset + wi.row) * F;
      if (whole && wi.k0 == wi.k1) {
        // no observations: the reference zeroes the row (_als.pyx:98-100)
        for (int m = lane; m < F; m += 32) xout[m] = 0.f;
      } else if (whole || finish) {
        const bool ok = factor_solve<NB>(st, U, zb, xout, lane);
        if (!ok && lane == 0) atomicMin(bad_row, (long long)(row_offset + wi.row));
        __syncwarp();
      }
    }
    cur = nxt;
    wi = wn;
  }
  cp_async_wait<0>();
}

__global__ void init_solver_scalars(int32_t *counters, long long *bad_row) {
  if (threadIdx.x < 16) counters[threadIdx.x] = 0;
  if (threadIdx.x == 0) bad_row[0] = LLONG_MAX;
}

template <int NB>
int run_cholesky(als_ctx *ctx, const als_csr *Cm, als_factors *X, const als_factors *Y) {
  using C = Cfg<NB>;
  const int smem = C::WARP_FLOATS * kWarpsPerCta * (int)sizeof(float);
  auto kern = cholesky_half_kernel<NB>;
  ALS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int ctas_per_sm = 0;
  ALS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 32 * kWarpsPerCta, smem));
  if (ctas_per_sm < 1) {
    set_error("cholesky: kernel does not fit on an SM (smem %d bytes)", smem);
    return ALS_E_CUDA;
  }
  float *slots = nullptr;
  if (Cm->n_slots) {
    int rc = ensure_scratch(ctx, (int64_t)Cm->n_slots * C::SLOT_FLOATS * (int64_t)sizeof(float));
    if (rc != ALS_OK) return rc;
    slots = (float *)ctx->scratch;
  }
  init_solver_scalars<<<1, 32, 0, ctx->stream>>>(ctx->counters, ctx->bad_row);
  ALS_CUDA(cudaGetLastError());
  ctx->launches++;
  if (Cm->n_work) {
    const int64_t want = ceil_div(Cm->n_work, kWarpsPerCta);
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * ctas_per_sm);
    ProfScope prof(ctx, kProfCholesky);
    kern<<<grid, 32 * kWarpsPerCta, smem, ctx->stream>>>(Cm->indices, Cm->data, Y->d, X->d, Cm->row_offset, ctx->Greg,
                                                          Cm->work, (int)Cm->n_work, ctx->counters, slots,
                                                          ctx->bad_row, 0);
    ALS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (Cm->n_finish) {
    const int64_t want = ceil_div(Cm->n_finish, kWarpsPerCta);
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * ctas_per_sm);
    ProfScope prof(ctx, kProfCholFinish);
    kern<<<grid, 32 * kWarpsPerCta, smem, ctx->stream>>>(Cm->indices, Cm->data, Y->d, X->d, Cm->row_offset, ctx->Greg,
                                                          Cm->finish, (int)Cm->n_finish, ctx->counters + 1, slots,
                                                          ctx->bad_row, 1);
    ALS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  return ALS_OK;
}

}  // namespace

int launch_cholesky(als_ctx *ctx, const als_csr *C, als_factors *X, const als_factors *Y) {
  if (X->ld != Y->ld) {
    set_error("cholesky: X and Y strides differ (%d vs %d)", X->ld, Y->ld);
    return ALS_E_INVALID;
  }
  switch (Y->ld / 16) {
    case 1: return run_cholesky<1>(ctx, C, X, Y);
    case 2: return run_cholesky<2>(ctx, C, X, Y);
    case 3: return run_cholesky<3>(ctx, C, X, Y);
    case 4: return run_cholesky<4>(ctx, C, X, Y);
    default:
      return launch_cholesky_wide(ctx, C, X, Y);  // 64 < padded factors <= 128: cholesky_wide.cu
  }
}

}  // namespace als
