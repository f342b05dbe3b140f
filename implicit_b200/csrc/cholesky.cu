// R1: fused Cholesky half-iteration (reference: _least_squares, implicit/cpu/_als.pyx:76-142).
//
// One warp owns one row u of the CSR at a time (persistent warps pull rows, longest first, from an
// atomic work counter):
//   gather   the factor rows Y[i] of the row's nonzeros are staged 8 at a time into shared memory
//            with 16-byte cp.async copies (3-deep ring per warp, prefetched across row boundaries;
//            indices / confidences are prefetched 32 at a time into registers one block ahead);
//   A, b     A_u = (Y^T Y + lambda I) + sum_k (|c_k| - 1) y_k y_k^T is accumulated in REGISTERS as the
//            upper-triangular set of 16x8 mma.sync.m16n8k8 TF32 tiles, with the 3xTF32 split
//            (hi*hi + hi*lo + lo*hi) so the result is fp32-faithful (plain TF32 would miss the 1e-4
//            parity bar); b_u = sum_{c_k > 0} c_k y_k rides along in fp32 FMAs;
//   solve    a right-looking blocked Cholesky with 8-row panels: each panel is spilled to shared
//            memory, one lane owns one panel column (the rhs slice and 8 unit vectors ride along as
//            extra columns), the 8 pivots are eliminated LDL^T-style with warp shuffles and the rows
//            scaled by 1/sqrt(d) afterwards; the trailing matrix is updated IN REGISTERS by the same
//            3xTF32 mma tiles; the back substitution resolves a panel at a time with the inverse of
//            its diagonal block (the forward-substituted unit vectors) on the packed U in shared memory.
// Giant rows are split into chunks whose partial (A, b) go to global scratch and are summed in a
// fixed order by a second "finish" launch, so results do not depend on scheduling.
#include <limits.h>
#include <stdlib.h>

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
// lo is handed over raw: the tensor core drops its low 13 bits, an error of at most 2^-21 |x| that is
// unbiased because, with hi rounded to nearest, lo is symmetric around zero.  (Rounding lo as well
// cost two more integer ops per value -- 17% of the kernel's instructions -- for no measurable gain.)
__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo) {
  hi = rn_tf32(x);
  lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int NB>
struct Cfg {
  static constexpr int F = 16 * NB;         // padded factors
  static constexpr int NT8 = 2 * NB;        // 8-wide column tiles == 8-row panels
  static constexpr int NTILES = NB * (NB + 1);
  static constexpr int LDS = F + 8;         // staged-row stride: conflict-free fragment reads
  static constexpr int NSTAGE = 3;
  static constexpr int STAGE_FLOATS = 8 * LDS + 16;  // 8 rows + sw[8] + cpos[8]
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

// 32 consecutive nonzeros of a row, one per lane, prefetched into registers well before the k-steps
// that gather them (the index load would otherwise sit on the critical path of every k-step).
struct Blk {
  int idx;   // column index, -1 past the end of the row
  float c;   // raw confidence; decoded only when the k-step is issued, so the load stays in flight
};

__device__ __forceinline__ Blk load_block(const WorkItem &wi, int b, const int32_t *__restrict__ indices,
                                          const float *__restrict__ data, int lane) {
  const int k = wi.k0 + 32 * b + lane;
  const bool valid = k < wi.k1;
  Blk r;
  r.idx = valid ? __ldg(indices + k) : -1;
  r.c = valid ? __ldg(data + k) : 0.f;
  return r;
}

// ---- gather: k-step s (0..3) of block `blk` -> stage -----------------------------------------------
template <int NB>
__device__ __forceinline__ void issue_kstep(float *stage, const Blk &blk, int s, bool active,
                                            const float *__restrict__ Y, int lane) {
  using C = Cfg<NB>;
  if (active) {  // warp uniform
    const int src = 8 * s + (lane & 7);
    const float c = __shfl_sync(0xffffffffu, blk.c, src);
    const int myidx = __shfl_sync(0xffffffffu, blk.idx, src);
    if (lane < 8) {
      // A += w y y^T with w = |c| - 1 = sign(w) (sqrt|w| y)(sqrt|w| y)^T; b += c y for c > 0   (_als.pyx:115-124)
      const float w = (myidx >= 0) ? fabsf(c) - 1.f : 0.f;
      stage[8 * C::LDS + lane] = copysignf(__fsqrt_rn(fabsf(w)), w);
      stage[8 * C::LDS + 8 + lane] = (myidx >= 0 && c > 0.f) ? c : 0.f;
    }
    const int first = __shfl_sync(0xffffffffu, blk.idx, 8 * s);  // the first row of an active k-step exists
    constexpr int CH = C::F / 4;  // 16-byte chunks per factor row
#pragma unroll
    for (int q = 0; q < NB; ++q) {
      const int id = q * 32 + lane;
      const int row = id / CH, ch = id % CH;
      int ridx = __shfl_sync(0xffffffffu, blk.idx, 8 * s + row);
      if (ridx < 0) ridx = first;  // padding rows carry sw = cp = 0
      cp_async16(stage + row * C::LDS + ch * 4, Y + (int64_t)ridx * C::F + ch * 4);
    }
  }
  cp_async_commit();
}

// ---- accumulate one k-step (8 nonzeros) --------------------------------------------------------
template <int NB>
__device__ __forceinline__ void consume_kstep(RowState<NB> &st, const float *stage, int g, int t) {
  using C = Cfg<NB>;
  const float s0 = stage[8 * C::LDS + t], s1 = stage[8 * C::LDS + t + 4];
  const float c0 = stage[8 * C::LDS + 8 + t], c1 = stage[8 * C::LDS + 8 + t + 4];
  const float a0 = fabsf(s0), a1 = fabsf(s1);
  const uint32_t m0 = __float_as_uint(s0) & kSignBit, m1 = __float_as_uint(s1) & kSignBit;
  uint32_t vh0[C::NT8], vl0[C::NT8], vh1[C::NT8], vl1[C::NT8];
#pragma unroll
  for (int c = 0; c < C::NT8; ++c) {
    const float y0 = stage[t * C::LDS + 8 * c + g];
    const float y1 = stage[(t + 4) * C::LDS + 8 * c + g];
    st.bp[c] = fmaf(c0, y0, st.bp[c]);
    st.bp[c] = fmaf(c1, y1, st.bp[c]);
    split_tf32(a0 * y0, vh0[c], vl0[c]);  // v = sqrt|w| y: one split serves both mma operands
    split_tf32(a1 * y1, vh1[c], vl1[c]);
  }
  // Term-major order: the three 3xTF32 terms of one tile chain through its accumulator, so they are
  // issued a full sweep of tiles apart instead of back to back (an HMMA result takes ~35 cycles).
#pragma unroll
  for (int term = 0; term < 3; ++term) {
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      // A fragment: rows 16i + g (a0, a2) and 16i + 8 + g (a1, a3) of sign(w) v; lo part for term 0
      const uint32_t a0 = (term == 0 ? vl0[2 * i] : vh0[2 * i]) ^ m0;
      const uint32_t a1 = (term == 0 ? vl0[2 * i + 1] : vh0[2 * i + 1]) ^ m0;
      const uint32_t a2 = (term == 0 ? vl1[2 * i] : vh1[2 * i]) ^ m1;
      const uint32_t a3 = (term == 0 ? vl1[2 * i + 1] : vh1[2 * i + 1]) ^ m1;
#pragma unroll
      for (int j = 2 * i; j < C::NT8; ++j) {
        float(&d)[4] = st.acc[C::tidx(i, j)];
        if (term == 1) mma_tf32(d, a0, a1, a2, a3, vl0[j], vl1[j]);  // hi * lo
        else mma_tf32(d, a0, a1, a2, a3, vh0[j], vh1[j]);            // lo * hi, then hi * hi
      }
    }
  }
}

// ---- blocked Cholesky + solves -----------------------------------------------------------------
// Right-looking, 8-row panels.  Panel p is spilled from the accumulator tiles to shared memory; one
// lane owns one panel column (plus the rhs slice as an extra column) and the 8 pivots are eliminated
// in order: the pivot lane broadcasts 1/sqrt(d), every lane scales its row-r entry, the lanes of the
// diagonal block broadcast U[r][r'] and every lane updates its later rows.  The trailing matrix is
// then updated in registers with 3xTF32 mma tiles.  A non-positive pivot yields a non-finite
// solution, which is how failure is detected (LAPACK posv info != 0, _als.pyx:131-138).
template <int NB>
__device__ __forceinline__ void factor_solve(RowState<NB> &st, float *U, float *zb,
                                             float *__restrict__ xout, int lane, bool &ok, int dbg,
                                             float *const *peers, int n_peers, int64_t xoff) {
  using C = Cfg<NB>;
  constexpr int F = C::F;
  const int g = lane >> 2, t = lane & 3;

#pragma unroll
  for (int p = 0; p < C::NT8; ++p) {
    const int i = p >> 1, h = p & 1;
    float *Up = U + C::poff(p);
    const int sp = C::pstride(p);
    const int Wp = F - 8 * p;
    constexpr int NJmax = (F + 9 + 31) / 32;
    const int NJ = (Wp + 9 + 31) / 32;
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
    // 2. one panel column per lane.  Local columns [0, Wp) are the matrix (the first 8 = the diagonal block),
    //    Wp is the rhs slice, and Wp+1 .. Wp+8 are the unit vectors e_0..e_7: forward-substituted with the rest
    //    they become the rows of U_d^-1, which lets the back substitution resolve a whole panel at once.
    float v[NJmax][8];
#pragma unroll
    for (int j = 0; j < NJmax; ++j) {
      if (j < NJ) {
        const int c = lane + 32 * j;
        const float *colp = (c < Wp) ? (Up + c) : (zb + 8 * p);
        const int rs = (c < Wp) ? sp : 1;
        const int e = c - Wp - 1;  // unit-vector index for the identity columns
#pragma unroll
        for (int r = 0; r < 8; ++r) v[j][r] = (c <= Wp) ? colp[r * rs] : (e == r ? 1.f : 0.f);
      }
    }
    // 3. eliminate the 8 pivots, LDL^T style: the only serial chain is  1/d_r -> (one shuffle) -> the next
    //    pivot's own update; the scaling by 1/sqrt(d_r) that turns the rows into U is applied afterwards,
    //    for all 8 rows at once.  u = a[r][r2] / d_r comes from the lane that owns diagonal-block column r2.
    if (!(dbg & 2)) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float d = v[0][r];  // meaningful on lane r: the pivot
        float rc = rcp_approx(d);
        rc = rc * fmaf(-d, rc, 2.f);  // Newton step
        const float rinv = __shfl_sync(0xffffffffu, rc, r);
#pragma unroll
        for (int r2 = r + 1; r2 < 8; ++r2) {
          const float u = __shfl_sync(0xffffffffu, v[0][r], r2) * rinv;
#pragma unroll
          for (int j = 0; j < NJmax; ++j)
            if (j < NJ) v[j][r2] = fmaf(-u, v[j][r], v[j][r2]);
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float d = __shfl_sync(0xffffffffu, v[0][r], r);
        float s = rsqrt_approx(d);
        s = s * fmaf(-0.5f * d * s, s, 1.5f);  // Newton step: full fp32 accuracy
#pragma unroll
        for (int j = 0; j < NJmax; ++j)
          if (j < NJ) v[j][r] *= s;
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < NJmax; ++j) {
      if (j < NJ) {
        const int c = lane + 32 * j;
        if (c >= 8 && c < Wp) {
#pragma unroll
          for (int r = 0; r < 8; ++r) Up[r * sp + c] = v[j][r];
        } else if (c == Wp) {
#pragma unroll
          for (int r = 0; r < 8; ++r) zb[8 * p + r] = v[j][r];
        } else if (c > Wp && c <= Wp + 8) {
          // row (c - Wp - 1) of U_d^-1 replaces that row of the diagonal block (U_d itself is not needed again)
          float *dst = Up + (c - Wp - 1) * sp;
          *reinterpret_cast<float4 *>(dst) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
          *reinterpret_cast<float4 *>(dst + 4) = make_float4(v[j][4], v[j][5], v[j][6], v[j][7]);
        }
      }
    }
    __syncwarp();
    // 4. trailing update in registers: A[m][n] -= sum_r U[r][m] U[r][n]; b[m] -= sum_r U[r][m] z[r]
    if (p + 1 < C::NT8 && !(dbg & 4)) {
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
      for (int term = 0; term < 3; ++term) {  // term-major, as in consume_kstep
#pragma unroll
        for (int ib = (p + 1) >> 1; ib < NB; ++ib) {
          // rows 16 ib + g (a0, a2) are still live only if 2 ib > p
          const bool top = (2 * ib > p);
          const int jt = top ? 2 * ib : p + 1;
          const uint32_t a0 = top ? ((term == 0 ? ul0[jt] : uh0[jt]) ^ kSignBit) : 0u;
          const uint32_t a2 = top ? ((term == 0 ? ul1[jt] : uh1[jt]) ^ kSignBit) : 0u;
          const uint32_t a1 = (term == 0 ? ul0[2 * ib + 1] : uh0[2 * ib + 1]) ^ kSignBit;
          const uint32_t a3 = (term == 0 ? ul1[2 * ib + 1] : uh1[2 * ib + 1]) ^ kSignBit;
#pragma unroll
          for (int j = (2 * ib > p + 1 ? 2 * ib : p + 1); j < C::NT8; ++j) {
            float(&d)[4] = st.acc[C::tidx(ib, j)];
            if (term == 1) mma_tf32(d, a0, a1, a2, a3, ul0[j], ul1[j]);
            else mma_tf32(d, a0, a1, a2, a3, uh0[j], uh1[j]);
          }
        }
      }
    }
  }

  // 5. back substitution U x = z, column oriented and blocked by panel: lane m (and m + 32) owns z[m];
  //    each panel's 8 unknowns are resolved in order (one shuffle + one multiply on the critical path),
  //    then every earlier row folds the panel in with two 16-byte loads and 8 FMAs.
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
    rowoff[q] = po + (mm & 7) * ps - 8 * pm;  // U[m][c] lives at U[rowoff + c] for c >= 8 pm
    zz[q] = zb[mm];
    xx[q] = 0.f;
  }
  if (!(dbg & 1))
#pragma unroll
  for (int p = C::NT8 - 1; p >= 0; --p) {
    const int qp = (8 * p) >> 5;        // register slot of the panel's rows
    const int l0 = (8 * p) & 31;        // their first lane
    // this lane's row of U_d^-1 (meaningful on lanes l0..l0+7 of slot qp)
    const float4 da = *reinterpret_cast<const float4 *>(U + rowoff[qp] + 8 * p);
    const float4 db = *reinterpret_cast<const float4 *>(U + rowoff[qp] + 8 * p + 4);
    float rhs[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) rhs[c] = __shfl_sync(0xffffffffu, zz[qp], l0 + c);
    // x_r = sum_{c >= r} U_d^-1[r][c] rhs[c]  (entries below the diagonal of the stored rows are exactly 0)
    const float xm = fmaf(da.x, rhs[0], fmaf(da.y, rhs[1], fmaf(da.z, rhs[2], da.w * rhs[3]))) +
                     fmaf(db.x, rhs[4], fmaf(db.y, rhs[5], fmaf(db.z, rhs[6], db.w * rhs[7])));
    const bool in_panel = (lane >= l0) && (lane < l0 + 8);
    if (in_panel) xx[qp] = xm;
    float xs[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) xs[c] = __shfl_sync(0xffffffffu, xm, l0 + c);
    // rows before the panel fold it in
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      if (32 * q < 8 * p) {
        const int m = lane + 32 * q;
        if (m < 8 * p) {
          const float4 ua = *reinterpret_cast<const float4 *>(U + rowoff[q] + 8 * p);
          const float4 ub = *reinterpret_cast<const float4 *>(U + rowoff[q] + 8 * p + 4);
          const float s0 = fmaf(ua.x, xs[0], fmaf(ua.y, xs[1], fmaf(ua.z, xs[2], ua.w * xs[3])));
          const float s1 = fmaf(ub.x, xs[4], fmaf(ub.y, xs[5], fmaf(ub.z, xs[6], ub.w * xs[7])));
          zz[q] -= s0 + s1;
        }
      }
    }
  }
  bool fin = true;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int m = lane + 32 * q;
    if (m < F) {
      fin = fin && (fabsf(xx[q]) <= 3.0e38f);  // false for inf and NaN
    }
  }
  ok = __all_sync(0xffffffffu, fin);
  if (ok) {
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int m = lane + 32 * q;
      if (m < F) {
        xout[m] = xx[q];
        for (int pi = 0; pi < n_peers; ++pi) peers[pi][xoff + m] = xx[q];  // NVLink stores into the peer replicas
      }
    }
  }
}

// ---- kernel ------------------------------------------------------------------------------------
// pass 0: whole rows and chunks of giant rows;  pass 1: finish giant rows from their chunk slots.
template <int NB>
__global__ void __launch_bounds__(32 * kWarpsPerCta, 3)
cholesky_half_kernel(const int32_t *__restrict__ indices, const float *__restrict__ data, const float *__restrict__ Y,
                     float *__restrict__ X, int64_t row_offset, const float *__restrict__ Greg,
                     const WorkItem *__restrict__ work, int n_work, int32_t *counter, float *slots,
                     long long *bad_row, int pass, int dbg, float *const *peers, int n_peers) {
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
  auto stage_ptr = [&](int s) -> float * { return stages + (s >= C::NSTAGE ? s - C::NSTAGE : s) * C::STAGE_FLOATS; };
  const Blk kNoBlk{-1, 0.f};

  // software pipeline over work items: `wi` is being processed, `wn` (+ its first index block) is
  // already in registers, the counter for the one after is fetched a whole row ahead
  int i0 = fetch();
  if (i0 < 0) return;
  WorkItem wi = load_item(i0);
  int i1 = fetch();
  WorkItem wn{0, 0, 0, -1};
  Blk b0 = kNoBlk, nb0 = kNoBlk;
  int stage = 0;
  if (pass == 0) {
    b0 = load_block(wi, 0, indices, data, lane);
    const int nks0 = (wi.k1 - wi.k0 + 7) >> 3;
    issue_kstep<NB>(stage_ptr(0), b0, 0, 0 < nks0, Y, lane);
    issue_kstep<NB>(stage_ptr(1), b0, 1, 1 < nks0, Y, lane);
  }
  if (i1 >= 0) {
    wn = load_item(i1);
    if (pass == 0) nb0 = load_block(wn, 0, indices, data, lane);
  }

  RowState<NB> st;
  for (;;) {
    const bool whole = (wi.slot == -1), chunk = (wi.slot >= 0), finish = (wi.slot == -2);
    const int i2 = (i1 >= 0) ? fetch() : -1;  // consumed after this row's accumulation
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

    WorkItem wnn{0, 0, 0, -1};
    Blk nnb0 = kNoBlk;
    if (pass == 0) {
      const int nks = (wi.k1 - wi.k0 + 7) >> 3;
      Blk cb = b0;
      Blk nb = (nks > 4) ? load_block(wi, 1, indices, data, lane) : kNoBlk;
      for (int ks = 0; ks < nks; ++ks) {
        cp_async_wait<1>();
        __syncwarp();
        const int tks = ks + 2;
        if ((tks & 3) == 0 && tks < nks) {  // the gathers move on to the next 32 nonzeros
          cb = nb;
          if (tks + 4 < nks) nb = load_block(wi, (tks >> 2) + 1, indices, data, lane);
        }
        issue_kstep<NB>(stage_ptr(stage + 2), cb, tks & 3, tks < nks, Y, lane);
        consume_kstep<NB>(st, stage_ptr(stage), g, t);
        if (++stage == C::NSTAGE) stage = 0;
      }
      __syncwarp();
      // the first two k-steps of the next item land while this row is factored
      if (i1 >= 0) {
        const int nks1 = (wn.k1 - wn.k0 + 7) >> 3;
        issue_kstep<NB>(stage_ptr(stage), nb0, 0, 0 < nks1, Y, lane);
        issue_kstep<NB>(stage_ptr(stage + 1), nb0, 1, 1 < nks1, Y, lane);
      }
      if (i2 >= 0) {
        wnn = load_item(i2);
        nnb0 = load_block(wnn, 0, indices, data, lane);
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
      if (i2 >= 0) wnn = load_item(i2);
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
        for (int m = lane; m < F; m += 32) {
          xout[m] = 0.f;
          for (int pi = 0; pi < n_peers; ++pi) peers[pi][(row_offset + wi.row) * F + m] = 0.f;
        }
      } else if (whole || finish) {
        bool ok = true;
        if (!(dbg & 8)) factor_solve<NB>(st, U, zb, xout, lane, ok, dbg, peers, n_peers, (row_offset + wi.row) * F);
        if (!ok && lane == 0) atomicMin(bad_row, (long long)(row_offset + wi.row));
        __syncwarp();
      }
    }
    if (i1 < 0) break;
    wi = wn;
    b0 = nb0;
    i1 = i2;
    wn = wnn;
    nb0 = nnb0;
  }
  cp_async_wait<0>();
}

__global__ void init_solver_scalars(int32_t *counters, long long *bad_row) {
  if (threadIdx.x < 16) counters[threadIdx.x] = 0;
  if (threadIdx.x == 0) bad_row[0] = LLONG_MAX;
}

// Timing ablations (results are WRONG when set): ALS_B200_DEBUG bit0 skip back-substitution, bit1 skip pivot
// elimination, bit2 skip trailing updates, bit3 skip the whole factorisation.  Never set in production.
static int debug_flags() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("ALS_B200_DEBUG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int NB>
int run_cholesky(als_ctx *ctx, const als_csr *Cm, als_factors *X, const als_factors *Y) {
  using C = Cfg<NB>;
  const int dbg = debug_flags();
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
                                                          ctx->bad_row, 0, dbg, X->peers_dev, X->n_peers);
    ALS_CUDA(cudaGetLastError());
    ctx->launches++;
  }
  if (Cm->n_finish) {
    const int64_t want = ceil_div(Cm->n_finish, kWarpsPerCta);
    const int grid = (int)std::min<int64_t>(want, (int64_t)ctx->sm_count * ctas_per_sm);
    ProfScope prof(ctx, kProfCholFinish);
    kern<<<grid, 32 * kWarpsPerCta, smem, ctx->stream>>>(Cm->indices, Cm->data, Y->d, X->d, Cm->row_offset, ctx->Greg,
                                                          Cm->finish, (int)Cm->n_finish, ctx->counters + 1, slots,
                                                          ctx->bad_row, 1, dbg, X->peers_dev, X->n_peers);
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
