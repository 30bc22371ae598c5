// Supernodal, level-scheduled sparse triangular solves on the inverted-diagonal
// panels produced by hb_numfact.cu -- the dominant kernel of the RAS apply.
//
// Replaces SUBDOMAIN<K>::solve (reference call sites include/HPDDM_schwarz.hpp:
// 535,542,544,557,567,590; third-party bodies include/HPDDM_SuiteSparse.hpp:388-423,
// include/HPDDM_MUMPS.hpp:304-317).
//
// Forward sweep, level by level from the leaves (one launch per level):
//     y[piv(f)]   += W_f  * b[piv(f)]
//     b[struct(f)] -= M_f * b[piv(f)]
// Backward sweep, from the root:
//     x[piv(f)]   += P_f^T * [ y[piv(f)] ; -x[struct(f)] ]
// Each panel value is read exactly once per sweep with 128-bit coalesced loads;
// a warp owns one work item (<= 32 rows x 512 columns forward, <= 128 rows x 256
// columns backward; 1, 2 or 4 right-hand sides per pass over the panels), reduces with shuffles and publishes with FP64 atomics
// (RED.ADD.F64).  Pure HBM streaming: 0.25 flop/byte.
#include <algorithm>
#include <cstring>

#include "hb_internal.h"

namespace hb {

namespace {

__device__ __forceinline__ double2 ldg_stream(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// 8 row sums spread over the warp -> lane L (L % 4 == 0) ends up with the total of row
// 4*bit4(L) + 2*bit3(L) + bit2(L): 9 shuffles instead of 40
__device__ __forceinline__ K reduce8(K (&a)[8], int lane) {
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const K send = u16 ? a[i] : a[i + 4], keep = u16 ? a[i + 4] : a[i];
    a[i] = keep + hb_shfl_xor(send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const K send = u8 ? a[i] : a[i + 2], keep = u8 ? a[i + 2] : a[i];
    a[i] = keep + hb_shfl_xor(send, 8);
  }
  {
    const K send = u4 ? a[0] : a[1], keep = u4 ? a[1] : a[0];
    a[0] = keep + hb_shfl_xor(send, 4);
  }
  a[0] += hb_shfl_xor(a[0], 2);
  a[0] += hb_shfl_xor(a[0], 1);
  return a[0];
}

// Forward sweep work item, MU right-hand sides at once (panel values are read once for all MU).
// R = 8 / MU rows and JU = MU column slabs are in flight together (8 x 128-bit panel loads per
// lane), the R * MU = 8 partial sums go through one reduce8.
// (MU = 1 is held to 64 registers -> 4 CTAs / SM: measured +2 % at m = 128 over the 80-register build, profiles/README.md)
// A 128-bit load carries VE elements of K: two reals or one complex (hb_scalar.h); complex accumulators
// double the register budget, so that build runs 2 CTAs / SM.
// WIDE = 2 (block right-hand sides only): the staged chunk covers 2 FCH / MU columns per pass (64 KB of dynamic shared memory per
// CTA) -- half as many shuffle reductions per panel byte as WIDE = 1.
template <int MU, int WIDE = 1>
__global__ void __launch_bounds__(256, IS_COMPLEX ? 2 : (MU == 1 ? 4 : 3)) k_fwd(const FwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts,
                                                const int *__restrict__ rowidx, const K *__restrict__ pan, K *b, K *y, int n) {
  constexpr int R = 8 / MU, JU = MU, CW = WIDE * FCH / MU;
  // WIDE = 1: 32 KB of static shared memory (the layout the MU = 1 kernel was tuned with); WIDE = 2: 64 KB, dynamic (opt-in size)
  __shared__ __align__(16) K bs_static[WIDE == 1 ? 8 : 1][WIDE == 1 ? FCH : 1];
  extern __shared__ __align__(16) unsigned char bs_raw[];
  K(*bs)[WIDE * FCH] = WIDE == 1 ? reinterpret_cast<K(*)[WIDE * FCH]>(bs_static) : reinterpret_cast<K(*)[WIDE * FCH]>(bs_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  const FwdItem w = items[it];
  const Front f = fronts[w.front];
  const int s1 = f.s1, nb1 = (s1 + RB - 1) / RB;
  const K *base;
  int nrows, stride, cmax;
  const bool pivot = w.rblk < nb1;
  if (pivot) {
    const int k = w.rblk;
    stride = hb_wblk(s1, k);
    base = pan + f.poff + hb_blk_off(k);
    nrows = min(RB, s1 - RB * k);
    cmax = min(s1, RB * (k + 1));
  } else {
    const int k2 = w.rblk - nb1;
    stride = hb_ldp(s1);
    base = pan + f.poff + hb_upd_off(s1) + (int64_t)k2 * RB * stride;
    nrows = min(RB, f.s2 - RB * k2);
    cmax = s1;
  }
  const int c1 = min(cmax, w.c0 + w.cw);
  K *mybs = bs[warp];
  const double2 *bs2 = reinterpret_cast<const double2 *>(mybs);
  const int st2 = stride / VE;  // row stride in 128-bit vectors
  const int rsel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  for (int cs = w.c0; cs < c1; cs += CW) {
    const int nc = min(CW, c1 - cs);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < MU; ++m) {
      for (int c = lane; c < nc; c += 32) mybs[m * CW + c] = b[(int64_t)m * n + f.p0 + cs + c];
      if (VE == 2 && (nc & 1) && lane == 0) mybs[m * CW + nc] = mk(0.0);  // real panels are zero-padded to even widths
    }
    __syncwarp();
    const int nv = (nc + VE - 1) / VE;
    for (int r = 0; r < nrows; r += R) {
      const double2 *p = reinterpret_cast<const double2 *>(base + (int64_t)r * stride + cs);
      const int nr = min(R, nrows - r);
      K a[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = mk(0.0);
      for (int j0 = lane; j0 < nv; j0 += 32 * JU) {
        double2 t[R][JU];
#pragma unroll
        for (int u = 0; u < JU; ++u)
#pragma unroll
          for (int q = 0; q < R; ++q) t[q][u] = (q < nr && j0 + 32 * u < nv) ? ldg_stream(p + (int64_t)q * st2 + j0 + 32 * u) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < JU; ++u) {
          if (j0 + 32 * u < nv) {
#pragma unroll
            for (int m = 0; m < MU; ++m) {
              const double2 bb = bs2[m * (CW / VE) + j0 + 32 * u];
#pragma unroll
              for (int q = 0; q < R; ++q) a[q * MU + m] = hb_vdot(t[q][u], bb, a[q * MU + m]);
            }
          }
        }
      }
      const K v = reduce8(a, lane);
      const int q = rsel / MU, m = rsel % MU;
      if ((lane & 3) == 0 && q < nr) {
        if (pivot) hb_atomic_add(&y[(int64_t)m * n + f.p0 + RB * w.rblk + r + q], v);
        else hb_atomic_add(&b[(int64_t)m * n + rowidx[f.rptr + RB * (w.rblk - nb1) + r + q]], -v);
      }
    }
  }
}

// VE right-hand-side values b[c .. c+VE) as one 128-bit vector, from L1 (read-only in this launch: a level's fronts only
// update rows of their ancestors).  Real scalars: the address is only 8-byte aligned in general -> two loads; the partner of
// the last column of an odd-width item is a zero (it multiplies the zero padding of the panel, but must not be a NaN).
template <bool COHERENT = false>
__device__ __forceinline__ double2 ld_rhs(const K *p, int c, int nc) {
  // COHERENT (persistent kernels: the vector is updated by other SMs during the launch): read through L2, never a stale L1 line
#ifdef HB_COMPLEX
  return COHERENT ? __ldcg(reinterpret_cast<const double2 *>(p + c)) : __ldg(reinterpret_cast<const double2 *>(p + c));
#else
  if (COHERENT) return make_double2(__ldcg(p + c), c + 1 < nc ? __ldcg(p + c + 1) : 0.0);
  return make_double2(__ldg(p + c), c + 1 < nc ? __ldg(p + c + 1) : 0.0);
#endif
}
__device__ __forceinline__ K ld_cg(const K *p) {
#ifdef HB_COMPLEX
  const double2 v = __ldcg(reinterpret_cast<const double2 *>(p));
  return mk(v.x, v.y);
#else
  return __ldcg(p);
#endif
}

// Forward sweep work item for MU = 2 / 4 right-hand sides.  Same data flow as k_fwd, but the right-hand-side block is read
// through L1 instead of being staged in shared memory, so that one pass covers the whole item width (up to FCH columns) for
// all MU columns: the shuffle reduction (reduce8) runs once per R = 8 / MU rows x FCH columns instead of once per FCH / MU
// columns -- it dominated the staged variant at MU = 4 (profiles/README.md).
// Narrow fronts (row width 8, 16 or 32: the small separators of the leaf-side levels, thousands per level).  The generic forward item
// gives such a front 8 rows x (width / 2) lanes per step -- 4 to 16 active lanes; ncu at 160^3 (profiles/r02_ncu_full_sptrsv_m160.csv)
// shows those levels at 1.8-3.9 TB/s.  A block of rows of a narrow panel is CONTIGUOUS (row stride = width), so k_fwd_narrow reads
// it as a flat array: every lane one 128-bit load per step, 32 / G rows per step with G = width / 2 lanes per row, a G-lane shuffle
// reduction, one RED per row.  Same item list as k_fwd_blk<1>: each of the two kernels skips the other's items (levels that have only
// one kind get only that launch).
__host__ __device__ inline bool hb_narrow_front(int s1) {
  const int l = hb_ldp(s1);
  return !IS_COMPLEX && (l == 8 || l == 16 || l == 32);
}
#ifndef HB_COMPLEX
__global__ void __launch_bounds__(256) k_fwd_narrow(const FwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts, const int *__restrict__ rowidx,
                                                    const double *__restrict__ pan, double *b, double *y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  const FwdItem w = items[it];
  const Front f = fronts[w.front];
  const int s1 = f.s1;
  if (!hb_narrow_front(s1)) return;
  const int stride = hb_ldp(s1);  // one pivot block (s1 <= 32): block 0 of the trapezoid has the full width too
  const bool pivot = w.rblk == 0;
  const double *base = pivot ? pan + f.poff : pan + f.poff + hb_upd_off(s1) + (int64_t)(w.rblk - 1) * RB * stride;
  const int nrows = pivot ? s1 : min(RB, f.s2 - RB * (w.rblk - 1));
  const int G = stride >> 1, rpl = 32 / G;  // lanes per row, rows per warp-wide load
  const int sub = lane % G, rsub = lane / G;
  const double *bc = b + f.p0;
  const double2 bz = make_double2(2 * sub < s1 ? __ldg(bc + 2 * sub) : 0.0, 2 * sub + 1 < s1 ? __ldg(bc + 2 * sub + 1) : 0.0);
  const double2 *flat = reinterpret_cast<const double2 *>(base) + lane;
  const int *rows = rowidx + f.rptr + (pivot ? 0 : RB * (w.rblk - 1));
  const int nloads = (nrows + rpl - 1) / rpl;
  for (int q0 = 0; q0 < nloads; q0 += 8) {
    double2 t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) t[u] = ((q0 + u) * rpl + rsub < nrows) ? ldg_stream(flat + 32 * (q0 + u)) : make_double2(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      double v = fma(t[u].x, bz.x, t[u].y * bz.y);
      for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int r = (q0 + u) * rpl + rsub;
      if (sub == 0 && r < nrows) {
        if (pivot) atomicAdd(&y[f.p0 + r], v);
        else atomicAdd(&b[rows[r]], -v);
      }
    }
  }
}
#endif

template <int MU, bool COHERENT>
__device__ __forceinline__ void fwd_blk_item(const FwdItem &w, int lane, const Front *__restrict__ fronts, const int *__restrict__ rowidx, const K *__restrict__ pan, K *b, K *y, int n) {
  constexpr int R = 8 / MU, JU = MU;
  const Front f = fronts[w.front];
  const int s1 = f.s1, nb1 = (s1 + RB - 1) / RB;
  const K *base;
  int nrows, stride, cmax;
  const bool pivot = w.rblk < nb1;
  if (pivot) {
    const int k = w.rblk;
    stride = hb_wblk(s1, k);
    base = pan + f.poff + hb_blk_off(k);
    nrows = min(RB, s1 - RB * k);
    cmax = min(s1, RB * (k + 1));
  } else {
    const int k2 = w.rblk - nb1;
    stride = hb_ldp(s1);
    base = pan + f.poff + hb_upd_off(s1) + (int64_t)k2 * RB * stride;
    nrows = min(RB, f.s2 - RB * k2);
    cmax = s1;
  }
  const int nc = min(cmax, w.c0 + w.cw) - w.c0;  // columns of this item
  if (nc <= 0) return;
  const K *bc = b + f.p0 + w.c0;  // column m of the right-hand-side block: bc + m * n
  const int nv = (nc + VE - 1) / VE;
  const int st2 = stride / VE;  // row stride in 128-bit vectors
  const int rsel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  for (int r = 0; r < nrows; r += R) {
    const double2 *p = reinterpret_cast<const double2 *>(base + (int64_t)r * stride + w.c0);
    const int nr = min(R, nrows - r);
    K a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = mk(0.0);
    for (int j0 = lane; j0 < nv; j0 += 32 * JU) {
      double2 t[R][JU];
#pragma unroll
      for (int u = 0; u < JU; ++u)
#pragma unroll
        for (int q = 0; q < R; ++q) t[q][u] = (q < nr && j0 + 32 * u < nv) ? ldg_stream(p + (int64_t)q * st2 + j0 + 32 * u) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < JU; ++u) {
        if (j0 + 32 * u < nv) {
#pragma unroll
          for (int m = 0; m < MU; ++m) {
            const double2 bb = ld_rhs<COHERENT>(bc + (int64_t)m * n, (j0 + 32 * u) * VE, nc);
#pragma unroll
            for (int q = 0; q < R; ++q) a[q * MU + m] = hb_vdot(t[q][u], bb, a[q * MU + m]);
          }
        }
      }
    }
    const K v = reduce8(a, lane);
    const int q = rsel / MU, m = rsel % MU;
    if ((lane & 3) == 0 && q < nr) {
      if (pivot) hb_atomic_add(&y[(int64_t)m * n + f.p0 + RB * w.rblk + r + q], v);
      else hb_atomic_add(&b[(int64_t)m * n + rowidx[f.rptr + RB * (w.rblk - nb1) + r + q]], -v);
    }
  }
}


template <int MU>
__global__ void __launch_bounds__(256, (IS_COMPLEX || MU == 4) ? 2 : (MU == 1 ? 4 : 3)) k_fwd_blk(const FwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts,
                                                                     const int *__restrict__ rowidx, const K *__restrict__ pan, K *b, K *y, int n, int skip_narrow = 0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  const FwdItem w = items[it];
  if (MU == 1 && skip_narrow && hb_narrow_front(fronts[w.front].s1)) return;  // k_fwd_narrow takes these
  fwd_blk_item<MU, false>(w, lane, fronts, rowidx, pan, b, y, n);
}

// Backward sweep work item.  NJ = slabs of 32 * VE columns covered per pass (NJ * MU <= 4 128-bit
// accumulators per lane); rows are processed in groups of 8 / NJ so that 8 128-bit loads are in flight.
// With SHARED (block right-hand sides) the multipliers u of a 32-row block are broadcast from shared memory (us: 32 * MU values
// of this warp) instead of one shuffle per (row, column): MU = 4 needed 8 SHFL per 128-bit panel load.
template <int NJ, int MU, bool SHARED, bool COHERENT = false>
__device__ __forceinline__ void bwd_pass(const BwdItem &w, const Front &f, int cbase, const int *__restrict__ rowidx, const K *__restrict__ pan,
                                         const K *__restrict__ y, K *x, int n, int lane, K *us) {
  constexpr int G = 8 / NJ;
  const int s1 = f.s1, ldp = hb_ldp(s1);
  const K *P = pan + f.poff;
  const K *Pu = P + hb_upd_off(s1);
  double2 acc[NJ][MU];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int m = 0; m < MU; ++m) acc[j][m] = make_double2(0.0, 0.0);
  const int cl = cbase + VE * lane;  // this lane's first column
  for (int rb = 0; rb < w.nr; rb += 32) {
    const int r = w.r0 + rb + lane;
    K u[MU];
#pragma unroll
    for (int m = 0; m < MU; ++m) {
      u[m] = mk(0.0);
      if (rb + lane < w.nr) {
        const K *src = (r < s1) ? y + (int64_t)m * n + f.p0 + r : x + (int64_t)m * n + rowidx[f.rptr + r - s1];
        const K v = COHERENT ? ld_cg(src) : *src;  // persistent kernel: other SMs finished these entries during this launch
        u[m] = (r < s1) ? v : -v;
      }
    }
    if (SHARED) {
      __syncwarp();  // the previous block's readers are done
#pragma unroll
      for (int m = 0; m < MU; ++m) us[lane * MU + m] = u[m];
      __syncwarp();
    }
    const int nq = min(32, w.nr - rb);
    for (int q = 0; q < nq; q += G) {
      double2 t[G][NJ];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int rr = w.r0 + rb + q + g;
        const K *rowp;
        int wlim;
        if (rr < s1) {
          const int k = rr / RB;
          wlim = hb_wblk(s1, k);
          rowp = P + hb_blk_off(k) + (int64_t)(rr - k * RB) * wlim;
        } else {
          wlim = ldp;
          rowp = Pu + (int64_t)(rr - s1) * ldp;
        }
        if (q + g >= nq) wlim = 0;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int c = cl + 32 * VE * j;
          t[g][j] = (c < wlim) ? ldg_stream(reinterpret_cast<const double2 *>(rowp + c)) : make_double2(0.0, 0.0);
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int m = 0; m < MU; ++m) {
          const K uq = SHARED ? us[((q + g) & 31) * MU + m] : hb_shfl(u[m], (q + g) & 31);
#pragma unroll
          for (int j = 0; j < NJ; ++j) hb_vaxpy(t[g][j], uq, acc[j][m]);
        }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = cl + 32 * VE * j;
#pragma unroll
    for (int m = 0; m < MU; ++m) hb_vpublish(x + (int64_t)m * n + f.p0, c, s1, acc[j][m]);
  }
}

template <int MU, bool SHARED, bool COHERENT>
__device__ __forceinline__ void bwd_item(const BwdItem &w, int lane, const Front *__restrict__ fronts, const int *__restrict__ rowidx, const K *__restrict__ pan, const K *y, K *x, int n,
                                         K *us) {
  const Front f = fronts[w.front];
  const int width = min(BCH, hb_ldp(f.s1) - w.c0);
  constexpr int NJMAX = 4 / MU;  // 4, 2, 1
  constexpr int SLAB = 32 * VE;  // columns covered by one 128-bit load per lane
  for (int cb = 0; cb < width; cb += SLAB * NJMAX) {
    const int left = width - cb;
    if (NJMAX >= 4 && left > 2 * SLAB) bwd_pass<(NJMAX >= 4 ? 4 : NJMAX), MU, SHARED, COHERENT>(w, f, w.c0 + cb, rowidx, pan, y, x, n, lane, us);
    else if (NJMAX >= 2 && left > SLAB) bwd_pass<(NJMAX >= 2 ? 2 : NJMAX), MU, SHARED, COHERENT>(w, f, w.c0 + cb, rowidx, pan, y, x, n, lane, us);
    else bwd_pass<1, MU, SHARED, COHERENT>(w, f, w.c0 + cb, rowidx, pan, y, x, n, lane, us);
  }
}

template <int MU, bool SHARED, int OCC = ((IS_COMPLEX || MU == 4) ? 2 : 3)>
__global__ void __launch_bounds__(256, OCC) k_bwd(const BwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts,
                                                const int *__restrict__ rowidx, const K *__restrict__ pan, const K *__restrict__ y, K *x, int n) {
  __shared__ __align__(16) K us_all[SHARED ? 8 * 32 * MU : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  K *us = us_all + (SHARED ? warp * 32 * MU : 0);
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  bwd_item<MU, SHARED, false>(items[it], lane, fronts, rowidx, pan, y, x, n, us);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Dependency-driven persistent sweeps (one launch per sweep instead of one per elimination-tree level).  CTAs claim groups of 8
// work items in list order (level order: leaves first in the forward sweep, root first in the backward sweep) from a global counter;
// a warp starts its item as soon as the fronts it depends on are complete --
//   forward:  every child of the item's front has finished all its items   (their updates to b[pivots of this front] are in)
//   backward: the parent front has finished all its items                  (x[struct rows of this front] is final)
// -- instead of waiting for the whole previous level: the tails of the many small levels overlap, and the launch / drain gaps between
// levels disappear.  An item only ever waits for items EARLIER in the list, which have been claimed by resident CTAs: no deadlock;
// the spin is bounded anyway (error flag, checked at the next synchronisation point).  Vectors written during the launch (b, y, x)
// are read through L2 (COHERENT).
struct SweepSync {
  int *pending;       // forward: children of front f not yet complete (reset from pending0 before every solve)
  int *done;          // items of front f completed in the running sweep
  const int *total;   // items of front f in this sweep
  unsigned long long *next;  // item claim counter
  int *err;
};
__device__ __forceinline__ bool spin_until_zero(const int *p) {
  for (long long k = 0; k < (1LL << 24); ++k) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (v <= 0) return true;
    __nanosleep(64);
  }
  return false;
}
template <int OCC>
__global__ void __launch_bounds__(256, OCC) k_fwd_persistent(const FwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts, const int *__restrict__ rowidx,
                                                             const K *__restrict__ pan, K *b, K *y, int n, SweepSync sy) {
  __shared__ unsigned long long base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) base_s = atomicAdd(sy.next, 8ULL);
    __syncthreads();
    const int64_t it = (int64_t)base_s + warp;
    if ((int64_t)base_s >= nitems) return;
    if (it >= nitems) continue;
    const FwdItem w = items[it];
    bool ok = true;
    if (lane == 0) ok = spin_until_zero(sy.pending + w.front);
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) {
      if (lane == 0) *reinterpret_cast<volatile int *>(sy.err) = 1;
    } else
      fwd_blk_item<1, true>(w, lane, fronts, rowidx, pan, b, y, n);
    __threadfence();  // this item's updates are visible device-wide before it is counted
    __syncwarp();
    if (lane == 0) {
      if (atomicAdd(sy.done + w.front, 1) + 1 == sy.total[w.front]) {  // last item of the front: release the parent
        const int parent = fronts[w.front].parent;
        if (parent >= 0) atomicSub(sy.pending + parent, 1);
      }
    }
  }
}
template <int OCC>
__global__ void __launch_bounds__(256, OCC) k_bwd_persistent(const BwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts, const int *__restrict__ rowidx,
                                                             const K *__restrict__ pan, const K *y, K *x, int n, SweepSync sy) {
  __shared__ unsigned long long base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) base_s = atomicAdd(sy.next, 8ULL);
    __syncthreads();
    const int64_t it = (int64_t)base_s + warp;
    if ((int64_t)base_s >= nitems) return;
    if (it >= nitems) continue;
    const BwdItem w = items[it];
    const int parent = fronts[w.front].parent;
    bool ok = true;
    if (lane == 0 && parent >= 0) ok = spin_until_zero(sy.pending + parent);  // backward: pending[f] = items of f still to do
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) {
      if (lane == 0) *reinterpret_cast<volatile int *>(sy.err) = 1;
    } else
      bwd_item<1, false, true>(w, lane, fronts, rowidx, pan, y, x, n, nullptr);
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicSub(sy.pending + w.front, 1);
  }
}

__global__ void k_perm_in(int n, int mu, const int *__restrict__ perm, const K *__restrict__ in, K *b, K *y, K *x) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < (int64_t)n * mu) {
    const int i = (int)(t % n);
    const int64_t c = t / n;
    b[t] = in[c * n + perm[i]];
    y[t] = mk(0.0);
    x[t] = mk(0.0);
  }
}
__global__ void k_perm_out(int n, int mu, const int *__restrict__ perm, const K *__restrict__ x, const double *__restrict__ d, K *out, int accumulate) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < (int64_t)n * mu) {
    const int i = (int)(t % n);
    const int64_t c = t / n;
    const int p = perm[i];
    K v = x[t];
    if (d) v = d[p] * v;
    out[c * n + p] = accumulate ? out[c * n + p] + v : v;
  }
}

// Block right-hand sides (MU >= 2).  Forward sweep: MU = 2 -> k_fwd_blk (right-hand sides through L1, one reduction per full item
// width); MU = 4, real scalars -> k_fwd<4, 2> (staged in 64 KB of shared memory, 2 FCH / MU columns per pass: the L1 variant is
// occupancy-bound at 106 registers).  Backward sweep: multipliers broadcast from shared memory; MU = 4 real at 78 registers / 3 CTAs
// per SM.  Measured at m = 96 (profiles/README.md): 4 RHS 5.04 ms (first generation) -> 4.48 ms, 2 RHS 3.12 -> 3.09 ms.
// HPDDM_B200_BLK = staged | l1 | wide and HPDDM_B200_BWD4 = 2 | 3 override the choice for A/B measurements.
template <int MU>
static int launch_levels(Sub *s, cudaStream_t st) {
  static const char *env = getenv("HPDDM_B200_BLK");
  static const bool staged = env && !strcmp(env, "staged");
  static const bool wide = env ? !strcmp(env, "wide") : (MU == 4 && !IS_COMPLEX);
  static const bool bwd3 = getenv("HPDDM_B200_BWD4") ? !strcmp(getenv("HPDDM_B200_BWD4"), "3") : true;
  const bool blk = MU > 1 && !staged;
  // MU = 1, real scalars: the forward sweep reads its right-hand-side chunk (4 KB per warp, L1-resident) through L1 as well instead of
  // staging it in shared memory -- same-box A/B at m = 128: 8.51 -> 8.25 ms, at m = 64: 0.728 -> 0.693 ms (profiles/README.md);
  // HPDDM_B200_FWD1=staged selects the shared-memory kernel
  static const bool fwd1_l1 = getenv("HPDDM_B200_FWD1") ? !strcmp(getenv("HPDDM_B200_FWD1"), "l1") : !IS_COMPLEX;
  constexpr size_t smem2 = (size_t)2 * 8 * FCH * sizeof(K);
  DeviceFactor &D = s->fac;
  const Symbolic &S = s->sym;
  const int n = S.n;
  for (int l = 0; l < S.nlevels; ++l) {
    const int64_t i0 = S.fwd_ptr[l], ni = S.fwd_ptr[l + 1] - i0;
    if (ni <= 0) continue;
    if (MU == 1 && fwd1_l1) {
      // HPDDM_B200_NARROW=1: items of narrow fronts go to k_fwd_narrow, the rest to the generic kernel
      int64_t narrow = 0;
#ifndef HB_COMPLEX
      // opt-in: same-box A/B (profiles/README.md) 8.398 -> 8.434 ms at 128^3, 0.720 -> 0.734 ms at 64^3, 20.21 -> 20.26 ms at 160^3 -- the
      // narrow levels hold 0.5 GB of 60, and the second launch per level costs what the better lane use gains
      static const bool narrow_on = getenv("HPDDM_B200_NARROW") && !strcmp(getenv("HPDDM_B200_NARROW"), "1");
      if (narrow_on)
        for (int64_t q = i0; q < i0 + ni; ++q) narrow += hb_narrow_front(S.fronts[S.fwd[q].front].s1);
      if (narrow > 0) k_fwd_narrow<<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y);
#endif
      if (narrow < ni) k_fwd_blk<1><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y, n, narrow > 0 ? 1 : 0);
    }
    else if (MU > 1 && wide) k_fwd<(MU > 1 ? MU : 2), 2><<<(unsigned)((ni + 7) / 8), 256, smem2, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y, n);
    else if (blk) k_fwd_blk<(MU > 1 ? MU : 2)><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y, n);
    else k_fwd<MU><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y, n);
  }
  for (int l = S.nlevels - 1; l >= 0; --l) {
    const int64_t i0 = S.bwd_ptr[l], ni = S.bwd_ptr[l + 1] - i0;
    if (ni <= 0) continue;
    if (blk && MU == 4 && bwd3 && !IS_COMPLEX) k_bwd<MU, (MU > 1), 3><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.bwd + i0, ni, D.fronts, D.rowidx, D.panU, D.y, D.x, n);
    else if (blk) k_bwd<MU, (MU > 1)><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.bwd + i0, ni, D.fronts, D.rowidx, D.panU, D.y, D.x, n);
    else k_bwd<MU, false><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.bwd + i0, ni, D.fronts, D.rowidx, D.panU, D.y, D.x, n);
  }
  HB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace


// tables of the persistent sweeps, built once per factorisation from the host-side symbolic structure
static int persistent_tables(Sub *s) {
  DeviceFactor &D = s->fac;
  const Symbolic &S = s->sym;
  cudaStream_t st = s->ctx->stream;
  const int F = (int)S.fronts.size();
  D.nfronts = F;
  std::vector<int> children(F, 0), ft(F, 0), bt(F, 0);
  for (int f = 0; f < F; ++f)
    if (S.fronts[f].parent >= 0) children[S.fronts[f].parent]++;
  for (const FwdItem &w : S.fwd) ft[w.front]++;
  for (const BwdItem &w : S.bwd) bt[w.front]++;
  // a front without forward items (no columns) would never release its parent: count it as complete from the start
  for (int f = 0; f < F; ++f)
    if (ft[f] == 0 && S.fronts[f].parent >= 0) children[S.fronts[f].parent]--;
  std::vector<BwdItem> ordered;
  ordered.reserve(S.bwd.size());
  for (int l = S.nlevels - 1; l >= 0; --l) ordered.insert(ordered.end(), S.bwd.begin() + S.bwd_ptr[l], S.bwd.begin() + S.bwd_ptr[l + 1]);
  auto upl = [&](const void *h, size_t bytes, void **d) -> int {
    HB_CUDA(cudaMalloc(d, std::max<size_t>(bytes, 16)));
    if (bytes) HB_CUDA(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  };
  HB_CHECK(upl(children.data(), F * sizeof(int), (void **)&D.fwd_children));
  HB_CHECK(upl(ft.data(), F * sizeof(int), (void **)&D.fwd_total));
  HB_CHECK(upl(bt.data(), F * sizeof(int), (void **)&D.bwd_total));
  HB_CHECK(upl(ordered.data(), ordered.size() * sizeof(BwdItem), (void **)&D.bwd_ordered));
  HB_CUDA(cudaMalloc(&D.sync_pending, std::max<size_t>(F, 4) * sizeof(int)));
  HB_CUDA(cudaMalloc(&D.sync_done, std::max<size_t>(F, 4) * sizeof(int)));
  HB_CUDA(cudaMalloc(&D.sync_next, sizeof(unsigned long long)));
  HB_CUDA(cudaHostAlloc(&D.sync_err_host, sizeof(int), cudaHostAllocMapped));
  *D.sync_err_host = 0;
  HB_CUDA(cudaHostGetDevicePointer(&D.sync_err, D.sync_err_host, 0));
  HB_CUDA(cudaStreamSynchronize(st));
  return 0;
}
int sptrsv_check(Sub *s) {
  if (s->fac.sync_err_host && *reinterpret_cast<volatile int *>(s->fac.sync_err_host)) {
    set_error("persistent SpTRSV sweep: a work item timed out waiting for the fronts it depends on");
    return HPDDM_B200_ERR_STATE;
  }
  return 0;
}
// HPDDM_B200_PERSISTENT=1: single right-hand side sweeps as two dependency-driven persistent launches instead of one launch per level.
// Opt-in: same-box A/B pairs (profiles/README.md) are mixed -- 128^3: 7.92 -> 7.73 ms (0.928 -> 0.950 of the HBM roofline), 160^3:
// 19.65 -> 20.15 ms, 64^3: 0.694 -> 0.748 ms; what the overlap of the level tails gains, the L2-coherent vector reads, the lower
// occupancy of the forward kernel (80 registers) and the claim / release traffic take back.
static bool persistent_on(const Symbolic &) {
  static const bool on = getenv("HPDDM_B200_PERSISTENT") && !strcmp(getenv("HPDDM_B200_PERSISTENT"), "1");
  return on;
}
static int launch_persistent(Sub *s, cudaStream_t st) {
  DeviceFactor &D = s->fac;
  const Symbolic &S = s->sym;
  const int n = S.n, F = D.nfronts;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  constexpr int OCC = IS_COMPLEX ? 2 : 3;
  const int64_t nf = (int64_t)S.fwd.size(), nb = (int64_t)S.bwd.size();
  SweepSync sy{D.sync_pending, D.sync_done, D.fwd_total, D.sync_next, D.sync_err};
  HB_CUDA(cudaMemcpyAsync(D.sync_pending, D.fwd_children, F * sizeof(int), cudaMemcpyDeviceToDevice, st));
  HB_CUDA(cudaMemsetAsync(D.sync_done, 0, F * sizeof(int), st));
  HB_CUDA(cudaMemsetAsync(D.sync_next, 0, sizeof(unsigned long long), st));
  if (nf > 0) k_fwd_persistent<OCC><<<(unsigned)std::min<int64_t>((int64_t)sms * OCC, (nf + 7) / 8), 256, 0, st>>>(D.fwd, nf, D.fronts, D.rowidx, D.panL, D.b, D.y, n, sy);
  HB_CUDA(cudaMemcpyAsync(D.sync_pending, D.bwd_total, F * sizeof(int), cudaMemcpyDeviceToDevice, st));
  HB_CUDA(cudaMemsetAsync(D.sync_next, 0, sizeof(unsigned long long), st));
  if (nb > 0) k_bwd_persistent<OCC><<<(unsigned)std::min<int64_t>((int64_t)sms * OCC, (nb + 7) / 8), 256, 0, st>>>(D.bwd_ordered, nb, D.fronts, D.rowidx, D.panU, D.y, D.x, n, sy);
  HB_CUDA(cudaGetLastError());
  return 0;
}

#ifndef HB_COMPLEX
// ---------------------------------------------------------------------------------------------------------------------------------
// Block right-hand sides on the FP64 tensor pipe: 1 .. 8 columns per pass over the panels.
//
// Data path: every warp is a self-contained producer/consumer pipeline.  Panel tiles (32 rows x 64 columns, one row = one 512-byte
// 1-D bulk copy, cp.async.bulk -> SASS UBLKCP) land in a per-warp ring of NST shared-memory stages; completion is tracked by one
// mbarrier per stage (expect_tx / complete_tx), so no register holds data in flight -- the limit of the register-tiled kernels above
// (MU = 4: 106 registers, 25 % occupancy, latency-bound at 0.5 of the HBM roofline).  The warp walks a STREAM of stages that runs
// across work-item boundaries (persistent warps: item = global warp index + k * total warps), so the ring stays full on the levels
// made of thousands of small fronts too.  The products are mma.sync.aligned.m8n8k4.f64 (SASS DMMA) with the right-hand-side block
// as the N dimension (padded to 8: one kernel serves 1 .. 8 columns):
//   forward   D[8 rows x 8 rhs]  += P[8 rows x 4 cols]   * b[4 cols x 8 rhs]      -- no cross-lane reduction left
//   backward  D[8 cols x 8 rhs]  += P^T[8 cols x 4 rows] * u[4 rows x 8 rhs]
// A fragments are 128-bit shared-memory loads feeding two MMAs each (even / odd columns); the row pitch of the stage (72 resp. 68
// doubles) makes them bank-conflict free for the two access patterns.  Panel bytes are still read exactly once per sweep.
namespace mma {
// Two ways of filling the ring: KIND 0 = one 1-D bulk copy per tile row (cp.async.bulk, SASS UBLKCP, mbarrier complete_tx; 32-row
// tiles, 4 warps per CTA); KIND 1 = per-lane 16-byte asynchronous copies (cp.async.cg, SASS LDGSTS, commit / wait groups; 16-row
// tiles, 8 warps per CTA).  Same consumers.
template <int KIND>
struct Cfg {
  static constexpr int ROWS = KIND == 0 ? 32 : 16, COLS = 64, NST = 3, WARPS = KIND == 0 ? 4 : 8;
  static constexpr int PITCH_F = COLS + 8;  // forward: lanes (g, t) read [8 rg + g][8 j + 2 t]
  static constexpr int PITCH_B = COLS + 4;  // backward: lanes (g, t) read [4 ks + t][16 jj + 2 g]
  static constexpr size_t SMEM_F = (size_t)WARPS * NST * ROWS * PITCH_F * sizeof(double) + WARPS * NST * sizeof(uint64_t);
  static constexpr size_t SMEM_B = (size_t)WARPS * NST * ROWS * PITCH_B * sizeof(double) + WARPS * NST * sizeof(uint64_t);
};

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared (bytes: multiple of 16; both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
// 16-byte asynchronous copy global -> shared, bypassing L1; src_bytes = 0 writes zeros without touching global memory
__device__ __forceinline__ void cp_async16(void *dst, const void *src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// one stage of the stream = one ROWS x 64 tile of one work item
struct FStage {  // forward
  int64_t it;    // item index (>= nitems: stream exhausted)
  int cs;        // first column of the tile
  int rs;        // first row of the tile inside the item's 32-row block
  const double *base;  // first row of the item's row block
  int stride, nrows, c0, c1, p0, rblk, nb1;
  int64_t rptr;
  bool pivot;
};
__device__ __forceinline__ void f_decode(FStage &s, const FwdItem *items, const Front *fronts, const double *pan) {
  const FwdItem w = items[s.it];
  const Front f = fronts[w.front];
  const int s1 = f.s1;
  s.nb1 = (s1 + RB - 1) / RB;
  s.pivot = w.rblk < s.nb1;
  int cmax;
  if (s.pivot) {
    s.stride = hb_wblk(s1, w.rblk);
    s.base = pan + f.poff + hb_blk_off(w.rblk);
    s.nrows = min(RB, s1 - RB * w.rblk);
    cmax = min(s1, RB * (w.rblk + 1));
  } else {
    const int k2 = w.rblk - s.nb1;
    s.stride = hb_ldp(s1);
    s.base = pan + f.poff + hb_upd_off(s1) + (int64_t)k2 * RB * s.stride;
    s.nrows = min(RB, f.s2 - RB * k2);
    cmax = s1;
  }
  s.c0 = w.c0;
  s.c1 = min(cmax, w.c0 + w.cw);
  s.p0 = f.p0;
  s.rblk = w.rblk;
  s.rptr = f.rptr;
  s.cs = w.c0;
  s.rs = 0;
}
// advance to the next tile of the stream: next row tile of the column chunk, next column chunk, else the warp's next non-empty item
template <int ROWS>
__device__ __forceinline__ void f_next(FStage &s, bool first, int64_t nitems, int64_t stride_items, const FwdItem *items, const Front *fronts, const double *pan) {
  if (!first) {
    s.rs += ROWS;
    if (s.rs < s.nrows) return;
    s.rs = 0;
    s.cs += 64;
    if (s.cs < s.c1) return;
    s.it += stride_items;
  }
  while (s.it < nitems) {
    f_decode(s, items, fronts, pan);
    if (s.cs < s.c1 && s.nrows > 0) return;
    s.it += stride_items;
  }
}

template <int KIND>
__global__ void __launch_bounds__(Cfg<KIND>::WARPS * 32, 1) k_fwd_mma(const FwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts, const int *__restrict__ rowidx,
                                                                     const double *__restrict__ pan, double *b, double *y, int n, int mu) {
  typedef Cfg<KIND> C;
  constexpr int ROWS = C::ROWS, NST = C::NST, WARPS = C::WARPS, PITCH = C::PITCH_F, RG = ROWS / 8;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double *ring = reinterpret_cast<double *>(smem_raw) + (size_t)warp * NST * ROWS * PITCH;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)WARPS * NST * ROWS * PITCH * sizeof(double)) + warp * NST;
  if (KIND == 0) {
    for (int i = lane; i < NST * ROWS * PITCH; i += 32) ring[i] = 0.0;  // stale tails are multiplied by zeros: keep them finite
    if (lane == 0)
      for (int q = 0; q < NST; ++q) mbar_init(bars + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, gstride = (int64_t)gridDim.x * WARPS;
  FStage prod, cons;
  prod.it = cons.it = gw;
  f_next<ROWS>(prod, true, nitems, gstride, items, fronts, pan);
  f_next<ROWS>(cons, true, nitems, gstride, items, fronts, pan);
  auto produce = [&](int slot) {
    double *tile = ring + (size_t)slot * ROWS * PITCH;
    if (prod.it >= nitems) {
      if (KIND == 1) cp_async_commit();  // keep one group per iteration so that wait_group counts stages
      return;
    }
    const int nr = min(ROWS, prod.nrows - prod.rs);
    if (KIND == 0) {  // lane r copies row r of the tile with one bulk copy
      const int nc = min(64, prod.stride - prod.cs);  // the stored row is zero-padded up to its stride: copy whole 32-byte groups
      const uint32_t bytes = (uint32_t)nc * 8u;
      if (lane == 0) mbar_expect_tx(bars + slot, bytes * (uint32_t)nr);
      __syncwarp();
      if (lane < nr) bulk_g2s(tile + (size_t)lane * PITCH, prod.base + (int64_t)(prod.rs + lane) * prod.stride + prod.cs, bytes, bars + slot);
    } else {  // every instruction moves one 512-byte row: lane l its 16 bytes at column 2 l (zeros past the stored width / the last row)
      const bool colok = prod.cs + 2 * lane < prod.stride;
      const double *src = prod.base + (int64_t)prod.rs * prod.stride + prod.cs + 2 * lane;
#pragma unroll
      for (int r = 0; r < ROWS; ++r) cp_async16(tile + (size_t)r * PITCH + 2 * lane, (colok && r < nr) ? src + (int64_t)r * prod.stride : prod.base, (colok && r < nr) ? 16 : 0);
      cp_async_commit();
    }
  };
  int pslot = 0, cslot = 0;
  uint32_t cphase = 0;
  for (int q = 0; q < NST - 1; ++q) {  // prologue: fill NST - 1 stages
    produce(pslot);
    pslot = (pslot + 1) % NST;
    if (prod.it < nitems) f_next<ROWS>(prod, false, nitems, gstride, items, fronts, pan);
  }
  double acc[4][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = 0.0;
  double bx[8], by[8];
  while (cons.it < nitems) {
    // keep the ring full: the slot freed by the previous iteration (all lanes are past their reads of it: __syncwarp below)
    produce(pslot);
    pslot = (pslot + 1) % NST;
    if (prod.it < nitems) f_next<ROWS>(prod, false, nitems, gstride, items, fronts, pan);
    if (cons.rs == 0) {  // right-hand-side fragments of this column chunk: B[k = t][n = g] for the even / odd column of each group of 8
      const double *bc = b + (int64_t)g * n + cons.p0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cons.cs + 8 * j + 2 * t;
        bx[j] = (g < mu && c < cons.c1) ? __ldg(bc + c) : 0.0;
        by[j] = (g < mu && c + 1 < cons.c1) ? __ldg(bc + c + 1) : 0.0;
      }
    }
    if (KIND == 0) mbar_wait(bars + cslot, cphase);
    else {
      cp_async_wait<NST - 1>();  // all but the NST - 1 youngest groups have landed: the stage being consumed is complete for this lane
      __syncwarp();              // ... and for every other lane of the warp
    }
    const double *tile = ring + (size_t)cslot * ROWS * PITCH;
#pragma unroll
    for (int part = 0; part < 4 / RG; ++part) {  // which of the item's 4 row groups this tile holds (static accumulator indices)
      if (cons.rs == part * ROWS) {
#pragma unroll
        for (int rg = 0; rg < RG; ++rg) {
          if (cons.rs + 8 * rg < cons.nrows) {
            const double2 *row = reinterpret_cast<const double2 *>(tile + (size_t)(8 * rg + g) * PITCH) + t;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const double2 a = row[4 * j];
              dmma(acc[part * RG + rg], a.x, bx[j]);
              dmma(acc[part * RG + rg], a.y, by[j]);
            }
          }
        }
      }
    }
    __syncwarp();  // every lane is done reading this slot before it is refilled
    if (KIND == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cslot = (cslot + 1) % NST;
    cphase ^= (cslot == 0);
    const bool last = cons.cs + 64 >= cons.c1 && cons.rs + ROWS >= cons.nrows;
    if (last) {  // publish the item: D[row = 8 rg + g][rhs = 2 t, 2 t + 1]
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) {
        const int r = 8 * rg + g;
        if (r < cons.nrows) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = 2 * t + h;
            if (m < mu) {
              if (cons.pivot) atomicAdd(&y[(int64_t)m * n + cons.p0 + RB * cons.rblk + r], acc[rg][h]);
              else atomicAdd(&b[(int64_t)m * n + rowidx[cons.rptr + RB * (cons.rblk - cons.nb1) + r]], -acc[rg][h]);
            }
          }
        }
        acc[rg][0] = acc[rg][1] = 0.0;
      }
    }
    f_next<ROWS>(cons, false, nitems, gstride, items, fronts, pan);
  }
  if (KIND == 1) cp_async_wait<0>();
}

struct BStage {  // backward: tile = rows [rr, rr + ROWS) x columns [cc, cc + 64) of one item
  int64_t it;
  int cc, rr;
  const double *P, *Pu;
  int s1, ldp, p0, r0, r1, cend;
  int64_t rptr;
};
__device__ __forceinline__ void b_decode(BStage &s, const BwdItem *items, const Front *fronts, const double *pan) {
  const BwdItem w = items[s.it];
  const Front f = fronts[w.front];
  s.s1 = f.s1;
  s.ldp = hb_ldp(f.s1);
  s.P = pan + f.poff;
  s.Pu = s.P + hb_upd_off(f.s1);
  s.p0 = f.p0;
  s.rptr = f.rptr;
  s.r0 = w.r0;
  s.r1 = w.r0 + w.nr;
  s.cend = min(w.c0 + BCH, s.ldp);
  s.cc = w.c0;
  s.rr = w.r0;
}
template <int ROWS>
__device__ __forceinline__ void b_next(BStage &s, bool first, int64_t nitems, int64_t stride_items, const BwdItem *items, const Front *fronts, const double *pan) {
  if (!first) {
    s.rr += ROWS;
    if (s.rr < s.r1) return;
    s.rr = s.r0;
    s.cc += 64;
    if (s.cc < s.cend) return;
    s.it += stride_items;
  }
  while (s.it < nitems) {
    b_decode(s, items, fronts, pan);
    if (s.cc < s.cend && s.rr < s.r1) return;
    s.it += stride_items;
  }
}
// stored row r of a panel: address of its first column and its stored width (pivot rows live in the trapezoid)
__device__ __forceinline__ const double *panel_row(const BStage &s, int r, int &width) {
  if (r < s.s1) {
    const int k = r / RB;
    width = hb_wblk(s.s1, k);
    return s.P + hb_blk_off(k) + (int64_t)(r - k * RB) * width;
  }
  width = s.ldp;
  return s.Pu + (int64_t)(r - s.s1) * s.ldp;
}

template <int KIND>
__global__ void __launch_bounds__(Cfg<KIND>::WARPS * 32, 1) k_bwd_mma(const BwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts, const int *__restrict__ rowidx,
                                                                     const double *__restrict__ pan, const double *__restrict__ y, double *x, int n, int mu) {
  typedef Cfg<KIND> C;
  constexpr int ROWS = C::ROWS, NST = C::NST, WARPS = C::WARPS, PITCH = C::PITCH_B, KS = ROWS / 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double *ring = reinterpret_cast<double *>(smem_raw) + (size_t)warp * NST * ROWS * PITCH;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)WARPS * NST * ROWS * PITCH * sizeof(double)) + warp * NST;
  if (KIND == 0) {
    for (int i = lane; i < NST * ROWS * PITCH; i += 32) ring[i] = 0.0;
    if (lane == 0)
      for (int q = 0; q < NST; ++q) mbar_init(bars + q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, gstride = (int64_t)gridDim.x * WARPS;
  BStage prod, cons;
  prod.it = cons.it = gw;
  b_next<ROWS>(prod, true, nitems, gstride, items, fronts, pan);
  b_next<ROWS>(cons, true, nitems, gstride, items, fronts, pan);
  auto produce = [&](int slot) {
    double *tile = ring + (size_t)slot * ROWS * PITCH;
    if (prod.it >= nitems) {
      if (KIND == 1) cp_async_commit();
      return;
    }
    if (KIND == 0) {  // lane r: row rr + r of the panel by one bulk copy; columns it does not store are zero-filled by hand
      const int r = prod.rr + lane;
      const double *src = prod.P;
      int width = 0;
      if (r < prod.r1) src = panel_row(prod, r, width);
      const int nc = max(0, min(64, width - prod.cc));
      double *dst = tile + (size_t)lane * PITCH;
      for (int c = nc; c < 64; ++c) dst[c] = 0.0;
      unsigned total = 8u * (unsigned)nc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
      if (lane == 0) mbar_expect_tx(bars + slot, total);
      __syncwarp();
      if (nc > 0) bulk_g2s(dst, src + prod.cc, 8u * (unsigned)nc, bars + slot);
    } else {
#pragma unroll
      for (int q = 0; q < ROWS; ++q) {
        const int r = prod.rr + q;
        int width = 0;
        const double *src = prod.P;
        if (r < prod.r1) src = panel_row(prod, r, width);
        const bool ok = prod.cc + 2 * lane < width;  // structural zeros / rows past the item read as zeros
        cp_async16(tile + (size_t)q * PITCH + 2 * lane, ok ? src + prod.cc + 2 * lane : prod.P, ok ? 16 : 0);
      }
      cp_async_commit();
    }
  };
  int pslot = 0, cslot = 0;
  uint32_t cphase = 0;
  for (int q = 0; q < NST - 1; ++q) {
    produce(pslot);
    pslot = (pslot + 1) % NST;
    if (prod.it < nitems) b_next<ROWS>(prod, false, nitems, gstride, items, fronts, pan);
  }
  double accE[4][2], accO[4][2];
#pragma unroll
  for (int q = 0; q < 4; ++q) accE[q][0] = accE[q][1] = accO[q][0] = accO[q][1] = 0.0;
  while (cons.it < nitems) {
    produce(pslot);
    pslot = (pslot + 1) % NST;
    if (prod.it < nitems) b_next<ROWS>(prod, false, nitems, gstride, items, fronts, pan);
    // multipliers of the rows of this tile: B[k = t][n = g] = u[row 4 ks + t][rhs g]
    double ub[KS];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int r = cons.rr + 4 * ks + t;
      double v = 0.0;
      if (g < mu && r < cons.r1) v = (r < cons.s1) ? y[(int64_t)g * n + cons.p0 + r] : -x[(int64_t)g * n + rowidx[cons.rptr + r - cons.s1]];
      ub[ks] = v;
    }
    if (KIND == 0) mbar_wait(bars + cslot, cphase);
    else cp_async_wait<NST - 1>();
    __syncwarp();  // other lanes' copies / zero tails are visible
    const double *tile = ring + (size_t)cslot * ROWS * PITCH;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const double2 *row = reinterpret_cast<const double2 *>(tile + (size_t)(4 * ks + t) * PITCH) + g;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const double2 a = row[8 * jj];
        dmma(accE[jj], a.x, ub[ks]);
        dmma(accO[jj], a.y, ub[ks]);
      }
    }
    __syncwarp();
    if (KIND == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cslot = (cslot + 1) % NST;
    cphase ^= (cslot == 0);
    const bool last = cons.rr + ROWS >= cons.r1;
    if (last) {  // all rows of this 64-column chunk are in: D[col = 16 jj + 2 g (+ 1)][rhs = 2 t, 2 t + 1]
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int c = cons.cc + 16 * jj + 2 * g;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = 2 * t + h;
          if (m < mu) {
            if (c < cons.s1) atomicAdd(&x[(int64_t)m * n + cons.p0 + c], accE[jj][h]);
            if (c + 1 < cons.s1) atomicAdd(&x[(int64_t)m * n + cons.p0 + c + 1], accO[jj][h]);
          }
        }
        accE[jj][0] = accE[jj][1] = accO[jj][0] = accO[jj][1] = 0.0;
      }
    }
    b_next<ROWS>(cons, false, nitems, gstride, items, fronts, pan);
  }
  if (KIND == 1) cp_async_wait<0>();
}
}  // namespace mma


// ---- register-staged tensor-pipe sweeps (the shipped block kernels): same DMMA formulation as above, but the A fragments come
// straight from HBM -- lane (g, t) issues 128-bit streaming loads of [row g][columns 8 j + 2 t, + 1] (forward) or
// [row 4 ks + t][columns 16 jj + 2 g, + 1] (backward), 8 in flight per lane, i.e. the access pattern of the scalar kernels with the
// FMA / shuffle-reduction work moved to the tensor pipe.  One warp = one work item, 8 warps per CTA.
template <int OCC>
__global__ void __launch_bounds__(256, OCC) k_fwd_dmma(const FwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts, const int *__restrict__ rowidx,
                                                       const double *__restrict__ pan, double *b, double *y, int n, int mu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  const FwdItem w = items[it];
  const Front f = fronts[w.front];
  const int s1 = f.s1, nb1 = (s1 + RB - 1) / RB;
  const double *base;
  int nrows, stride, cmax;
  const bool pivot = w.rblk < nb1;
  if (pivot) {
    stride = hb_wblk(s1, w.rblk);
    base = pan + f.poff + hb_blk_off(w.rblk);
    nrows = min(RB, s1 - RB * w.rblk);
    cmax = min(s1, RB * (w.rblk + 1));
  } else {
    const int k2 = w.rblk - nb1;
    stride = hb_ldp(s1);
    base = pan + f.poff + hb_upd_off(s1) + (int64_t)k2 * RB * stride;
    nrows = min(RB, f.s2 - RB * k2);
    cmax = s1;
  }
  const int c1 = min(cmax, w.c0 + w.cw);
  if (c1 <= w.c0) return;
  const double *bc = b + (int64_t)g * n + f.p0;              // right-hand side g (B fragment: n = g), L1-resident chunk
  // ncu (profiles/r02_ncu_dmma4_m96.csv) showed the first version of this kernel latency-bound: 16 warps per SM, 8 loads each = 64 KB
  // in flight per SM (warps stalled on long scoreboard 19 : 1 per issue, DRAM at 55-65 %).  Registers are the budget for bytes in
  // flight, so the chunk is narrow (32 columns: 8 B-fragment registers) and covers all 4 row groups: 16 x 128-bit loads per lane.
  double acc[4][2];
#pragma unroll
  for (int rg = 0; rg < 4; ++rg) acc[rg][0] = acc[rg][1] = 0.0;
  const double2 *row0 = reinterpret_cast<const double2 *>(base + (int64_t)g * stride) + t;
  const int64_t rgstep = (int64_t)8 * stride / 2;          // 8 rows further, in 128-bit units
  for (int cs = w.c0; cs < c1; cs += 32) {
    double2 a[4][4];
#pragma unroll
    for (int rg = 0; rg < 4; ++rg)
#pragma unroll
      for (int u = 0; u < 4; ++u)
        a[rg][u] = (8 * rg + g < nrows && cs + 8 * u + 2 * t < stride) ? ldg_stream(row0 + rg * rgstep + cs / 2 + 4 * u) : make_double2(0.0, 0.0);  // rows are zero-padded to an even width
    double bx[4], by[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = cs + 8 * u + 2 * t;
      bx[u] = (g < mu && c < c1) ? __ldg(bc + c) : 0.0;
      by[u] = (g < mu && c + 1 < c1) ? __ldg(bc + c + 1) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) mma::dmma(acc[rg], a[rg][u].x, bx[u]);  // consecutive MMAs go to different accumulators
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int rg = 0; rg < 4; ++rg) mma::dmma(acc[rg], a[rg][u].y, by[u]);
  }
#pragma unroll
  for (int rg = 0; rg < 4; ++rg) {
    const int r = 8 * rg + g;
    if (r < nrows) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = 2 * t + h;
        if (m < mu) {
          if (pivot) atomicAdd(&y[(int64_t)m * n + f.p0 + RB * w.rblk + r], acc[rg][h]);
          else atomicAdd(&b[(int64_t)m * n + rowidx[f.rptr + RB * (w.rblk - nb1) + r]], -acc[rg][h]);
        }
      }
    }
  }
}

template <int OCC>
__global__ void __launch_bounds__(256, OCC) k_bwd_dmma(const BwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts, const int *__restrict__ rowidx,
                                                       const double *__restrict__ pan, const double *__restrict__ y, double *x, int n, int mu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  const BwdItem w = items[it];
  const Front f = fronts[w.front];
  const int s1 = f.s1, ldp = hb_ldp(s1);
  const double *P = pan + f.poff, *Pu = P + hb_upd_off(s1);
  const int cend = min(w.c0 + BCH, ldp), r1 = w.r0 + w.nr;
  for (int cc = w.c0; cc < cend; cc += 64) {               // 64-column sub-chunks: 4 groups of 16 columns = 8 accumulator fragments
    double accE[4][2], accO[4][2];
#pragma unroll
    for (int q = 0; q < 4; ++q) accE[q][0] = accE[q][1] = accO[q][0] = accO[q][1] = 0.0;
    // multipliers u[row][rhs g] of a trip (B fragments), fetched one trip ahead: the struct rows need two dependent loads (index, value)
    auto multipliers = [&](int rr, double (&ub)[4]) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = rr + 4 * ks + t;
        double v = 0.0;
        if (g < mu && r < r1) v = (r < s1) ? y[(int64_t)g * n + f.p0 + r] : -x[(int64_t)g * n + rowidx[f.rptr + r - s1]];
        ub[ks] = v;
      }
    };
    double ub[4], ubn[4];
    multipliers(w.r0, ub);
    for (int rr = w.r0; rr < r1; rr += 16) {                 // four k-steps (16 rows) per trip: 16 loads in flight per lane
      double2 a[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = rr + 4 * ks + t;
        const double *rowp = P;
        int width = 0;
        if (r < r1) {
          if (r < s1) {
            const int k = r / RB;
            width = hb_wblk(s1, k);
            rowp = P + hb_blk_off(k) + (int64_t)(r - k * RB) * width;
          } else {
            width = ldp;
            rowp = Pu + (int64_t)(r - s1) * ldp;
          }
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int c = cc + 16 * jj + 2 * g;
          a[ks][jj] = (c < width) ? ldg_stream(reinterpret_cast<const double2 *>(rowp + c)) : make_double2(0.0, 0.0);
        }
      }
      multipliers(rr + 16, ubn);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          mma::dmma(accE[jj], a[ks][jj].x, ub[ks]);
          mma::dmma(accO[jj], a[ks][jj].y, ub[ks]);
        }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) ub[ks] = ubn[ks];
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int c = cc + 16 * jj + 2 * g;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = 2 * t + h;
        if (m < mu) {
          if (c < s1) atomicAdd(&x[(int64_t)m * n + f.p0 + c], accE[jj][h]);
          if (c + 1 < s1) atomicAdd(&x[(int64_t)m * n + f.p0 + c + 1], accO[jj][h]);
        }
      }
    }
  }
}

// sweeps for 1 .. 8 right-hand sides on the tensor pipe (one persistent launch per level and sweep)
// HPDDM_B200_MMA_VARIANT = reg (A fragments by 128-bit loads from HBM) | ring (per-warp shared-memory ring filled by 16-byte cp.async,
// persistent warps) | tma (the same ring filled by one cp.async.bulk per tile row: measured issue-bound on 512-byte copies) -- A/B table
// in profiles/README.md
static int launch_levels_mma(Sub *s, cudaStream_t st, int mu) {
  static const char *var = getenv("HPDDM_B200_MMA_VARIANT");
  static const int kind = !var ? 2 : (!strcmp(var, "tma") ? 0 : (!strcmp(var, "ring") ? 1 : 2));
  static const int occ = getenv("HPDDM_B200_MMA_OCC") ? atoi(getenv("HPDDM_B200_MMA_OCC")) : 2;
  DeviceFactor &D = s->fac;
  const Symbolic &S = s->sym;
  const int n = S.n;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  auto grid = [&](int64_t ni, int warps) { return (unsigned)std::min<int64_t>(sms, (ni + warps - 1) / warps); };
  for (int l = 0; l < S.nlevels; ++l) {
    const int64_t i0 = S.fwd_ptr[l], ni = S.fwd_ptr[l + 1] - i0;
    if (ni <= 0) continue;
    if (kind == 0) mma::k_fwd_mma<0><<<grid(ni, 4), 128, mma::Cfg<0>::SMEM_F, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y, n, mu);
    else if (kind == 1) mma::k_fwd_mma<1><<<grid(ni, 8), 256, mma::Cfg<1>::SMEM_F, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y, n, mu);
    else if (occ == 3) k_fwd_dmma<3><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y, n, mu);
    else k_fwd_dmma<2><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y, n, mu);
  }
  for (int l = S.nlevels - 1; l >= 0; --l) {
    const int64_t i0 = S.bwd_ptr[l], ni = S.bwd_ptr[l + 1] - i0;
    if (ni <= 0) continue;
    if (kind == 0) mma::k_bwd_mma<0><<<grid(ni, 4), 128, mma::Cfg<0>::SMEM_B, st>>>(D.bwd + i0, ni, D.fronts, D.rowidx, D.panU, D.y, D.x, n, mu);
    else if (kind == 1) mma::k_bwd_mma<1><<<grid(ni, 8), 256, mma::Cfg<1>::SMEM_B, st>>>(D.bwd + i0, ni, D.fronts, D.rowidx, D.panU, D.y, D.x, n, mu);
    else if (occ == 3) k_bwd_dmma<3><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.bwd + i0, ni, D.fronts, D.rowidx, D.panU, D.y, D.x, n, mu);
    else k_bwd_dmma<2><<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.bwd + i0, ni, D.fronts, D.rowidx, D.panU, D.y, D.x, n, mu);
  }
  HB_CUDA(cudaGetLastError());
  return 0;
}
// HPDDM_B200_MMA: "0" = register-tiled kernels only; "1" (default) = tensor-pipe kernels for 3 .. 8 right-hand sides;
// "2" = also for 2; "all" = also for a single right-hand side (A/B measurements, profiles/README.md)
static int mma_min_mu() {
  static const int v = [] {
    const char *e = getenv("HPDDM_B200_MMA");
    if (!e) return 3;
    if (!strcmp(e, "0")) return 1000;
    if (!strcmp(e, "2")) return 2;
    if (!strcmp(e, "all")) return 1;
    return 3;
  }();
  return v;
}
int sptrsv_prepare(Sub *s) {  // once per factorisation, outside any stream capture: opt-in shared-memory sizes are per device
  HB_CUDA(cudaFuncSetAttribute(mma::k_fwd_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma::Cfg<0>::SMEM_F));
  HB_CUDA(cudaFuncSetAttribute(mma::k_bwd_mma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma::Cfg<0>::SMEM_B));
  HB_CUDA(cudaFuncSetAttribute(mma::k_fwd_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma::Cfg<1>::SMEM_F));
  HB_CUDA(cudaFuncSetAttribute(mma::k_bwd_mma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mma::Cfg<1>::SMEM_B));
  HB_CUDA(cudaFuncSetAttribute(k_fwd<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * 8 * FCH * sizeof(K))));
  HB_CUDA(cudaFuncSetAttribute(k_fwd<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * 8 * FCH * sizeof(K))));
  return persistent_tables(s);
}
int sptrsv_max_block() { return mma_min_mu() <= 8 ? 8 : 4; }
#else
int sptrsv_prepare(Sub *s) {
  HB_CUDA(cudaFuncSetAttribute(k_fwd<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * 8 * FCH * sizeof(K))));
  HB_CUDA(cudaFuncSetAttribute(k_fwd<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)2 * 8 * FCH * sizeof(K))));
  return persistent_tables(s);
}
int sptrsv_max_block() { return 4; }
#endif

// x = A^{-1} b for mu in {1, 2, 4} columns at once (natural ordering in/out, device pointers,
// column stride n).  The 2 * nlevels sweep launches are replayed from a CUDA graph captured on
// first use (their arguments never change); only the two permutation kernels see the caller's
// pointers.
int sptrsv_group(int left) {
#ifndef HB_COMPLEX
  if (left >= mma_min_mu()) return std::min(left, 8);
#endif
  return left >= 4 ? 4 : (left >= 2 ? 2 : 1);
}

int sptrsv_solve(Sub *s, const K *b, K *x, int mu, const double *scale, bool accumulate) {
  DeviceFactor &D = s->fac;
  if (!D.valid) {
    set_error("solve: no factorisation (call numfact first)");
    return HPDDM_B200_ERR_STATE;
  }
  bool tensor = false;
#ifndef HB_COMPLEX
  tensor = mu >= mma_min_mu() && mu <= 8;
#endif
  if (!tensor && mu != 1 && mu != 2 && mu != 4) {
    set_error("sptrsv_solve: %d right-hand sides in one pass are not supported by the register-tiled kernels (1, 2 or 4)", mu);
    return HPDDM_B200_ERR_ARG;
  }
  const Symbolic &S = s->sym;
  cudaStream_t st = s->ctx->stream;
  const int n = S.n;
  if (n == 0) return 0;
  auto sweeps = [&]() -> int {
#ifndef HB_COMPLEX
    if (tensor) return launch_levels_mma(s, st, mu);
#endif
    if (mu == 1 && persistent_on(S)) return launch_persistent(s, st);
    return mu == 1 ? launch_levels<1>(s, st) : (mu == 2 ? launch_levels<2>(s, st) : launch_levels<4>(s, st));
  };
  static const bool graphs_on = getenv("HPDDM_B200_NO_GRAPH") == nullptr;
  // inside an outer capture (the whole-apply graph of hb_api.cu) the sweep launches are recorded straight into that graph: a graph
  // cannot be launched into a capturing stream
  cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cst);
  const bool use_graph = graphs_on && cst == cudaStreamCaptureStatusNone;
  k_perm_in<<<(unsigned)(((int64_t)n * mu + 255) / 256), 256, 0, st>>>(n, mu, D.perm, b, D.b, D.y, D.x);
  if (use_graph && !D.graph[mu]) {
    cudaGraph_t g = nullptr;
    HB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = sweeps();
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc < 0 || e != cudaSuccess) {
      set_error("CUDA graph capture of the SpTRSV sweeps failed (%s)", cudaGetErrorString(e));
      return HPDDM_B200_ERR_CUDA;
    }
    HB_CUDA(cudaGraphInstantiate(&D.graph[mu], g, 0));
    cudaGraphDestroy(g);
  }
  if (use_graph) HB_CUDA(cudaGraphLaunch(D.graph[mu], st));
  else HB_CHECK(sweeps());
  k_perm_out<<<(unsigned)(((int64_t)n * mu + 255) / 256), 256, 0, st>>>(n, mu, D.perm, D.x, scale, x, accumulate ? 1 : 0);
  s->ctx->launches += 2 + D.sweep_launches;
  HB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hb
