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
// columns backward), reduces with shuffles and publishes with FP64 atomics
// (RED.ADD.F64).  Pure HBM streaming: 0.25 flop/byte.
#include "hb_internal.h"

namespace hb {

namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double2 ldg_stream(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// 8 row sums spread over the warp -> lane L (L % 4 == 0) ends up with the total of row
// 4*bit4(L) + 2*bit3(L) + bit2(L): 9 shuffles instead of 40
__device__ __forceinline__ double reduce8(double (&a)[8], int lane) {
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = u16 ? a[i] : a[i + 4], keep = u16 ? a[i + 4] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = u8 ? a[i] : a[i + 2], keep = u8 ? a[i + 2] : a[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const double send = u4 ? a[0] : a[1], keep = u4 ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 2);
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
  return a[0];
}

__global__ void __launch_bounds__(256, 3) k_fwd(const FwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts,
                                                const int *__restrict__ rowidx, const double *__restrict__ pan, double *b, double *y) {
  __shared__ __align__(16) double bs[8][FCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  const FwdItem w = items[it];
  const Front f = fronts[w.front];
  const int s1 = f.s1, nb1 = (s1 + RB - 1) / RB;
  const double *base;
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
  const int c0 = w.c0, c1 = min(cmax, c0 + FCH);
  const int nc = c1 - c0;
  double *mybs = bs[warp];
  for (int c = lane; c < nc; c += 32) mybs[c] = b[f.p0 + c0 + c];
  if ((nc & 1) && lane == 0) mybs[nc] = 0.0;  // panels are zero-padded to even widths
  __syncwarp();
  const int nv = (nc + 1) >> 1;
  const double2 *bs2 = reinterpret_cast<const double2 *>(mybs);
  const int st2 = stride >> 1;  // row stride in double2
  const int rsel = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
  for (int r = 0; r < nrows; r += 8) {
    const double2 *p = reinterpret_cast<const double2 *>(base + (int64_t)r * stride + c0);
    const int nr = min(8, nrows - r);
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.0;
    for (int j = lane; j < nv; j += 32) {
      double2 t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t[q] = (q < nr) ? ldg_stream(p + (int64_t)q * st2 + j) : make_double2(0.0, 0.0);
      const double2 bb = bs2[j];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = fma(t[q].x, bb.x, fma(t[q].y, bb.y, a[q]));
    }
    const double v = reduce8(a, lane);
    if ((lane & 3) == 0 && rsel < nr) {
      if (pivot) atomicAdd(&y[f.p0 + RB * w.rblk + r + rsel], v);
      else atomicAdd(&b[rowidx[f.rptr + RB * (w.rblk - nb1) + r + rsel]], -v);
    }
  }
}

// NJ = number of 64-column slabs of the chunk this lane covers (1, 2 or 4); rows are
// processed in groups of 8/NJ so that 8 128-bit loads are in flight per lane
template <int NJ>
__device__ __forceinline__ void bwd_item(const BwdItem &w, const Front &f, const int *__restrict__ rowidx, const double *__restrict__ pan,
                                         const double *__restrict__ y, double *x, int lane) {
  constexpr int G = 8 / NJ;
  const int s1 = f.s1, ldp = hb_ldp(s1);
  const double *P = pan + f.poff;
  const double *Pu = P + hb_upd_off(s1);
  double2 acc[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) acc[j] = make_double2(0.0, 0.0);
  const int cl = w.c0 + 2 * lane;  // this lane's first column
  for (int rb = 0; rb < w.nr; rb += 32) {
    const int r = w.r0 + rb + lane;
    double u = 0.0;
    if (rb + lane < w.nr) u = (r < s1) ? y[f.p0 + r] : -x[rowidx[f.rptr + r - s1]];
    const int nq = min(32, w.nr - rb);
    for (int q = 0; q < nq; q += G) {
      double2 t[G][NJ];
      double uq[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        uq[g] = __shfl_sync(0xffffffffu, u, (q + g) & 31);
        const int rr = w.r0 + rb + q + g;
        const double *rowp;
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
          const int c = cl + 64 * j;
          t[g][j] = (c < wlim) ? ldg_stream(reinterpret_cast<const double2 *>(rowp + c)) : make_double2(0.0, 0.0);
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          acc[j].x = fma(t[g][j].x, uq[g], acc[j].x);
          acc[j].y = fma(t[g][j].y, uq[g], acc[j].y);
        }
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = cl + 64 * j;
    if (c < s1) atomicAdd(&x[f.p0 + c], acc[j].x);
    if (c + 1 < s1) atomicAdd(&x[f.p0 + c + 1], acc[j].y);
  }
}

__global__ void __launch_bounds__(256, 3) k_bwd(const BwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts,
                                                const int *__restrict__ rowidx, const double *__restrict__ pan, const double *__restrict__ y, double *x) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  const BwdItem w = items[it];
  const Front f = fronts[w.front];
  const int width = min(BCH, hb_ldp(f.s1) - w.c0);
  if (width <= 64) bwd_item<1>(w, f, rowidx, pan, y, x, lane);
  else if (width <= 128) bwd_item<2>(w, f, rowidx, pan, y, x, lane);
  else bwd_item<4>(w, f, rowidx, pan, y, x, lane);
}

__global__ void k_perm_in(int n, const int *__restrict__ perm, const double *__restrict__ in, double *b, double *y, double *x) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    b[i] = in[perm[i]];
    y[i] = 0.0;
    x[i] = 0.0;
  }
}
__global__ void k_perm_out(int n, const int *__restrict__ perm, const double *__restrict__ x, const double *__restrict__ d, double *out, int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int p = perm[i];
    double v = x[i];
    if (d) v *= d[p];
    out[p] = accumulate ? out[p] + v : v;
  }
}

}  // namespace

int sptrsv_solve(Sub *s, const double *b, double *x, const double *scale, bool accumulate) {
  DeviceFactor &D = s->fac;
  if (!D.valid) {
    set_error("solve: no factorisation (call numfact first)");
    return HPDDM_B200_ERR_STATE;
  }
  const Symbolic &S = s->sym;
  cudaStream_t st = s->ctx->stream;
  const int n = S.n;
  if (n == 0) return 0;
  k_perm_in<<<(n + 255) / 256, 256, 0, st>>>(n, D.perm, b, D.b, D.y, D.x);
  s->ctx->launches++;
  for (int l = 0; l < S.nlevels; ++l) {
    const int64_t i0 = S.fwd_ptr[l], ni = S.fwd_ptr[l + 1] - i0;
    if (ni <= 0) continue;
    k_fwd<<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.fwd + i0, ni, D.fronts, D.rowidx, D.panL, D.b, D.y);
    s->ctx->launches++;
  }
  for (int l = S.nlevels - 1; l >= 0; --l) {
    const int64_t i0 = S.bwd_ptr[l], ni = S.bwd_ptr[l + 1] - i0;
    if (ni <= 0) continue;
    k_bwd<<<(unsigned)((ni + 7) / 8), 256, 0, st>>>(D.bwd + i0, ni, D.fronts, D.rowidx, D.panU, D.y, D.x);
    s->ctx->launches++;
  }
  k_perm_out<<<(n + 255) / 256, 256, 0, st>>>(n, D.perm, D.x, scale, x, accumulate ? 1 : 0);
  s->ctx->launches++;
  HB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hb
