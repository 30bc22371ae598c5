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

__global__ void __launch_bounds__(256) k_fwd(const FwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts,
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
  for (int r = 0; r < nrows; r += 4) {
    const double2 *p0 = reinterpret_cast<const double2 *>(base + (int64_t)r * stride + c0);
    const double2 *p1 = reinterpret_cast<const double2 *>(base + (int64_t)min(r + 1, nrows - 1) * stride + c0);
    const double2 *p2 = reinterpret_cast<const double2 *>(base + (int64_t)min(r + 2, nrows - 1) * stride + c0);
    const double2 *p3 = reinterpret_cast<const double2 *>(base + (int64_t)min(r + 3, nrows - 1) * stride + c0);
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll 2
    for (int j = lane; j < nv; j += 32) {
      const double2 t0 = ldg_stream(p0 + j), t1 = ldg_stream(p1 + j), t2 = ldg_stream(p2 + j), t3 = ldg_stream(p3 + j);
      const double2 bb = bs2[j];
      a0 = fma(t0.x, bb.x, fma(t0.y, bb.y, a0));
      a1 = fma(t1.x, bb.x, fma(t1.y, bb.y, a1));
      a2 = fma(t2.x, bb.x, fma(t2.y, bb.y, a2));
      a3 = fma(t3.x, bb.x, fma(t3.y, bb.y, a3));
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    a3 = warp_sum(a3);
    if (lane < 4 && r + lane < nrows) {
      const double v = lane == 0 ? a0 : (lane == 1 ? a1 : (lane == 2 ? a2 : a3));
      if (pivot) atomicAdd(&y[f.p0 + RB * w.rblk + r + lane], v);
      else atomicAdd(&b[rowidx[f.rptr + RB * (w.rblk - nb1) + r + lane]], -v);
    }
  }
}

__global__ void __launch_bounds__(256) k_bwd(const BwdItem *__restrict__ items, int64_t nitems, const Front *__restrict__ fronts,
                                             const int *__restrict__ rowidx, const double *__restrict__ pan, const double *__restrict__ y, double *x) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t it = (int64_t)blockIdx.x * 8 + warp;
  if (it >= nitems) return;
  const BwdItem w = items[it];
  const Front f = fronts[w.front];
  const int s1 = f.s1, ldp = hb_ldp(s1);
  const double *P = pan + f.poff;
  const double *Pu = P + hb_upd_off(s1);
  double2 acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = make_double2(0.0, 0.0);
  const int cl = w.c0 + 2 * lane;  // this lane's first column
  for (int rb = 0; rb < w.nr; rb += 32) {
    const int r = w.r0 + rb + lane;
    double u = 0.0;
    if (rb + lane < w.nr) u = (r < s1) ? y[f.p0 + r] : -x[rowidx[f.rptr + r - s1]];
    const int nq = min(32, w.nr - rb);
#pragma unroll 4
    for (int q = 0; q < nq; ++q) {
      const double uq = __shfl_sync(0xffffffffu, u, q);
      const int rr = w.r0 + rb + q;
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
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cl + 64 * j;
        if (c < wlim) {
          const double2 t = ldg_stream(reinterpret_cast<const double2 *>(rowp + c));
          acc[j].x = fma(t.x, uq, acc[j].x);
          acc[j].y = fma(t.y, uq, acc[j].y);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cl + 64 * j;
    if (c < s1) atomicAdd(&x[f.p0 + c], acc[j].x);
    if (c + 1 < s1) atomicAdd(&x[f.p0 + c + 1], acc[j].y);
  }
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
