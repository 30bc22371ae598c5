// Vector-level kernels of the RAS apply: partition-of-unity scaling fused into
// the producing kernels, CSR SpMV, tall-skinny deflation products, halo
// pack / deterministic unpack-add, D-weighted dots, replicated coarse solve.
//
// Reference counterparts: Wrapper::diag (include/HPDDM_wrapper.hpp:820-831),
// Wrapper::gthr (314-318), generic csrmm (697-733), the two GEMMs of
// Schwarz::deflation (include/HPDDM_schwarz.hpp:1616,1618), the gather /
// scatter-add loops of Subdomain::exchange (include/HPDDM_subdomain.hpp:118-127),
// the coarse solve of CoarseOperator::callSolver
// (include/HPDDM_coarse_operator_impl.hpp:1706-1720).
#include <cstdlib>
#include <cstring>

#include "hb_internal.h"

namespace hb {

namespace {

__device__ __forceinline__ K warp_sum(K v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += hb_shfl_xor(v, o);
  return v;
}

__global__ void kk_scale(int n, int mu, const double *__restrict__ d, const K *__restrict__ in, K *out) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < (int64_t)n * mu) out[t] = d[t % n] * in[t];
}
__global__ void kk_axpy(int64_t n, double a, const K *__restrict__ x, K *y) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) y[t] = hb_fma(a, x[t], y[t]);
}
__global__ void kk_copy(int64_t n, const K *__restrict__ x, K *y) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) y[t] = x[t];
}
__global__ void kk_fill(int64_t n, K v, K *y) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) y[t] = v;
}

// out[i,c] = (d ? d[i] : 1) * (beta * yin[i,c] + alpha * sum_k a[k] x[ja[k],c])
// CSR "vector" kernel: LPR lanes cooperate on a row (coalesced a/ja reads), LPR chosen from the mean row length
template <int LPR>
__global__ void __launch_bounds__(256) kk_spmv(int n, int mu, const int *__restrict__ ia, const int *__restrict__ ja, const K *__restrict__ a, double alpha,
                                               const K *__restrict__ x, double beta, const K *__restrict__ yin, K *out, const double *__restrict__ d) {
  const int64_t gt = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int i = (int)(gt / LPR), sub = (int)(gt % LPR);
  if (i >= n) return;  // LPR divides the warp size: a whole row group leaves together
  const int k0 = ia[i], k1 = ia[i + 1];
  const double di = d ? d[i] : 1.0;
  for (int c = 0; c < mu; ++c) {
    const K *xc = x + (int64_t)c * n;
    K acc = mk(0.0);
    for (int k = k0 + sub; k < k1; k += LPR) acc = hb_fma(a[k], xc[ja[k]], acc);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc += hb_shfl_down(acc, o, LPR);
    if (sub == 0) {
      K v = alpha * acc;
      if (beta != 0.0) v += beta * yin[i + (int64_t)c * n];
      out[i + (int64_t)c * n] = di * v;
    }
  }
}

// Experimental variant for short rows (stencil matrices, <= 12 entries per row on average): a CTA owns 256 consecutive rows, whose
// entries are one contiguous piece of ja / a -- staged into shared memory with coalesced loads, then one thread per row walks its
// entries there.  Motivated by the cold-cache ncu figure of the thread-per-row kernel (1.5 TB/s at 160^3), but slower than it in the
// live (warm, back-to-back) measurement: kept opt-in for re-measurement.
constexpr int SPMV_CAP = IS_COMPLEX ? 2048 : 3072;  // staged entries per CTA (36 / 40 KB)
__global__ void __launch_bounds__(256) kk_spmv_staged(int n, int mu, const int *__restrict__ ia, const int *__restrict__ ja, const K *__restrict__ a, double alpha,
                                                      const K *__restrict__ x, double beta, const K *__restrict__ yin, K *out, const double *__restrict__ d) {
  __shared__ __align__(16) K as[SPMV_CAP];
  __shared__ int js[SPMV_CAP];
  const int r0 = blockIdx.x * 256, r1 = min(n, r0 + 256);
  const int kb = ia[r0], ke = ia[r1];
  const bool staged = ke - kb <= SPMV_CAP;
  if (staged)
    for (int k = kb + threadIdx.x; k < ke; k += 256) {
      as[k - kb] = a[k];
      js[k - kb] = ja[k];
    }
  __syncthreads();
  const int i = r0 + threadIdx.x;
  if (i >= r1) return;
  const int k0 = ia[i], k1 = ia[i + 1];
  const double di = d ? d[i] : 1.0;
  for (int c = 0; c < mu; ++c) {
    const K *xc = x + (int64_t)c * n;
    K acc = mk(0.0);
    if (staged)
      for (int k = k0; k < k1; ++k) acc = hb_fma(as[k - kb], xc[js[k - kb]], acc);
    else
      for (int k = k0; k < k1; ++k) acc = hb_fma(a[k], xc[ja[k]], acc);
    K v = alpha * acc;
    if (beta != 0.0) v += beta * yin[i + (int64_t)c * n];
    out[i + (int64_t)c * n] = di * v;
  }
}

// T[k + ldT*c] += sum_i conj(Z[i + k*n]) * d[i] * x[i + c*n]  (Z^H: Wrapper<K>::transc, schwarz.hpp:1616) ; 1024 rows per CTA, Z read once per column group.
// Vectors are processed in chunks of KC: KC * 4 independent 8-byte loads per thread are in flight, the
// KC * MB partial sums are reduced once per chunk (shuffles + one shared-memory pass) and published with
// one atomic per (CTA, vector, column).
// ldz = stride between the vectors of Z (n for a contiguous n x nu block; mu * n for one column of a block Krylov basis)
template <int MB>
__global__ void __launch_bounds__(256) kk_zt(int n, int nu, int c0, const K *__restrict__ Z, int64_t ldz, const double *__restrict__ d, const K *__restrict__ x,
                                             K *T, int ldT) {
  constexpr int KCR = 8 / MB >= 2 ? 8 / MB : 2;     // 8, 4, 2 vectors per chunk for real scalars
  constexpr int KC = KCR / KD >= 1 ? KCR / KD : 1;  // complex: half as many (same bytes in flight)
  __shared__ K red[8][KC * MB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int base = blockIdx.x * 1024;
  K w[4][MB];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = base + tid + 256 * q;
#pragma unroll
    for (int m = 0; m < MB; ++m) w[q][m] = (i < n) ? d[i] * x[i + (int64_t)(c0 + m) * n] : mk(0.0);
  }
  for (int k0 = 0; k0 < nu; k0 += KC) {
    K z[KC][4];
#pragma unroll
    for (int kk = 0; kk < KC; ++kk)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = base + tid + 256 * q;
        z[kk][q] = (k0 + kk < nu && i < n) ? hb_conj(Z[i + (int64_t)(k0 + kk) * ldz]) : mk(0.0);
      }
    K acc[KC][MB];
#pragma unroll
    for (int kk = 0; kk < KC; ++kk)
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        K a = mk(0.0);
#pragma unroll
        for (int q = 0; q < 4; ++q) a = hb_fma(z[kk][q], w[q][m], a);
        acc[kk][m] = warp_sum(a);
      }
    if (lane == 0) {
#pragma unroll
      for (int kk = 0; kk < KC; ++kk)
#pragma unroll
        for (int m = 0; m < MB; ++m) red[warp][kk * MB + m] = acc[kk][m];
    }
    __syncthreads();
    if (tid < KC * MB) {
      const int kk = tid / MB, m = tid % MB;
      if (k0 + kk < nu) {
        K sum = mk(0.0);
#pragma unroll
        for (int q = 0; q < 8; ++q) sum += red[q][tid];
        hb_atomic_add(&T[k0 + kk + (int64_t)ldT * (c0 + m)], sum);
      }
    }
    __syncthreads();
  }
}

// out[i + c*n] = d[i] * sum_k Z[i + k*n] * Y[k + ldY*c]
template <int MB>
__global__ void __launch_bounds__(256) kk_zexp(int n, int nu, int c0, const K *__restrict__ Z, const double *__restrict__ d, const K *__restrict__ Y,
                                               int ldY, K *out) {
  extern __shared__ __align__(16) unsigned char ys_raw[];
  K *ys = reinterpret_cast<K *>(ys_raw);  // nu * MB
  for (int t = threadIdx.x; t < nu * MB; t += blockDim.x) ys[t] = Y[(t % nu) + (int64_t)ldY * (c0 + t / nu)];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  K acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = mk(0.0);
  for (int k = 0; k < nu; ++k) {
    const K z = Z[i + (int64_t)k * n];
#pragma unroll
    for (int m = 0; m < MB; ++m) acc[m] = hb_fma(z, ys[k + nu * m], acc[m]);
  }
  const double di = d[i];
#pragma unroll
  for (int m = 0; m < MB; ++m) out[i + (int64_t)(c0 + m) * n] = di * acc[m];
}

// send[ebase*mu + c*esize + (e - ebase)] = x[map[e] + c*n]
__global__ void kk_pack(int h, int n, int mu, const int *__restrict__ map, const int *__restrict__ ebase, const int *__restrict__ esize,
                        const K *__restrict__ x, K *send) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)h * mu) return;
  const int e = (int)(t % h), c = (int)(t / h);
  send[(int64_t)ebase[e] * mu + (int64_t)c * esize[e] + (e - ebase[e])] = x[map[e] + (int64_t)c * n];
}
// deterministic unpack-add: one thread per unique target dof, contributions summed in neighbour order
__global__ void kk_unpack(int nuniq, int n, int mu, const int *__restrict__ uidx, const int *__restrict__ useg, const int *__restrict__ upos,
                          const int *__restrict__ ebase, const int *__restrict__ esize, const K *__restrict__ recv, K *x) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)nuniq * mu) return;
  const int u = (int)(t % nuniq), c = (int)(t / nuniq);
  K acc = x[uidx[u] + (int64_t)c * n];
  for (int q = useg[u]; q < useg[u + 1]; ++q) {
    const int e = upos[q];
    acc += recv[(int64_t)ebase[e] * mu + (int64_t)c * esize[e] + (e - ebase[e])];
  }
  x[uidx[u] + (int64_t)c * n] = acc;
}

// res[c] += sum_i d[i] conj(x[i,c]) y[i,c]   (iterative.hpp:503)
__global__ void __launch_bounds__(256) kk_dot(int n, int mu, const double *__restrict__ d, const K *__restrict__ x, const K *__restrict__ y, K *res) {
  __shared__ K red[8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int c = 0; c < mu; ++c) {
    K acc = mk(0.0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + tid; i < n; i += (int64_t)gridDim.x * blockDim.x)
      acc = hb_fma(d[i] * hb_conj(x[i + (int64_t)c * n]), y[i + (int64_t)c * n], acc);
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      K s = mk(0.0);
      for (int q = 0; q < 8; ++q) s += red[q];
      hb_atomic_add(&res[c], s);
    }
    __syncthreads();
  }
}

// Reductions of the Krylov layer that know the boundary-condition rows (bcflag[i] != 0, Subdomain::boundaryConditions).
//  MODE 0: res[c] += sum_i d_i |b'_ic|^2 with b' = b / HPDDM_PEN on flagged rows whose entry exceeds PEN * EPS -- the ||b|| of
//          IterativeMethod::initializeNorm (include/HPDDM_iterative.hpp:455-468);
//  MODE 1..3 (Schwarz::computeResidual, include/HPDDM_schwarz.hpp:761-803; 1 = l2, 2 = l1, 3 = l-infinity), two values per column:
//          res[2c] from f (every entry larger than EPS * PEN divided by PEN), res[2c+1] from t = A x - f with flagged rows skipped.
template <int MODE>
__global__ void __launch_bounds__(256) kk_bcnorm(int n, int mu, const double *__restrict__ d, const unsigned char *__restrict__ bcflag, const K *__restrict__ f,
                                                 const K *__restrict__ t, double *res) {
  __shared__ double red[2][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr double PEN = 1.0e30, EPS = 1.0e-12;
  for (int c = 0; c < mu; ++c) {
    double a0 = 0.0, a1 = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + tid; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const K fv = f[i + (int64_t)c * n];
      const bool flagged = bcflag && bcflag[i];
      if (MODE == 0) {
        const double af = hb_abs(fv);
        a0 += d[i] * ((af > PEN * EPS && flagged) ? hb_norm(fv / PEN) : hb_norm(fv));
      } else {
        const double af = hb_abs(fv) > EPS * PEN ? hb_abs(fv / PEN) : hb_abs(fv);
        const double at = flagged ? 0.0 : hb_abs(t[i + (int64_t)c * n]);
        if (MODE == 1) {
          a0 += d[i] * af * af;
          a1 += d[i] * at * at;
        } else if (MODE == 2) {
          a0 += d[i] * af;
          a1 += d[i] * at;
        } else {
          a0 = fmax(a0, af);
          a1 = fmax(a1, at);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double b0 = __shfl_xor_sync(0xffffffffu, a0, o), b1 = __shfl_xor_sync(0xffffffffu, a1, o);
      a0 = MODE == 3 ? fmax(a0, b0) : a0 + b0;
      a1 = MODE == 3 ? fmax(a1, b1) : a1 + b1;
    }
    if (lane == 0) {
      red[0][warp] = a0;
      red[1][warp] = a1;
    }
    __syncthreads();
    if (tid == 0) {
      double s0 = 0.0, s1 = 0.0;
      for (int q = 0; q < 8; ++q) {
        s0 = MODE == 3 ? fmax(s0, red[0][q]) : s0 + red[0][q];
        s1 = MODE == 3 ? fmax(s1, red[1][q]) : s1 + red[1][q];
      }
      if (MODE == 0) atomicAdd(&res[c], s0);
      else if (MODE == 3) {  // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long *>(&res[2 * c]), (unsigned long long)__double_as_longlong(s0));
        atomicMax(reinterpret_cast<unsigned long long *>(&res[2 * c + 1]), (unsigned long long)__double_as_longlong(s1));
      } else {
        atomicAdd(&res[2 * c], s0);
        atomicAdd(&res[2 * c + 1], s1);
      }
    }
    __syncthreads();
  }
}

// Replicated coarse solve (one CTA): Y = Einv T ; R = T - E Y ; Y += Einv R.
// Coarse vectors use the communication layout [proc][col][row-in-proc]:
//   v(r, c) = buf[rowproc[r] * Lmax * mu + c * Lmax + rowloc[r]]   (process blocks padded to Lmax rows)
__global__ void __launch_bounds__(256) kk_coarse(int Nc, int mu, int Lnu, const int *__restrict__ rowproc, const int *__restrict__ rowloc,
                                                 const K *__restrict__ E, const K *__restrict__ Einv, const K *__restrict__ T, K *Y, K *R) {
  auto at = [&](int r, int c) -> int64_t { return (int64_t)rowproc[r] * Lnu * mu + (int64_t)c * Lnu + rowloc[r]; };
  for (int c = 0; c < mu; ++c) {
    for (int r = threadIdx.x; r < Nc; r += blockDim.x) {
      K acc = mk(0.0);
      for (int k = 0; k < Nc; ++k) acc = hb_fma(Einv[r + (int64_t)k * Nc], T[at(k, c)], acc);
      Y[at(r, c)] = acc;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < Nc; r += blockDim.x) {
      K acc = T[at(r, c)];
      for (int k = 0; k < Nc; ++k) acc = hb_fma(-E[r + (int64_t)k * Nc], Y[at(k, c)], acc);
      R[r] = acc;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < Nc; r += blockDim.x) {
      K acc = mk(0.0);
      for (int k = 0; k < Nc; ++k) acc = hb_fma(Einv[r + (int64_t)k * Nc], R[k], acc);
      Y[at(r, c)] += acc;
    }
    __syncthreads();
  }
}

// Large coarse spaces (N_c > 512): the same three passes as kk_coarse, one launch each over all SMs (thread per row, columns of the
// matrix read coalesced).  PASS 0: Y = Einv T   PASS 1: R = T - E Y   PASS 2: Y += Einv R     (R: N_c x mu, column-major)
template <int PASS>
__global__ void __launch_bounds__(256) kk_coarse_pass(int Nc, int mu, int Lnu, const int *__restrict__ rowproc, const int *__restrict__ rowloc,
                                                      const K *__restrict__ M, const K *__restrict__ T, K *Y, K *R) {
  auto at = [&](int r, int c) -> int64_t { return (int64_t)rowproc[r] * Lnu * mu + (int64_t)c * Lnu + rowloc[r]; };
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Nc) return;
  for (int c = 0; c < mu; ++c) {
    K acc = PASS == 1 ? T[at(r, c)] : mk(0.0);
    if (PASS == 0)
      for (int k = 0; k < Nc; ++k) acc = hb_fma(M[r + (int64_t)k * Nc], T[at(k, c)], acc);
    else if (PASS == 1)
      for (int k = 0; k < Nc; ++k) acc = hb_fma(-M[r + (int64_t)k * Nc], Y[at(k, c)], acc);
    else
      for (int k = 0; k < Nc; ++k) acc = hb_fma(M[r + (int64_t)k * Nc], R[k + (int64_t)c * Nc], acc);
    if (PASS == 0) Y[at(r, c)] = acc;
    else if (PASS == 1) R[r + (int64_t)c * Nc] = acc;
    else Y[at(r, c)] += acc;
  }
}

// v[i] = 0 where |v[i]| < tiny (Schwarz::solveGEVP post-processing, schwarz.hpp:713)
__global__ void kk_flush_tiny(int64_t n, double tiny, K *v) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n && hb_abs(v[t]) < tiny) v[t] = mk(0.0);
}

__global__ void kk_bc(int nbc, int n, int mu, const int *__restrict__ idx, const K *__restrict__ val, const K *__restrict__ b, K *x) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nbc * mu) return;
  const int q = t % nbc, c = t / nbc;
  x[idx[q] + (int64_t)c * n] = b[idx[q] + (int64_t)c * n] / val[q];
}

// w[i] += sign * sum_j V[i + j*ldv] h[j]   (Gram-Schmidt update / Krylov linear combination)
__global__ void __launch_bounds__(256) kk_vupdate(int n, int k, const K *__restrict__ V, int64_t ldv, const K *__restrict__ h, double sign, K *w) {
  extern __shared__ __align__(16) unsigned char hs_raw[];
  K *hs = reinterpret_cast<K *>(hs_raw);
  for (int t = threadIdx.x; t < k; t += blockDim.x) hs[t] = h[t];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  K acc = mk(0.0);
  for (int j = 0; j < k; ++j) acc = hb_fma(V[i + (int64_t)j * ldv], hs[j], acc);
  w[i] = hb_fma(sign, acc, w[i]);
}
// W[i, c] += sign * sum_j V[i + j*n] * H[j + ldh*c]  for MB columns c0 .. c0+MB (block Gram-Schmidt update / block Krylov
// linear combination: gemm("N", "N", n, mu, k) of blockOrthogonalization and addSol, iterative.hpp:519,544,322)
template <int MB>
__global__ void __launch_bounds__(256) kk_vupdate_blk(int n, int k, int c0, const K *__restrict__ V, const K *__restrict__ H, int ldh, double sign, K *W) {
  extern __shared__ __align__(16) unsigned char hb_raw[];
  K *hs = reinterpret_cast<K *>(hb_raw);  // k * MB
  for (int t = threadIdx.x; t < k * MB; t += blockDim.x) hs[t] = H[(t % k) + (int64_t)ldh * (c0 + t / k)];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  K acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = mk(0.0);
  for (int j = 0; j < k; ++j) {
    const K v = V[i + (int64_t)j * n];
#pragma unroll
    for (int m = 0; m < MB; ++m) acc[m] = hb_fma(v, hs[j + k * m], acc[m]);
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) W[i + (int64_t)(c0 + m) * n] = hb_fma(sign, acc[m], W[i + (int64_t)(c0 + m) * n]);
}
// W <- W * R  in place, R = mu x mu upper triangular (column-major, ld = mu; the inverse Cholesky factor of CholQR:
// trsm("R", "U", "N", "N") of IterativeMethod::QR, iterative.hpp:635), mu <= 8
__global__ void __launch_bounds__(256) kk_rmul_upper(int n, int mu, const K *__restrict__ R, K *W) {
  __shared__ K rs[64];
  if ((int)threadIdx.x < mu * mu) rs[threadIdx.x] = R[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  K w[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) w[c] = c < mu ? W[i + (int64_t)c * n] : mk(0.0);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (c < mu) {
      K acc = mk(0.0);
#pragma unroll
      for (int r = 0; r < 8; ++r)
        if (r <= c) acc = hb_fma(w[r], rs[r + mu * c], acc);
      W[i + (int64_t)c * n] = acc;
    }
  }
}
__global__ void kk_scal_copy(int64_t n, double a, const K *__restrict__ x, K *y) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) y[t] = a * x[t];
}

inline unsigned grid1(int64_t n, int bs = 256) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

#define HB_LAUNCH_END(c)       \
  (c)->launches++;             \
  HB_CUDA(cudaGetLastError()); \
  return 0

int k_scale(Ctx *c, int n, int mu, const double *d, const K *in, K *out) {
  if ((int64_t)n * mu == 0) return 0;
  kk_scale<<<grid1((int64_t)n * mu), 256, 0, c->stream>>>(n, mu, d, in, out);
  HB_LAUNCH_END(c);
}
int k_axpy(Ctx *c, int64_t n, double a, const K *x, K *y) {
  if (n == 0) return 0;
  kk_axpy<<<grid1(n), 256, 0, c->stream>>>(n, a, x, y);
  HB_LAUNCH_END(c);
}
int k_copy(Ctx *c, int64_t n, const K *x, K *y) {
  if (n == 0 || x == y) return 0;
  kk_copy<<<grid1(n), 256, 0, c->stream>>>(n, x, y);
  HB_LAUNCH_END(c);
}
int k_fill(Ctx *c, int64_t n, K v, K *y) {
  if (n == 0) return 0;
  kk_fill<<<grid1(n), 256, 0, c->stream>>>(n, v, y);
  HB_LAUNCH_END(c);
}
int k_spmv(Ctx *c, const Sub *s, int mu, double alpha, const K *x, double beta, const K *yin, K *out, const double *d) {
  if (s->n == 0) return 0;
  return k_spmv_raw(c, s->n, s->A.ia[s->n], s->d_ia, s->d_ja, s->d_a, mu, alpha, x, beta, yin, out, d);
}
int k_spmv_raw(Ctx *c, int n, int64_t nnz, const int *ia, const int *ja, const K *a, int mu, double alpha, const K *x, double beta, const K *yin, K *out,
               const double *d) {
  if (n == 0) return 0;
  const double avg = (double)nnz / n;
#define HB_SPMV(L) kk_spmv<L><<<grid1((int64_t)n * L), 256, 0, c->stream>>>(n, mu, ia, ja, a, alpha, x, beta, yin, out, d)
  // short rows (7-point stencils): one thread per row measured 3x faster than 4 lanes per row on B200
  // staged variant: opt-in (HPDDM_B200_SPMV=staged) -- live A/B (profiles/README.md): 12.7 vs 8.6 us at 64^3, 0.28 vs 0.16 ms at 160^3 for
  // SpMV + D-scale: the thread-per-row kernel wins (its strided reads hit L1 / L2; the staged one pays two passes and 36 KB per CTA)
  static const bool staged = getenv("HPDDM_B200_SPMV") && !strcmp(getenv("HPDDM_B200_SPMV"), "staged");
  if (avg <= 12 && staged) kk_spmv_staged<<<grid1(n), 256, 0, c->stream>>>(n, mu, ia, ja, a, alpha, x, beta, yin, out, d);
  else if (avg <= 16) HB_SPMV(1);
  else if (avg <= 40) HB_SPMV(8);
  else if (avg <= 96) HB_SPMV(16);
  else HB_SPMV(32);
#undef HB_SPMV
  HB_LAUNCH_END(c);
}
int k_zt_project(Ctx *c, const Sub *s, int mu, const K *x, K *T, int ldT) { return k_zt_raw(c, s->n, s->nu, s->d_Z, s->d_d, mu, x, T, ldT); }
// T[k + ldT*col] += sum_i Z[i,k] d[i] x[i,col]  on raw arrays (Z: n x nu column-major)
int k_zt_raw(Ctx *c, int n, int nu, const K *Z, const double *d, int mu, const K *x, K *T, int ldT) {
  if (n == 0 || nu == 0) return 0;
  const unsigned g = grid1(n, 1024);
  int c0 = 0;
  while (c0 < mu) {
    const int left = mu - c0;
    if (left >= 4) {
      kk_zt<4><<<g, 256, 0, c->stream>>>(n, nu, c0, Z, n, d, x, T, ldT);
      c0 += 4;
    } else if (left >= 2) {
      kk_zt<2><<<g, 256, 0, c->stream>>>(n, nu, c0, Z, n, d, x, T, ldT);
      c0 += 2;
    } else {
      kk_zt<1><<<g, 256, 0, c->stream>>>(n, nu, c0, Z, n, d, x, T, ldT);
      c0 += 1;
    }
    c->launches++;
  }
  HB_CUDA(cudaGetLastError());
  return 0;
}
int k_z_expand(Ctx *c, const Sub *s, int mu, const K *Y, int ldY, K *out) {
  if (s->n == 0) return 0;
  if (s->nu == 0) return k_fill(c, (int64_t)s->n * mu, mk(0.0), out);
  return k_zexp_raw(c, s->n, s->nu, s->d_Z, s->d_d, mu, Y, ldY, out);
}
// out[i,col] = d[i] * sum_k Z[i,k] Y[k,col]  on raw arrays
int k_zexp_raw(Ctx *c, int n, int nu, const K *Z, const double *d, int mu, const K *Y, int ldY, K *out) {
  if (n == 0 || nu == 0) return 0;
  const unsigned g = grid1(n);
  int c0 = 0;
  while (c0 < mu) {
    const int left = mu - c0;
    if (left >= 4) {
      kk_zexp<4><<<g, 256, nu * 4 * sizeof(K), c->stream>>>(n, nu, c0, Z, d, Y, ldY, out);
      c0 += 4;
    } else if (left >= 2) {
      kk_zexp<2><<<g, 256, nu * 2 * sizeof(K), c->stream>>>(n, nu, c0, Z, d, Y, ldY, out);
      c0 += 2;
    } else {
      kk_zexp<1><<<g, 256, nu * sizeof(K), c->stream>>>(n, nu, c0, Z, d, Y, ldY, out);
      c0 += 1;
    }
    c->launches++;
  }
  HB_CUDA(cudaGetLastError());
  return 0;
}
int k_pack(Ctx *c, const Sub *s, int mu, const K *x, K *send) {
  if (s->h == 0) return 0;
  kk_pack<<<grid1((int64_t)s->h * mu), 256, 0, c->stream>>>(s->h, s->n, mu, s->d_map, s->d_ebase, s->d_esize, x, send);
  HB_LAUNCH_END(c);
}
int k_unpack(Ctx *c, const Sub *s, int mu, K *x) {
  if (s->nuniq == 0) return 0;
  kk_unpack<<<grid1((int64_t)s->nuniq * mu), 256, 0, c->stream>>>(s->nuniq, s->n, mu, s->d_uidx, s->d_useg, s->d_upos, s->d_ebase, s->d_esize, s->d_recv, x);
  HB_LAUNCH_END(c);
}
int k_dot(Ctx *c, const Sub *s, int mu, const K *x, const K *y, K *res) {
  if (s->n == 0) return 0;
  unsigned g = grid1(s->n);
  if (g > 1184) g = 1184;
  kk_dot<<<g, 256, 0, c->stream>>>(s->n, mu, s->d_d, x, y, res);
  HB_LAUNCH_END(c);
}
int k_rhs_norm(Ctx *c, const Sub *s, int mu, const K *b, double *res) {
  if (s->n == 0) return 0;
  unsigned g = grid1(s->n);
  if (g > 1184) g = 1184;
  kk_bcnorm<0><<<g, 256, 0, c->stream>>>(s->n, mu, s->d_d, s->d_bcflag, b, nullptr, res);
  HB_LAUNCH_END(c);
}
int k_residual_norms(Ctx *c, const Sub *s, int mu, int norm, const K *f, const K *t, double *res) {
  if (s->n == 0) return 0;
  unsigned g = grid1(s->n);
  if (g > 1184) g = 1184;
  if (norm == 1) kk_bcnorm<2><<<g, 256, 0, c->stream>>>(s->n, mu, s->d_d, s->d_bcflag, f, t, res);
  else if (norm == 2) kk_bcnorm<3><<<g, 256, 0, c->stream>>>(s->n, mu, s->d_d, s->d_bcflag, f, t, res);
  else kk_bcnorm<1><<<g, 256, 0, c->stream>>>(s->n, mu, s->d_d, s->d_bcflag, f, t, res);
  HB_LAUNCH_END(c);
}
int k_coarse_solve(Ctx *c, int mu) {
  if (c->Nc == 0) return 0;
  if (c->coarse_multipass) {  // d_R holds N_c x mu_cap values (install_coarse / ensure_capacity)
    const unsigned g = grid1(c->Nc);
    kk_coarse_pass<0><<<g, 256, 0, c->stream>>>(c->Nc, mu, c->Lnu, c->d_rowproc, c->d_rowloc, c->d_Einv, c->d_T, c->d_Y, c->d_R);
    kk_coarse_pass<1><<<g, 256, 0, c->stream>>>(c->Nc, mu, c->Lnu, c->d_rowproc, c->d_rowloc, c->d_E, c->d_T, c->d_Y, c->d_R);
    kk_coarse_pass<2><<<g, 256, 0, c->stream>>>(c->Nc, mu, c->Lnu, c->d_rowproc, c->d_rowloc, c->d_Einv, c->d_T, c->d_Y, c->d_R);
    c->launches += 2;
    HB_LAUNCH_END(c);
  }
  kk_coarse<<<1, 256, 0, c->stream>>>(c->Nc, mu, c->Lnu, c->d_rowproc, c->d_rowloc, c->d_E, c->d_Einv, c->d_T, c->d_Y, c->d_R);
  HB_LAUNCH_END(c);
}
int k_vdots(Ctx *c, const Sub *s, int k, const K *V, int64_t ldv, const K *w, K *T) {
  if (s->n == 0 || k == 0) return 0;
  kk_zt<1><<<grid1(s->n, 1024), 256, 0, c->stream>>>(s->n, k, 0, V, ldv, s->d_d, w, T, k);
  HB_LAUNCH_END(c);
}
int k_vupdate(Ctx *c, const Sub *s, int k, const K *V, int64_t ldv, const K *h, double sign, K *w) {
  if (s->n == 0 || k == 0) return 0;
  kk_vupdate<<<grid1(s->n), 256, k * sizeof(K), c->stream>>>(s->n, k, V, ldv, h, sign, w);
  HB_LAUNCH_END(c);
}
int k_vupdate_blk(Ctx *c, int n, int k, int mu, const K *V, const K *H, int ldh, double sign, K *W) {
  if (n == 0 || k == 0 || mu == 0) return 0;
  const unsigned g = grid1(n);
  int c0 = 0;
  while (c0 < mu) {
    const int left = mu - c0;
    if (left >= 4) {
      kk_vupdate_blk<4><<<g, 256, (size_t)k * 4 * sizeof(K), c->stream>>>(n, k, c0, V, H, ldh, sign, W);
      c0 += 4;
    } else if (left >= 2) {
      kk_vupdate_blk<2><<<g, 256, (size_t)k * 2 * sizeof(K), c->stream>>>(n, k, c0, V, H, ldh, sign, W);
      c0 += 2;
    } else {
      kk_vupdate_blk<1><<<g, 256, (size_t)k * sizeof(K), c->stream>>>(n, k, c0, V, H, ldh, sign, W);
      c0 += 1;
    }
    c->launches++;
  }
  HB_CUDA(cudaGetLastError());
  return 0;
}
int k_rmul_upper(Ctx *c, int n, int mu, const K *R, K *W) {
  if (n == 0 || mu == 0) return 0;
  if (mu > 8) {
    set_error("block Krylov methods support at most 8 right-hand sides per block (got %d)", mu);
    return HPDDM_B200_ERR_ARG;
  }
  kk_rmul_upper<<<grid1(n), 256, 0, c->stream>>>(n, mu, R, W);
  HB_LAUNCH_END(c);
}
int k_scal_copy(Ctx *c, int64_t n, double a, const K *x, K *y) {
  if (n == 0) return 0;
  kk_scal_copy<<<grid1(n), 256, 0, c->stream>>>(n, a, x, y);
  HB_LAUNCH_END(c);
}
int k_flush_tiny(Ctx *c, int64_t n, double tiny, K *v) {
  if (n == 0) return 0;
  kk_flush_tiny<<<grid1(n), 256, 0, c->stream>>>(n, tiny, v);
  HB_LAUNCH_END(c);
}
int k_bc(Ctx *c, const Sub *s, int mu, const K *b, K *x) {
  const int nbc = (int)s->bc.size();
  if (nbc == 0) return 0;
  kk_bc<<<grid1((int64_t)nbc * mu), 256, 0, c->stream>>>(nbc, s->n, mu, s->d_bc_idx, s->d_bc_val, b, x);
  HB_LAUNCH_END(c);
}

}  // namespace hb
