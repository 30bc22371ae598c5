// GPU multifrontal numeric factorisation producing the *inverted-diagonal-block*
// supernodal panels that the SpTRSV kernels stream (hb_solve.cu).
//
// Replaces Schwarz::callNumfact -> SUBDOMAIN<K>::numfact (reference:
// include/HPDDM_schwarz.hpp:337-368; the arithmetic itself is third-party there:
// include/HPDDM_SuiteSparse.hpp:264-371, include/HPDDM_MUMPS.hpp:228-291).
// Setup phase only -- off the measured hot path; dense front kernels of large
// fronts go through cuSOLVER/cuBLAS, small fronts through a batched CTA-per-front
// kernel.
//
// Per front (pivots s1, border s2, frontal matrix F column-major, ld = s1+s2):
//   SPD :  F11 = L L^T ; L21 = F21 L^-T ; S = F22 - L21 L21^T
//          panel = [ W ; M ] , W = L^-1 , M = L21 W           (used by both sweeps)
//   LU  :  F11 = L U (no pivoting) ; L21 = F21 U^-1 ; U12 = L^-1 F12 ; S = F22 - L21 U12
//          panL = [ L^-1 ; L21 L^-1 ] , panU = [ U^-T ; (U^-1 U12)^T ]
// so that  forward : y1 = W b1 ,  b2 -= M b1          (one row-major GEMV)
//          backward: x1 = panU^T [ y1 ; -x2 ]         (one transposed GEMV)
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <chrono>
#include <cmath>

#include "hb_internal.h"

namespace hb {
const char *get_error();
static const char *get_error_msg() { return get_error(); }

namespace {

static int SMALL = 160;   // fronts with s1+s2 <= SMALL are factored by the batched kernel

__global__ void k_assemble(int64_t ne, const int *__restrict__ efront, const int *__restrict__ erow, const int *__restrict__ ecol,
                           const K *__restrict__ eval, const int64_t *__restrict__ foff, const int *__restrict__ fs, K *F) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= ne) return;
  int f = efront[e];
  hb_atomic_add(&F[foff[f] + erow[e] + (int64_t)ecol[e] * fs[f]], eval[e]);
}

struct EATask {
  int child;
  int j0;  // first column of the child's update matrix handled by this CTA
};
// extend-add: F_parent[rel[i], rel[j]] += U_child[i, j]
__global__ void k_extend_add(const EATask *__restrict__ tasks, const Front *__restrict__ fronts, const int *__restrict__ rel,
                             const int64_t *__restrict__ foff, const K *__restrict__ Fchild, K *Fpar, int lower_only) {
  EATask t = tasks[blockIdx.x];
  const Front c = fronts[t.child];
  const Front p = fronts[c.parent];
  const int lc = c.s1 + c.s2, lp = p.s1 + p.s2;
  const K *U = Fchild + foff[t.child] + c.s1 + (int64_t)c.s1 * lc;
  K *P = Fpar + foff[c.parent];
  const int *r = rel + c.rptr;
  const int j1 = min(t.j0 + 16, c.s2);
  for (int j = t.j0; j < j1; ++j) {
    const int64_t pj = (int64_t)r[j] * lp;
    for (int i = (lower_only ? j : 0) + threadIdx.x; i < c.s2; i += blockDim.x) hb_atomic_add(&P[r[i] + pj], U[i + (int64_t)j * lc]);
  }
}

// ------------------------------------------------------------------ small fronts
// one CTA per front; F in global memory (L1/L2 resident for these sizes)
__global__ void __launch_bounds__(256) k_factor_small(const int *__restrict__ list, const Front *__restrict__ fronts, const int64_t *__restrict__ foff,
                                                        K *Fbuf, K *panL, K *panU, int symmetric, int *info) {
  const int f = list[blockIdx.x];
  const Front fr = fronts[f];
  const int s1 = fr.s1, s2 = fr.s2, s = s1 + s2;
  K *F = Fbuf + foff[f];
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  __syncthreads();
  for (int k = 0; k < s1; ++k) {
    const K piv = F[k + (int64_t)k * s];
    if (symmetric ? !(hb_real(piv) > 0.0) : !(hb_abs(piv) > 0.0)) {
      if (tid == 0) bad = 1;
    }
    __syncthreads();
    if (bad) break;
    if (symmetric) {  // real scalars only (numfact_device never selects LL^T for complex)
      const double sq = sqrt(hb_real(piv));
      for (int i = k + tid; i < s; i += nt) F[i + (int64_t)k * s] = (i == k) ? mk(sq) : F[i + (int64_t)k * s] / sq;
      __syncthreads();
      const int m = s - k - 1;
      for (int idx = tid; idx < m * m; idx += nt) {
        const int j = k + 1 + idx / m, i = k + 1 + idx % m;
        if (i >= j) F[i + (int64_t)j * s] -= F[i + (int64_t)k * s] * F[j + (int64_t)k * s];
      }
    } else {
      for (int i = k + 1 + tid; i < s; i += nt) F[i + (int64_t)k * s] = F[i + (int64_t)k * s] / piv;
      __syncthreads();
      const int m = s - k - 1;
      for (int idx = tid; idx < m * m; idx += nt) {
        const int j = k + 1 + idx / m, i = k + 1 + idx % m;
        F[i + (int64_t)j * s] -= F[i + (int64_t)k * s] * F[k + (int64_t)j * s];
      }
    }
    __syncthreads();
  }
  if (bad) {
    if (tid == 0) atomicExch(info, f + 1);
    return;
  }
  // ---- W = (lower factor)^-1 written straight into the panel (row-major trapezoid)
  K *P = panL + fr.poff;
  const int ldp = hb_ldp(s1);
  auto prow = [&](K *base, int r) -> K * {
    const int k = r / RB;
    return base + hb_blk_off(k) + (int64_t)(r - k * RB) * hb_wblk(s1, k);
  };
  // zero-fill pivot trapezoid (structural zeros above the diagonal + padding)
  for (int64_t q = tid; q < hb_upd_off(s1); q += nt) P[q] = mk(0.0);
  if (!symmetric) {
    K *Q = panU + fr.poff;
    for (int64_t q = tid; q < hb_upd_off(s1); q += nt) Q[q] = mk(0.0);
  }
  __syncthreads();
  // column j of W by forward substitution, one thread per column
  for (int j = tid; j < s1; j += nt) {
    for (int i = j; i < s1; ++i) {
      K acc = mk((i == j) ? 1.0 : 0.0);
      for (int k = j; k < i; ++k) acc -= F[i + (int64_t)k * s] * prow(P, k)[j];
      prow(P, i)[j] = symmetric ? acc / F[i + (int64_t)i * s] : acc;  // LU: unit lower
    }
  }
  if (!symmetric) {
    // V = U11^-1 (upper); store V^T: Q[r][c] = V[c][r], c <= r.  V^T = (U11^T)^-1, U11^T lower with Lt[i][k] = U[k][i]
    K *Q = panU + fr.poff;
    for (int j = tid; j < s1; j += nt) {
      for (int i = j; i < s1; ++i) {
        K acc = mk((i == j) ? 1.0 : 0.0);
        for (int k = j; k < i; ++k) acc -= F[k + (int64_t)i * s] * prow(Q, k)[j];
        prow(Q, i)[j] = acc / F[i + (int64_t)i * s];
      }
    }
  }
  __syncthreads();
  // ---- update rows: M = L21 * W
  K *Pu = P + hb_upd_off(s1);
  for (int idx = tid; idx < s2 * ldp; idx += nt) {
    const int i = idx / ldp, c = idx % ldp;
    K acc = mk(0.0);
    if (c < s1)
      for (int k = c; k < s1; ++k) acc += F[s1 + i + (int64_t)k * s] * prow(P, k)[c];
    Pu[(int64_t)i * ldp + c] = acc;
  }
  if (!symmetric) {
    // N12 = V * U12 ; panU update row i, col c = N12[c][i] = sum_{k>=c} V[c][k] U12[k][i], V[c][k] = Q[k][c]
    K *Q = panU + fr.poff;
    K *Qu = Q + hb_upd_off(s1);
    for (int idx = tid; idx < s2 * ldp; idx += nt) {
      const int i = idx / ldp, c = idx % ldp;
      K acc = mk(0.0);
      if (c < s1)
        for (int k = c; k < s1; ++k) acc += prow(Q, k)[c] * F[k + (int64_t)(s1 + i) * s];
      Qu[(int64_t)i * ldp + c] = acc;
    }
  }
}

// ------------------------------------------------------------------ large fronts: F (col-major) -> panel (row-major trapezoid)
// mode 0: lower part  panel[r][c] = F[r + c*ld]   (pivot rows: c<=r, unit_diag -> 1 on the diagonal)
// mode 1: upper part transposed  panel[r][c] = F[c + r*ld]
__global__ void k_to_panel(const K *__restrict__ Fpiv, int ldpiv, const K *__restrict__ F, int s1, int s2, int mode, K *P) {
  // mode 0: pivot rows from Fpiv (lower triangular, ld = ldpiv), update rows from F21 = F + s1 (ld = s1+s2)
  // mode 1: everything from the upper part of F, transposed (Fpiv unused)
  const int ld = s1 + s2, ldp = hb_ldp(s1);
  const int r0 = blockIdx.x * 32;  // 32 panel rows per CTA
  __shared__ K tile[32][33];
  const int nrows = min(32, s1 + s2 - r0);
  for (int c0 = 0; c0 < ldp; c0 += 32) {
    if (mode == 0) {
      for (int ty = threadIdx.y; ty < 32; ty += blockDim.y) {
        const int r = r0 + threadIdx.x, c = c0 + ty;
        K v = mk(0.0);
        if (r < s1 + s2 && c < s1) v = (r < s1) ? Fpiv[r + (int64_t)c * ldpiv] : F[r + (int64_t)c * ld];
        tile[ty][threadIdx.x] = v;
      }
    } else {
      for (int ty = threadIdx.y; ty < 32; ty += blockDim.y) {
        const int r = r0 + ty, c = c0 + threadIdx.x;  // F[c + r*ld]: coalesced along c
        tile[threadIdx.x][ty] = (r < s1 + s2 && c < s1) ? F[c + (int64_t)r * ld] : mk(0.0);
      }
    }
    __syncthreads();
    for (int ty = threadIdx.y; ty < nrows; ty += blockDim.y) {
      const int r = r0 + ty, c = c0 + threadIdx.x;
      K v = tile[threadIdx.x][ty];
      if (r < s1) {
        const int k = r / RB, w = hb_wblk(s1, k);
        if (c < w) {
          if (c > r) v = mk(0.0);
          P[hb_blk_off(k) + (int64_t)(r - k * RB) * w + c] = v;
        }
      } else if (c < ldp) {
        P[hb_upd_off(s1) + (int64_t)(r - s1) * ldp + c] = v;
      }
    }
    __syncthreads();
  }
}

// S = unit-lower part of the packed LU factors in F11 (strict lower + ones), zero above
__global__ void k_unit_lower(const K *__restrict__ F, int s1, int ld, K *S) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)s1 * s1) return;
  const int r = (int)(t % s1), c = (int)(t / s1);
  S[t] = r > c ? F[r + (int64_t)c * ld] : mk(r == c ? 1.0 : 0.0);
}

struct Libs {
  cusolverDnHandle_t so = nullptr;
  cublasHandle_t bl = nullptr;
  cusolverDnParams_t params = nullptr;
  void *work = nullptr;
  size_t work_bytes = 0;
  void *hwork = nullptr;
  size_t hwork_bytes = 0;
  int *dinfo = nullptr;
  K *scratch = nullptr;
  size_t scratch_bytes = 0;
  ~Libs() {
    if (scratch) cudaFree(scratch);
    if (work) cudaFree(work);
    if (hwork) free(hwork);
    if (dinfo) cudaFree(dinfo);
    if (params) cusolverDnDestroyParams(params);
    if (so) cusolverDnDestroy(so);
    if (bl) cublasDestroy(bl);
  }
  int ensure(size_t dev_bytes, size_t host_bytes) {
    if (dev_bytes > work_bytes) {
      if (work) cudaFree(work);
      HB_CUDA(cudaMalloc(&work, dev_bytes));
      work_bytes = dev_bytes;
    }
    if (host_bytes > hwork_bytes) {
      if (hwork) free(hwork);
      hwork = malloc(host_bytes);
      hwork_bytes = host_bytes;
    }
    return 0;
  }
};

#define HB_SOLVER(call)                                                      \
  do {                                                                       \
    cusolverStatus_t st__ = (call);                                          \
    if (st__ != CUSOLVER_STATUS_SUCCESS) {                                   \
      set_error("cuSOLVER status %d at %s:%d", (int)st__, __FILE__, __LINE__); \
      return HPDDM_B200_ERR_CUDA;                                            \
    }                                                                        \
  } while (0)
#define HB_BLAS(call)                                                      \
  do {                                                                     \
    cublasStatus_t st__ = (call);                                          \
    if (st__ != CUBLAS_STATUS_SUCCESS) {                                   \
      set_error("cuBLAS status %d at %s:%d", (int)st__, __FILE__, __LINE__); \
      return HPDDM_B200_ERR_CUDA;                                          \
    }                                                                      \
  } while (0)

// cuSOLVER / cuBLAS entry points of this build's scalar type
#ifdef HB_COMPLEX
typedef cuDoubleComplex CK;
#define HB_CUDA_K CUDA_C_64F
#define HB_LIB(d, z) z
#else
typedef double CK;
#define HB_CUDA_K CUDA_R_64F
#define HB_LIB(d, z) d
#endif
static inline CK *ck(K *p) { return reinterpret_cast<CK *>(p); }
static inline const CK *ck(const K *p) { return reinterpret_cast<const CK *>(p); }

// `info`: this front's own 4 status words (every cuSOLVER call overwrites its devInfo, a successful one with 0: a slot shared by the
// fronts of a level would only remember the last front)
static int factor_large(Libs &L, cudaStream_t st, K *F, int s1, int s2, bool symmetric, K *panL, K *panU, int *info) {
  const int ld = s1 + s2;
  const K one = mk(1.0), mone = mk(-1.0);
  K *F11 = F, *F21 = F + s1, *F12 = F + (int64_t)s1 * ld, *F22 = F + s1 + (int64_t)s1 * ld;
  size_t wd = 0, wh = 0;
  if (symmetric) {
#ifdef HB_COMPLEX
    set_error("numfact: LL^T is not available for complex scalars");
    return HPDDM_B200_ERR_STATE;
#else
    int lwork = 0;
    HB_SOLVER(cusolverDnDpotrf_bufferSize(L.so, CUBLAS_FILL_MODE_LOWER, s1, F11, ld, &lwork));
    HB_CHECK(L.ensure((size_t)lwork * sizeof(double), 0));
    HB_SOLVER(cusolverDnDpotrf(L.so, CUBLAS_FILL_MODE_LOWER, s1, F11, ld, (double *)L.work, lwork, info));
    if (s2 > 0) {
      HB_BLAS(cublasDtrsm(L.bl, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, s2, s1, &one, F11, ld, F21, ld));
      HB_BLAS(cublasDsyrk(L.bl, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, s2, s1, &mone, F21, ld, &one, F22, ld));
    }
    HB_SOLVER(cusolverDnXtrtri_bufferSize(L.so, CUBLAS_FILL_MODE_LOWER, CUBLAS_DIAG_NON_UNIT, s1, CUDA_R_64F, F11, ld, &wd, &wh));
    HB_CHECK(L.ensure(wd, wh));
    HB_SOLVER(cusolverDnXtrtri(L.so, CUBLAS_FILL_MODE_LOWER, CUBLAS_DIAG_NON_UNIT, s1, CUDA_R_64F, F11, ld, L.work, wd, L.hwork, wh, info + 1));
    if (s2 > 0) HB_BLAS(cublasDtrmm(L.bl, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, s2, s1, &one, F11, ld, F21, ld, F21, ld));
    k_to_panel<<<(s1 + s2 + 31) / 32, dim3(32, 8), 0, st>>>(F, ld, F, s1, s2, 0, panL);
#endif
  } else {
    int lwork = 0;
    HB_SOLVER(HB_LIB(cusolverDnDgetrf_bufferSize, cusolverDnZgetrf_bufferSize)(L.so, s1, s1, ck(F11), ld, &lwork));
    HB_CHECK(L.ensure((size_t)lwork * sizeof(K), 0));
    HB_SOLVER(HB_LIB(cusolverDnDgetrf, cusolverDnZgetrf)(L.so, s1, s1, ck(F11), ld, (CK *)L.work, nullptr, info));  // devIpiv = NULL: no pivoting
    if (s2 > 0) {
      HB_BLAS(HB_LIB(cublasDtrsm, cublasZtrsm)(L.bl, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, s2, s1, ck(&one), ck(F11), ld, ck(F21), ld));
      HB_BLAS(HB_LIB(cublasDtrsm, cublasZtrsm)(L.bl, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_UNIT, s1, s2, ck(&one), ck(F11), ld, ck(F12), ld));
      HB_BLAS(HB_LIB(cublasDgemm, cublasZgemm)(L.bl, CUBLAS_OP_N, CUBLAS_OP_N, s2, s2, s1, ck(&mone), ck(F21), ld, ck(F12), ld, ck(&one), ck(F22), ld));
    }
    // L11 and U11 share F11: split the unit-lower factor into a scratch matrix so that both
    // inversions are plain non-unit trtri calls
    if ((size_t)s1 * s1 * sizeof(K) > L.scratch_bytes) {
      if (L.scratch) cudaFree(L.scratch);
      L.scratch = nullptr;
      HB_CUDA(cudaMalloc(&L.scratch, (size_t)s1 * s1 * sizeof(K)));
      L.scratch_bytes = (size_t)s1 * s1 * sizeof(K);
    }
    K *S = L.scratch;
    k_unit_lower<<<(unsigned)(((int64_t)s1 * s1 + 255) / 256), 256, 0, st>>>(F11, s1, ld, S);
    HB_SOLVER(cusolverDnXtrtri_bufferSize(L.so, CUBLAS_FILL_MODE_LOWER, CUBLAS_DIAG_NON_UNIT, s1, HB_CUDA_K, S, s1, &wd, &wh));
    HB_CHECK(L.ensure(wd, wh));
    HB_SOLVER(cusolverDnXtrtri(L.so, CUBLAS_FILL_MODE_LOWER, CUBLAS_DIAG_NON_UNIT, s1, HB_CUDA_K, S, s1, L.work, wd, L.hwork, wh, info + 1));
    HB_SOLVER(cusolverDnXtrtri_bufferSize(L.so, CUBLAS_FILL_MODE_UPPER, CUBLAS_DIAG_NON_UNIT, s1, HB_CUDA_K, F11, ld, &wd, &wh));
    HB_CHECK(L.ensure(wd, wh));
    HB_SOLVER(cusolverDnXtrtri(L.so, CUBLAS_FILL_MODE_UPPER, CUBLAS_DIAG_NON_UNIT, s1, HB_CUDA_K, F11, ld, L.work, wd, L.hwork, wh, info + 2));
    if (s2 > 0) {
      HB_BLAS(HB_LIB(cublasDtrmm, cublasZtrmm)(L.bl, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, s2, s1, ck(&one), ck(S), s1, ck(F21), ld, ck(F21), ld));
      HB_BLAS(HB_LIB(cublasDtrmm, cublasZtrmm)(L.bl, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, s1, s2, ck(&one), ck(F11), ld, ck(F12), ld, ck(F12), ld));
    }
    k_to_panel<<<(s1 + s2 + 31) / 32, dim3(32, 8), 0, st>>>(S, s1, F, s1, s2, 0, panL);
    k_to_panel<<<(s1 + s2 + 31) / 32, dim3(32, 8), 0, st>>>(nullptr, 0, F, s1, s2, 1, panU);
  }
  return 0;
}

// NOTE: a plain cudaMemcpy from pageable memory may return before the DMA has landed and the
// legacy default stream does not order against our non-blocking stream: every upload is
// issued on the context's stream and waited for.
template <class T>
static int upload(const std::vector<T> &v, T **d, cudaStream_t st) {
  *d = nullptr;
  if (v.empty()) return 0;
  HB_CUDA(cudaMalloc(d, v.size() * sizeof(T)));
  HB_CUDA(cudaMemcpyAsync(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  HB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

static int numfact_try(Sub *s, const HostCSR &A, bool symmetric) {
  Symbolic &S = s->sym;
  DeviceFactor &D = s->fac;
  cudaStream_t st = s->ctx->stream;
  const int F = (int)S.fronts.size();
  const int n = S.n;
  // ---- host: matrix entries -> (front,row,col), grouped by level
  std::vector<int64_t> lvl_cnt(S.nlevels + 1, 0);
  auto loc = [&](int f, int p) {
    const Front &fr = S.fronts[f];
    if (p < fr.p0 + fr.s1) return p - fr.p0;
    const int *b = S.rowidx.data() + fr.rptr;
    return fr.s1 + (int)(std::lower_bound(b, b + fr.s2, p) - b);
  };
  std::vector<int> efront, erow, ecol;
  std::vector<K> eval;
  {
    const int64_t nnz = A.ia[n];
    std::vector<int> tf, tr, tc;
    std::vector<K> tv;
    tf.reserve(nnz);
    tr.reserve(nnz);
    tc.reserve(nnz);
    tv.reserve(nnz);
    for (int i = 0; i < n; ++i) {
      const int pi = S.iperm[i];
      for (int k = A.ia[i]; k < A.ia[i + 1]; ++k) {
        const int pj = S.iperm[A.ja[k]];
        if (symmetric && pi < pj) continue;
        int f, r, c;
        if (pi >= pj) {
          f = S.front_of[pj];
          r = loc(f, pi);
          c = pj - S.fronts[f].p0;
        } else {
          f = S.front_of[pi];
          r = pi - S.fronts[f].p0;
          c = loc(f, pj);
        }
        tf.push_back(f);
        tr.push_back(r);
        tc.push_back(c);
        tv.push_back(A.a[k]);
        lvl_cnt[S.fronts[f].level + 1]++;
      }
    }
    for (int l = 0; l < S.nlevels; ++l) lvl_cnt[l + 1] += lvl_cnt[l];
    const size_t ne = tf.size();
    efront.resize(ne);
    erow.resize(ne);
    ecol.resize(ne);
    eval.resize(ne);
    std::vector<int64_t> pos(lvl_cnt.begin(), lvl_cnt.end() - 1);
    for (size_t e = 0; e < ne; ++e) {
      int64_t q = pos[S.fronts[tf[e]].level]++;
      efront[q] = tf[e];
      erow[q] = tr[e];
      ecol[q] = tc[e];
      eval[q] = tv[e];
    }
  }
  // ---- per-front F offsets (per level), uploaded once
  std::vector<int64_t> foff(F, 0);
  std::vector<int> fs(F, 0);
  std::vector<int64_t> lvl_elems(S.nlevels, 0);
  for (int l = 0; l < S.nlevels; ++l) {
    int64_t off = 0;
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
      int f = S.level_order[q];
      const Front &fr = S.fronts[f];
      foff[f] = off;
      fs[f] = fr.s1 + fr.s2;
      off += (int64_t)fs[f] * fs[f];
    }
    lvl_elems[l] = off;
  }
  int *d_efront = nullptr, *d_erow = nullptr, *d_ecol = nullptr, *d_fs = nullptr, *d_rel = nullptr, *d_list = nullptr;
  K *d_eval = nullptr;
  int64_t *d_foff = nullptr;
  EATask *d_tasks = nullptr;
  int *d_info = nullptr;
  K *Fprev = nullptr, *Fcur = nullptr;
  Libs L;
  int rc = 0;
  auto cleanup = [&]() {
    cudaFree(d_efront);
    cudaFree(d_erow);
    cudaFree(d_ecol);
    cudaFree(d_eval);
    cudaFree(d_fs);
    cudaFree(d_foff);
    cudaFree(d_rel);
    cudaFree(d_list);
    cudaFree(d_tasks);
    cudaFree(d_info);
    cudaFree(Fprev);
    cudaFree(Fcur);
  };
#define NF_CHECK(x)   \
  do {                \
    rc = (x);         \
    if (rc < 0) {     \
      cleanup();      \
      return rc;      \
    }                 \
  } while (0)
#define NF_CUDA(call)                                                                             \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      cleanup();                                                                                  \
      return e__ == cudaErrorMemoryAllocation ? HPDDM_B200_ERR_NOMEM : HPDDM_B200_ERR_CUDA;        \
    }                                                                                             \
  } while (0)
  NF_CHECK(upload(efront, &d_efront, st));
  NF_CHECK(upload(erow, &d_erow, st));
  NF_CHECK(upload(ecol, &d_ecol, st));
  NF_CHECK(upload(eval, &d_eval, st));
  NF_CHECK(upload(fs, &d_fs, st));
  NF_CHECK(upload(foff, &d_foff, st));
  NF_CHECK(upload(S.rel, &d_rel, st));
  NF_CUDA(cudaMalloc(&d_info, 4 * sizeof(int)));
  NF_CUDA(cudaMemset(d_info, 0, 4 * sizeof(int)));
  if (cusolverDnCreate(&L.so) != CUSOLVER_STATUS_SUCCESS || cublasCreate(&L.bl) != CUBLAS_STATUS_SUCCESS) {
    set_error("cannot create cuSOLVER/cuBLAS handles");
    cleanup();
    return HPDDM_B200_ERR_CUDA;
  }
  cusolverDnSetStream(L.so, st);
  cublasSetStream(L.bl, st);
  int dinfo_cap = 0;  // status slots (4 words per cuSOLVER front of the current level)
  // ---- panel store
  if (!D.panL) NF_CUDA(cudaMalloc(&D.panL, std::max<int64_t>(S.panel_elems, 16) * sizeof(K)));
  if (!symmetric) {
    if (!D.panU || D.panU == D.panL) NF_CUDA(cudaMalloc(&D.panU, std::max<int64_t>(S.panel_elems, 16) * sizeof(K)));
  } else {
    if (D.panU && D.panU != D.panL) cudaFree(D.panU);
    D.panU = D.panL;
  }
  D.symmetric = symmetric;
  std::vector<int> small_list;
  std::vector<EATask> tasks;
  for (int l = 0; l < S.nlevels; ++l) {
    NF_CUDA(cudaMalloc(&Fcur, std::max<int64_t>(lvl_elems[l], 1) * sizeof(K)));
    NF_CUDA(cudaMemsetAsync(Fcur, 0, lvl_elems[l] * sizeof(K), st));
    const int64_t e0 = lvl_cnt[l], e1 = lvl_cnt[l + 1];
    if (e1 > e0) {
      k_assemble<<<(unsigned)((e1 - e0 + 255) / 256), 256, 0, st>>>(e1 - e0, d_efront + e0, d_erow + e0, d_ecol + e0, d_eval + e0, d_foff, d_fs, Fcur);
      s->ctx->launches++;
    }
    if (l > 0) {
      tasks.clear();
      for (int q = S.level_ptr[l - 1]; q < S.level_ptr[l]; ++q) {
        int c = S.level_order[q];
        if (S.fronts[c].parent < 0) continue;
        for (int j0 = 0; j0 < S.fronts[c].s2; j0 += 16) tasks.push_back({c, j0});
      }
      if (!tasks.empty()) {
        cudaFree(d_tasks);
        d_tasks = nullptr;
        NF_CHECK(upload(tasks, &d_tasks, st));
        k_extend_add<<<(unsigned)tasks.size(), 128, 0, st>>>(d_tasks, D.fronts, d_rel, d_foff, Fprev, Fcur, symmetric ? 1 : 0);
        s->ctx->launches++;
      }
    }
    small_list.clear();
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
      int f = S.level_order[q];
      if (fs[f] <= SMALL) small_list.push_back(f);
    }
    if (!small_list.empty()) {
      cudaFree(d_list);
      d_list = nullptr;
      NF_CHECK(upload(small_list, &d_list, st));
      k_factor_small<<<(unsigned)small_list.size(), 256, 0, st>>>(d_list, D.fronts, d_foff, Fcur, D.panL, D.panU, symmetric ? 1 : 0, d_info);
      s->ctx->launches++;
    }
    int nlarge = 0;
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) nlarge += fs[S.level_order[q]] > SMALL;
    if (nlarge > dinfo_cap) {
      if (L.dinfo) cudaFree(L.dinfo);
      L.dinfo = nullptr;
      NF_CUDA(cudaMalloc(&L.dinfo, (size_t)4 * nlarge * sizeof(int)));
      dinfo_cap = nlarge;
    }
    if (nlarge) NF_CUDA(cudaMemsetAsync(L.dinfo, 0, (size_t)4 * nlarge * sizeof(int), st));
    std::vector<int> large_fronts;
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
      int f = S.level_order[q];
      if (fs[f] <= SMALL) continue;
      const Front &fr = S.fronts[f];
      NF_CHECK(factor_large(L, st, Fcur + foff[f], fr.s1, fr.s2, symmetric, D.panL + fr.poff, D.panU + fr.poff, L.dinfo + 4 * large_fronts.size()));
      large_fronts.push_back(f);
      s->ctx->launches += 8;
    }
    // pivots / library status of this level: every cuSOLVER front has its own slot
    int hinfo[4] = {0, 0, 0, 0}, linfo[4] = {0, 0, 0, 0};
    std::vector<int> slots((size_t)4 * nlarge, 0);
    NF_CUDA(cudaMemcpyAsync(hinfo, d_info, sizeof(hinfo), cudaMemcpyDeviceToHost, st));
    if (nlarge) NF_CUDA(cudaMemcpyAsync(slots.data(), L.dinfo, slots.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
    NF_CUDA(cudaStreamSynchronize(st));
    int bad_large = -1;
    for (int q = 0; q < nlarge && bad_large < 0; ++q)
      if (slots[4 * q] != 0 || slots[4 * q + 1] != 0 || slots[4 * q + 2] != 0) {
        bad_large = large_fronts[q];
        for (int k = 0; k < 3; ++k) linfo[k] = slots[4 * q + k];
      }
    if (getenv("HPDDM_B200_DEBUG")) {
      int nl = 0;
      for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) nl += fs[S.level_order[q]] > SMALL;
      fprintf(stderr, "[hpddm_b200] level %d: %d fronts (%d via cuSOLVER), F buffer %.3f GB, info small=%d lib=%d/%d/%d\n", l, S.level_ptr[l + 1] - S.level_ptr[l], nl,
              lvl_elems[l] * sizeof(K) * 1e-9, hinfo[0], linfo[0], linfo[1], linfo[2]);
      if (hinfo[0] != 0) {
        const int f = hinfo[0] - 1;
        const Front &fr = S.fronts[f];
        fprintf(stderr, "[hpddm_b200]   failing front %d: p0 %d s1 %d s2 %d level %d parent %d children:", f, fr.p0, fr.s1, fr.s2, fr.level, fr.parent);
        for (int c : S.children[f]) fprintf(stderr, " %d(s1 %d s2 %d lvl %d)", c, S.fronts[c].s1, S.fronts[c].s2, S.fronts[c].level);
        fprintf(stderr, "\n");
        std::vector<K> hf((size_t)fs[f] * fs[f]);
        cudaMemcpyAsync(hf.data(), Fcur + foff[f], hf.size() * sizeof(K), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "[hpddm_b200]   diag(F) after partial factorisation:");
        for (int k = 0; k < fs[f] && k < 40; ++k) fprintf(stderr, " %.3g", hb_real(hf[k + (size_t)k * fs[f]]));
        fprintf(stderr, "\n");
      }
    }
    if (hinfo[0] != 0 || linfo[0] != 0 || linfo[1] != 0 || linfo[2] != 0) {
      set_error("numfact: %s pivot breakdown at level %d (front %d, potrf/getrf info %d, trtri info %d)", symmetric ? "Cholesky" : "LU", l,
                hinfo[0] != 0 ? hinfo[0] - 1 : bad_large, linfo[0], linfo[1]);
      cleanup();
      return HPDDM_B200_ERR_NUMERIC;
    }
    cudaFree(Fprev);
    Fprev = Fcur;
    Fcur = nullptr;
  }
  cleanup();
  NF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// Einv = E^-1 on the device for large coarse operators (N_c beyond what a host Gauss-Jordan should do): LU with partial pivoting
// (cuSOLVER getrf) + getrs on the identity.  dE is overwritten by its factors' input copy only through a scratch buffer.
__global__ void k_identity(int N, K *M) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < (int64_t)N * N) M[t] = mk((t % N) == (t / N) ? 1.0 : 0.0);
}
int dense_inverse_device(Ctx *c, int N, const K *dE, K *dEinv) {
  cudaStream_t st = c->stream;
  Libs L;
  if (cusolverDnCreate(&L.so) != CUSOLVER_STATUS_SUCCESS) {
    set_error("cannot create a cuSOLVER handle");
    return HPDDM_B200_ERR_CUDA;
  }
  cusolverDnSetStream(L.so, st);
  K *LU = nullptr;
  int *ipiv = nullptr, *info = nullptr, lwork = 0, hinfo = 0;
  auto done = [&](int rc) {
    cudaFree(LU);
    cudaFree(ipiv);
    cudaFree(info);
    return rc;
  };
  if (cudaMalloc(&LU, (size_t)N * N * sizeof(K)) != cudaSuccess || cudaMalloc(&ipiv, N * sizeof(int)) != cudaSuccess || cudaMalloc(&info, sizeof(int)) != cudaSuccess) {
    set_error("out of device memory inverting the %d x %d coarse operator", N, N);
    return done(HPDDM_B200_ERR_NOMEM);
  }
  cudaMemcpyAsync(LU, dE, (size_t)N * N * sizeof(K), cudaMemcpyDeviceToDevice, st);
  if (HB_LIB(cusolverDnDgetrf_bufferSize, cusolverDnZgetrf_bufferSize)(L.so, N, N, ck(LU), N, &lwork) != CUSOLVER_STATUS_SUCCESS || L.ensure((size_t)lwork * sizeof(K), 0) < 0)
    return done(HPDDM_B200_ERR_CUDA);
  if (HB_LIB(cusolverDnDgetrf, cusolverDnZgetrf)(L.so, N, N, ck(LU), N, (CK *)L.work, ipiv, info) != CUSOLVER_STATUS_SUCCESS) return done(HPDDM_B200_ERR_CUDA);
  k_identity<<<(unsigned)(((int64_t)N * N + 255) / 256), 256, 0, st>>>(N, dEinv);
  if (HB_LIB(cusolverDnDgetrs, cusolverDnZgetrs)(L.so, CUBLAS_OP_N, N, N, ck(LU), N, ipiv, ck(dEinv), N, info) != CUSOLVER_STATUS_SUCCESS) return done(HPDDM_B200_ERR_CUDA);
  cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess || hinfo != 0) {
    set_error("coarse operator is singular (cuSOLVER info %d)", hinfo);
    return done(HPDDM_B200_ERR_NUMERIC);
  }
  return done(0);
}

void free_factor(DeviceFactor &f) {
  for (cudaGraphExec_t &g : f.graph)
    if (g) cudaGraphExecDestroy(g);
  if (f.panU && f.panU != f.panL) cudaFree(f.panU);
  cudaFree(f.panL);
  cudaFree(f.fronts);
  cudaFree(f.rowidx);
  cudaFree(f.fwd);
  cudaFree(f.bwd);
  cudaFree(f.perm);
  cudaFree(f.b);
  cudaFree(f.y);
  cudaFree(f.x);
  for (void *q : {(void *)f.fwd_children, (void *)f.fwd_total, (void *)f.bwd_total, (void *)f.bwd_ordered, (void *)f.sync_pending, (void *)f.sync_done, (void *)f.sync_next})
    if (q) cudaFree(q);
  if (f.sync_err_host) cudaFreeHost(f.sync_err_host);
  f = DeviceFactor();
}

int numfact_device(Sub *s, const HostCSR &A) {
  cudaStream_t st = s->ctx->stream;
  auto t0 = std::chrono::steady_clock::now();
  free_factor(s->fac);
  int leaf = 64;
  if (const char *e = getenv("HPDDM_B200_SMALL")) SMALL = atoi(e);
  if (const char *e = getenv("HPDDM_B200_LEAF")) leaf = std::max(1, atoi(e));
  HB_CHECK(symbolic_analyze(A, s->gx, s->gy, s->gz, s->gdof, leaf, s->sym));
  auto t1 = std::chrono::steady_clock::now();
  s->t_symbolic = std::chrono::duration<double>(t1 - t0).count();
  Symbolic &S = s->sym;
  DeviceFactor &D = s->fac;
  HB_CHECK(upload(S.fronts, &D.fronts, st));
  HB_CHECK(upload(S.rowidx, &D.rowidx, st));
  HB_CHECK(upload(S.fwd, &D.fwd, st));
  HB_CHECK(upload(S.bwd, &D.bwd, st));
  HB_CHECK(upload(S.perm, &D.perm, st));
  HB_CUDA(cudaMalloc(&D.b, (size_t)S.n * 8 * sizeof(K)));
  HB_CUDA(cudaMalloc(&D.y, (size_t)S.n * 8 * sizeof(K)));
  HB_CUDA(cudaMalloc(&D.x, (size_t)S.n * 8 * sizeof(K)));
  HB_CHECK(sptrsv_prepare(s));
  D.sweep_launches = 0;
  for (int l = 0; l < S.nlevels; ++l) D.sweep_launches += (S.fwd_ptr[l + 1] > S.fwd_ptr[l]) + (S.bwd_ptr[l + 1] > S.bwd_ptr[l]);
  int rc = HPDDM_B200_ERR_NUMERIC;
  const bool force_lu = getenv("HPDDM_B200_FORCE_LU") != nullptr;
  if (A.symmetric && !force_lu) rc = numfact_try(s, A, true);
  if (rc == HPDDM_B200_ERR_NUMERIC) {
    if (A.symmetric && !force_lu) fprintf(stderr, "[hpddm_b200] subdomain %d: Cholesky failed (%s); retrying with LU\n", s->grank, get_error_msg());
    rc = numfact_try(s, A, false);
  }
  if (rc < 0) {
    free_factor(D);
    return rc;
  }
  D.valid = true;
  s->t_numfact = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
  return 0;
}

}  // namespace hb
