// CPU stand-in for everything BELOW the orchestration layer of libhpddm_b200.so, so that hb_api.cu, hb_krylov.cu, hb_geneo.cu,
// hb_gcrodr.cpp and hb_symbolic.cpp -- every exported entry point and all of its host logic: halo planning and message order, coarse
// layout, staging of caller memory, lazy host registration, apply / deflation orchestration for every correction and Prcndtnr, the device
// Krylov drivers -- compile with g++ and run on a machine without a GPU (and under AddressSanitizer).  Mocked: the CUDA runtime calls
// those files make ("device" memory is host memory, streams are synchronous), the kernel LAUNCHERS of hb_kernels.cu (a plain loop each,
// written from the kernel it stands for), the local factorisation / triangular solves of hb_numfact.cu + hb_solve.cu (a banded LU with
// partial pivoting on the host), the dense coarse inverse (Gauss-Jordan), and the peer-memory fabric (single process: off).
// TEST INFRASTRUCTURE ONLY (tests/tools/run_gpu_tests_on_stand_in.py); never linked into the product, never loaded by hpddm_b200/.
// Compiled twice like every source of the library (K = double, and -DHB_COMPLEX).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>

#include "../../hpddm_b200/csrc/hb_internal.h"

#ifndef HB_COMPLEX
// ---- CUDA runtime (defined once for both scalar builds)
extern "C" {
cudaError_t cudaMalloc(void **p, size_t bytes) {
  *p = malloc(bytes ? bytes : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) {
  free(p);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) {
  memset(p, v, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
  memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memmove(static_cast<char *>(dst) + r * dpitch, static_cast<const char *>(src) + r * spitch, width);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) {
  *n = 1;
  return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "device stand-in"; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
  *s = reinterpret_cast<cudaStream_t>(malloc(1));
  return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s) {
  free(s);
  return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) {
  *e = reinterpret_cast<cudaEvent_t>(malloc(1));
  return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e) {
  free(e);
  return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
// lazy pinning of caller memory: bookkeeping only
static std::map<uintptr_t, size_t> g_registered;
cudaError_t cudaHostRegister(void *p, size_t bytes, unsigned) {
  g_registered[reinterpret_cast<uintptr_t>(p)] = bytes;
  return cudaSuccess;
}
cudaError_t cudaHostUnregister(void *p) { return g_registered.erase(reinterpret_cast<uintptr_t>(p)) ? cudaSuccess : cudaErrorHostMemoryNotRegistered; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *at, const void *p) {
  memset(at, 0, sizeof(*at));
  at->type = cudaMemoryTypeUnregistered;
  const uintptr_t q = reinterpret_cast<uintptr_t>(p);
  for (const auto &r : g_registered)
    if (q >= r.first && q < r.first + r.second) at->type = cudaMemoryTypeHost;
  return cudaSuccess;
}
// stream capture / graphs: not available on the stand-in (the whole-apply graph is an opt-in of the real library)
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorNotSupported; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *) { return cudaErrorNotSupported; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long) { return cudaErrorNotSupported; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
}
#endif

namespace hb {

#define MOCK_END(c) \
  (c)->launches++;  \
  return 0

// ---- elementwise / reductions (hb_kernels.cu)
int k_scale(Ctx *c, int n, int mu, const double *d, const K *in, K *out) {
  for (int64_t t = 0; t < (int64_t)n * mu; ++t) out[t] = d[t % n] * in[t];
  MOCK_END(c);
}
int k_axpy(Ctx *c, int64_t n, double a, const K *x, K *y) {
  for (int64_t t = 0; t < n; ++t) y[t] = y[t] + a * x[t];
  MOCK_END(c);
}
int k_copy(Ctx *c, int64_t n, const K *x, K *y) {
  if (n && x != y) memmove(y, x, n * sizeof(K));
  MOCK_END(c);
}
int k_fill(Ctx *c, int64_t n, K v, K *y) {
  for (int64_t t = 0; t < n; ++t) y[t] = v;
  MOCK_END(c);
}
int k_scal_copy(Ctx *c, int64_t n, double a, const K *x, K *y) {
  for (int64_t t = 0; t < n; ++t) y[t] = a * x[t];
  MOCK_END(c);
}
int k_flush_tiny(Ctx *c, int64_t n, double tiny, K *v) {
  for (int64_t t = 0; t < n; ++t)
    if (hb_abs(v[t]) < tiny) v[t] = mk(0.0);
  MOCK_END(c);
}
// out[i,c] = (d ? d[i] : 1) * (beta * yin[i,c] + alpha * sum_k a[k] x[ja[k],c])
int k_spmv_raw(Ctx *c, int n, int64_t, const int *ia, const int *ja, const K *a, int mu, double alpha, const K *x, double beta, const K *yin, K *out, const double *d) {
  std::vector<K> col(n);  // out may alias yin, never x
  for (int cc = 0; cc < mu; ++cc) {
    for (int i = 0; i < n; ++i) {
      K acc = mk(0.0);
      for (int k = ia[i]; k < ia[i + 1]; ++k) acc = acc + a[k] * x[ja[k] + (int64_t)cc * n];
      K v = alpha * acc;
      if (beta != 0.0) v = v + beta * yin[i + (int64_t)cc * n];
      col[i] = (d ? d[i] : 1.0) * v;
    }
    for (int i = 0; i < n; ++i) out[i + (int64_t)cc * n] = col[i];
  }
  MOCK_END(c);
}
int k_spmv(Ctx *c, const Sub *s, int mu, double alpha, const K *x, double beta, const K *yin, K *out, const double *d) {
  if (s->n == 0) return 0;
  return k_spmv_raw(c, s->n, s->A.ia[s->n], s->d_ia, s->d_ja, s->d_a, mu, alpha, x, beta, yin, out, d);
}
// T[k + ldT*col] += sum_i conj(Z[i + k*n]) d[i] x[i + col*n]
int k_zt_raw(Ctx *c, int n, int nu, const K *Z, const double *d, int mu, const K *x, K *T, int ldT) {
  for (int col = 0; col < mu; ++col)
    for (int k = 0; k < nu; ++k) {
      K acc = mk(0.0);
      for (int i = 0; i < n; ++i) acc = acc + hb_conj(Z[i + (int64_t)k * n]) * (d[i] * x[i + (int64_t)col * n]);
      T[k + (int64_t)ldT * col] = T[k + (int64_t)ldT * col] + acc;
    }
  MOCK_END(c);
}
int k_zt_project(Ctx *c, const Sub *s, int mu, const K *x, K *T, int ldT) { return k_zt_raw(c, s->n, s->nu, s->d_Z, s->d_d, mu, x, T, ldT); }
// out[i + col*n] = d[i] * sum_k Z[i + k*n] Y[k + ldY*col]
int k_zexp_raw(Ctx *c, int n, int nu, const K *Z, const double *d, int mu, const K *Y, int ldY, K *out) {
  if (n == 0 || nu == 0) return 0;
  for (int col = 0; col < mu; ++col)
    for (int i = 0; i < n; ++i) {
      K acc = mk(0.0);
      for (int k = 0; k < nu; ++k) acc = acc + Z[i + (int64_t)k * n] * Y[k + (int64_t)ldY * col];
      out[i + (int64_t)col * n] = d[i] * acc;
    }
  MOCK_END(c);
}
int k_z_expand(Ctx *c, const Sub *s, int mu, const K *Y, int ldY, K *out) {
  if (s->n == 0) return 0;
  if (s->nu == 0) return k_fill(c, (int64_t)s->n * mu, mk(0.0), out);
  return k_zexp_raw(c, s->n, s->nu, s->d_Z, s->d_d, mu, Y, ldY, out);
}
// send[ebase*mu + c*esize + (e - ebase)] = x[map[e] + c*n]
int k_pack(Ctx *c, const Sub *s, int mu, const K *x, K *send) {
  for (int cc = 0; cc < mu; ++cc)
    for (int e = 0; e < s->h; ++e) send[(int64_t)s->d_ebase[e] * mu + (int64_t)cc * s->d_esize[e] + (e - s->d_ebase[e])] = x[s->d_map[e] + (int64_t)cc * s->n];
  MOCK_END(c);
}
// x[uidx[u]] += contributions of the unique target u, in neighbour order
int k_unpack(Ctx *c, const Sub *s, int mu, K *x) {
  for (int cc = 0; cc < mu; ++cc)
    for (int u = 0; u < s->nuniq; ++u) {
      K acc = x[s->d_uidx[u] + (int64_t)cc * s->n];
      for (int q = s->d_useg[u]; q < s->d_useg[u + 1]; ++q) {
        const int e = s->d_upos[q];
        acc = acc + s->d_recv[(int64_t)s->d_ebase[e] * mu + (int64_t)cc * s->d_esize[e] + (e - s->d_ebase[e])];
      }
      x[s->d_uidx[u] + (int64_t)cc * s->n] = acc;
    }
  MOCK_END(c);
}
int k_dot(Ctx *c, const Sub *s, int mu, const K *x, const K *y, K *res) {
  for (int cc = 0; cc < mu; ++cc) {
    K acc = mk(0.0);
    for (int i = 0; i < s->n; ++i) acc = acc + (s->d_d[i] * hb_conj(x[i + (int64_t)cc * s->n])) * y[i + (int64_t)cc * s->n];
    res[cc] = res[cc] + acc;
  }
  MOCK_END(c);
}
int k_rhs_norm(Ctx *c, const Sub *s, int mu, const K *b, double *res) {
  constexpr double PEN = 1.0e30, EPS = 1.0e-12;
  for (int cc = 0; cc < mu; ++cc)
    for (int i = 0; i < s->n; ++i) {
      const K fv = b[i + (int64_t)cc * s->n];
      const bool flagged = s->d_bcflag && s->d_bcflag[i];
      res[cc] += s->d_d[i] * ((hb_abs(fv) > PEN * EPS && flagged) ? hb_norm(fv / PEN) : hb_norm(fv));
    }
  MOCK_END(c);
}
int k_residual_norms(Ctx *c, const Sub *s, int mu, int norm, const K *f, const K *t, double *res) {
  constexpr double PEN = 1.0e30, EPS = 1.0e-12;
  for (int cc = 0; cc < mu; ++cc)
    for (int i = 0; i < s->n; ++i) {
      const K fv = f[i + (int64_t)cc * s->n];
      const bool flagged = s->d_bcflag && s->d_bcflag[i];
      const double af = hb_abs(fv) > EPS * PEN ? hb_abs(fv / PEN) : hb_abs(fv);
      const double at = flagged ? 0.0 : hb_abs(t[i + (int64_t)cc * s->n]);
      if (norm == 1) {
        res[2 * cc] += s->d_d[i] * af;
        res[2 * cc + 1] += s->d_d[i] * at;
      } else if (norm == 2) {
        res[2 * cc] = std::max(res[2 * cc], af);
        res[2 * cc + 1] = std::max(res[2 * cc + 1], at);
      } else {
        res[2 * cc] += s->d_d[i] * af * af;
        res[2 * cc + 1] += s->d_d[i] * at * at;
      }
    }
  MOCK_END(c);
}
int k_bc(Ctx *c, const Sub *s, int mu, const K *b, K *x) {
  const int nbc = (int)s->bc.size();
  for (int cc = 0; cc < mu; ++cc)
    for (int q = 0; q < nbc; ++q) x[s->d_bc_idx[q] + (int64_t)cc * s->n] = b[s->d_bc_idx[q] + (int64_t)cc * s->n] / s->d_bc_val[q];
  MOCK_END(c);
}
// replicated coarse solve: Y = Einv T ; R = T - E Y ; Y += Einv R, vectors in the layout [proc][col][row-in-proc]
int k_coarse_solve(Ctx *c, int mu) {
  const int Nc = c->Nc;
  if (Nc == 0) return 0;
  auto at = [&](int r, int cc) -> int64_t { return (int64_t)c->d_rowproc[r] * c->Lnu * mu + (int64_t)cc * c->Lnu + c->d_rowloc[r]; };
  std::vector<K> R(Nc);
  for (int cc = 0; cc < mu; ++cc) {
    for (int r = 0; r < Nc; ++r) {
      K acc = mk(0.0);
      for (int k = 0; k < Nc; ++k) acc = acc + c->d_Einv[r + (int64_t)k * Nc] * c->d_T[at(k, cc)];
      c->d_Y[at(r, cc)] = acc;
    }
    for (int r = 0; r < Nc; ++r) {
      K acc = c->d_T[at(r, cc)];
      for (int k = 0; k < Nc; ++k) acc = acc - c->d_E[r + (int64_t)k * Nc] * c->d_Y[at(k, cc)];
      R[r] = acc;
    }
    for (int r = 0; r < Nc; ++r) {
      K acc = mk(0.0);
      for (int k = 0; k < Nc; ++k) acc = acc + c->d_Einv[r + (int64_t)k * Nc] * R[k];
      c->d_Y[at(r, cc)] = c->d_Y[at(r, cc)] + acc;
    }
  }
  MOCK_END(c);
}
// Krylov helpers
int k_vdots(Ctx *c, const Sub *s, int k, const K *V, int64_t ldv, const K *w, K *T) {
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < s->n; ++i) T[j] = T[j] + s->d_d[i] * (hb_conj(V[i + j * ldv]) * w[i]);
  MOCK_END(c);
}
int k_vupdate(Ctx *c, const Sub *s, int k, const K *V, int64_t ldv, const K *h, double sign, K *w) {
  for (int i = 0; i < s->n; ++i) {
    K acc = mk(0.0);
    for (int j = 0; j < k; ++j) acc = acc + V[i + j * ldv] * h[j];
    w[i] = w[i] + sign * acc;
  }
  MOCK_END(c);
}
int k_vupdate_blk(Ctx *c, int n, int k, int mu, const K *V, const K *H, int ldh, double sign, K *W) {
  for (int col = 0; col < mu; ++col)
    for (int i = 0; i < n; ++i) {
      K acc = mk(0.0);
      for (int j = 0; j < k; ++j) acc = acc + V[i + (size_t)j * n] * H[j + (size_t)ldh * col];
      W[i + (size_t)col * n] = W[i + (size_t)col * n] + sign * acc;
    }
  MOCK_END(c);
}
int k_rmul_upper(Ctx *c, int n, int mu, const K *R, K *W) {
  std::vector<K> row(mu);
  for (int i = 0; i < n; ++i) {
    for (int col = 0; col < mu; ++col) {
      K acc = mk(0.0);
      for (int l = 0; l <= col; ++l) acc = acc + W[i + (size_t)l * n] * R[l + (size_t)col * mu];
      row[col] = acc;
    }
    for (int col = 0; col < mu; ++col) W[i + (size_t)col * n] = row[col];
  }
  MOCK_END(c);
}

// ---- local factorisation and solves: banded LU with partial pivoting (natural ordering), one per Sub
namespace {
struct BandLU {
  int n = 0, kl = 0, ku = 0;
  std::vector<K> a;      // dense row-major n x n (test sizes only)
  std::vector<int> piv;
};
std::map<const Sub *, BandLU> g_lu;
}  // namespace

int numfact_device(Sub *s, const HostCSR &A) {
  BandLU &f = g_lu[s];
  const int n = A.n;
  if ((int64_t)n * n > (int64_t)1 << 27) {
    set_error("device stand-in: local matrix of order %d is too large for the dense-band LU of the test harness", n);
    return HPDDM_B200_ERR_ARG;
  }
  f.n = n;
  f.kl = f.ku = 0;
  f.a.assign((size_t)n * n, mk(0.0));
  f.piv.assign(n, 0);
  for (int i = 0; i < n; ++i)
    for (int k = A.ia[i]; k < A.ia[i + 1]; ++k) {
      f.a[(size_t)i * n + A.ja[k]] = A.a[k];
      f.kl = std::max(f.kl, i - A.ja[k]);
      f.ku = std::max(f.ku, A.ja[k] - i);
    }
  for (int k = 0; k < n; ++k) {
    const int rmax = std::min(n - 1, k + f.kl), cmax = std::min(n - 1, k + f.kl + f.ku);
    int p = k;
    for (int i = k + 1; i <= rmax; ++i)
      if (hb_abs(f.a[(size_t)i * n + k]) > hb_abs(f.a[(size_t)p * n + k])) p = i;
    f.piv[k] = p;
    if (hb_abs(f.a[(size_t)p * n + k]) == 0.0) {
      set_error("numfact: zero pivot at row %d (device stand-in)", k);
      return HPDDM_B200_ERR_NUMERIC;
    }
    if (p != k)
      for (int j = k; j <= cmax; ++j) std::swap(f.a[(size_t)k * n + j], f.a[(size_t)p * n + j]);
    for (int i = k + 1; i <= rmax; ++i) {
      const K m = f.a[(size_t)i * n + k] / f.a[(size_t)k * n + k];
      f.a[(size_t)i * n + k] = m;
      if (m == mk(0.0)) continue;
      for (int j = k + 1; j <= cmax; ++j) f.a[(size_t)i * n + j] = f.a[(size_t)i * n + j] - m * f.a[(size_t)k * n + j];
    }
  }
  // the host symbolic phase of the real factorisation (ordering, fronts, work items): its results feed the statistics entry points
  HB_CHECK(symbolic_analyze(A, s->gx, s->gy, s->gz, s->gdof, 64, s->sym));
  s->fac.valid = true;
  s->fac.symmetric = A.symmetric;
  s->t_numfact = 0.0;
  return 0;
}
void free_factor(DeviceFactor &f) {
  for (auto it = g_lu.begin(); it != g_lu.end();)
    if (&it->first->fac == &f) it = g_lu.erase(it);
    else ++it;
  f.valid = false;
}
int sptrsv_prepare(Sub *) { return 0; }
int sptrsv_check(Sub *) { return 0; }
int sptrsv_max_block() { return 8; }
int sptrsv_group(int left) { return left >= 4 ? 4 : (left >= 2 ? 2 : 1); }
// out (=|+=) (scale ? scale .* : ) A^-1 b, mu columns of stride n
int sptrsv_solve(Sub *s, const K *b, K *x, int mu, const double *scale, bool accumulate) {
  auto it = g_lu.find(s);
  if (it == g_lu.end()) {
    set_error("solve: no factorisation (device stand-in)");
    return HPDDM_B200_ERR_STATE;
  }
  const BandLU &f = it->second;
  const int n = f.n;
  std::vector<K> y(n);
  for (int cc = 0; cc < mu; ++cc) {
    for (int i = 0; i < n; ++i) y[i] = b[i + (int64_t)cc * n];
    for (int k = 0; k < n; ++k) {
      if (f.piv[k] != k) std::swap(y[k], y[f.piv[k]]);
      const int rmax = std::min(n - 1, k + f.kl);
      for (int i = k + 1; i <= rmax; ++i) y[i] = y[i] - f.a[(size_t)i * n + k] * y[k];
    }
    for (int k = n - 1; k >= 0; --k) {
      const int cmax = std::min(n - 1, k + f.kl + f.ku);
      K acc = y[k];
      for (int j = k + 1; j <= cmax; ++j) acc = acc - f.a[(size_t)k * n + j] * y[j];
      y[k] = acc / f.a[(size_t)k * n + k];
    }
    for (int i = 0; i < n; ++i) {
      const K v = scale ? scale[i] * y[i] : y[i];
      x[i + (int64_t)cc * n] = accumulate ? x[i + (int64_t)cc * n] + v : v;
    }
  }
  s->ctx->launches++;
  return 0;
}
// dE -> dEinv (N x N, column-major): Gauss-Jordan with partial pivoting
int dense_inverse_device(Ctx *c, int N, const K *dE, K *dEinv) {
  std::vector<K> a(dE, dE + (size_t)N * N);
  for (int j = 0; j < N; ++j)
    for (int i = 0; i < N; ++i) dEinv[i + (size_t)j * N] = mk(i == j ? 1.0 : 0.0);
  for (int k = 0; k < N; ++k) {
    int p = k;
    for (int i = k + 1; i < N; ++i)
      if (hb_abs(a[i + (size_t)k * N]) > hb_abs(a[p + (size_t)k * N])) p = i;
    if (hb_abs(a[p + (size_t)k * N]) == 0.0) {
      set_error("coarse operator is singular (device stand-in)");
      return HPDDM_B200_ERR_NUMERIC;
    }
    for (int j = 0; j < N; ++j) {
      std::swap(a[k + (size_t)j * N], a[p + (size_t)j * N]);
      std::swap(dEinv[k + (size_t)j * N], dEinv[p + (size_t)j * N]);
    }
    const K piv = a[k + (size_t)k * N];
    for (int j = 0; j < N; ++j) {
      a[k + (size_t)j * N] = a[k + (size_t)j * N] / piv;
      dEinv[k + (size_t)j * N] = dEinv[k + (size_t)j * N] / piv;
    }
    for (int i = 0; i < N; ++i) {
      if (i == k) continue;
      const K m = a[i + (size_t)k * N];
      if (m == mk(0.0)) continue;
      for (int j = 0; j < N; ++j) {
        a[i + (size_t)j * N] = a[i + (size_t)j * N] - m * a[k + (size_t)j * N];
        dEinv[i + (size_t)j * N] = dEinv[i + (size_t)j * N] - m * dEinv[k + (size_t)j * N];
      }
    }
  }
  c->launches++;
  return 0;
}

// ---- peer-memory fabric.  One process: off.  Several processes (hpddm_b200_ctx_comm_init_host): the data plane of the stand-in is the
// host program's own all-gather callback (torch.distributed / gloo in the tests), so that the N > 1 host logic of the library -- links
// between subdomains, coarse layout and gather, Krylov reductions, the device drivers in lockstep on all ranks -- runs between real
// processes on a machine without a GPU.  The halo round gathers every process' packed send buffers (k_pack layout) with a small
// directory and copies the segments addressed to the local subdomains into their receive buffers; the unpack is the library's.
namespace {
int host_gather(Ctx *c, const void *send, void *recv, size_t bytes) {
  if (c->host_allgather(send, recv, bytes, c->host_allgather_user) != 0) {
    set_error("device stand-in: the host all-gather callback failed");
    return HPDDM_B200_ERR_NCCL;
  }
  return 0;
}
}  // namespace
int fabric_setup(Ctx *, int) { return 0; }
bool fabric_on(Ctx *c) { return c->nproc > 1 && c->host_allgather != nullptr; }
int fabric_allgather(Ctx *c, K *buf, int count) {
  if (!fabric_on(c)) return 0;
  std::vector<K> mine(buf + (size_t)c->proc_rank * count, buf + (size_t)(c->proc_rank + 1) * count);
  if (count) HB_CHECK(host_gather(c, mine.data(), buf, (size_t)count * sizeof(K)));
  c->launches++;
  return 1;
}
int fabric_allreduce(Ctx *c, double *buf, int count, int op) {  // op 0: sum in rank order, 1: max
  if (!fabric_on(c)) return 0;
  std::vector<double> all((size_t)c->nproc * count);
  if (count) HB_CHECK(host_gather(c, buf, all.data(), (size_t)count * sizeof(double)));
  for (int i = 0; i < count; ++i) {
    double acc = all[i];
    for (int p = 1; p < c->nproc; ++p) acc = op == 0 ? acc + all[(size_t)p * count + i] : std::max(acc, all[(size_t)p * count + i]);
    buf[i] = acc;
  }
  c->launches++;
  return 1;
}
int p2p_halo(Ctx *c, K *const *x, int mu) {
  if (!fabric_on(c)) return 0;
  // directory of this process: [nsub, then per subdomain: grank, nnb, (neighbour rank, entries, offset into the data block) ...]
  std::vector<long long> dir(1, (long long)c->subs.size());
  std::vector<K> data;
  int li = 0;
  for (Sub *s : c->subs) {
    HB_CHECK(k_pack(c, s, mu, x[li++], s->d_send));
    dir.push_back(s->grank);
    dir.push_back((long long)s->nb_rank.size());
    for (size_t i = 0; i < s->nb_rank.size(); ++i) {
      const long long cnt = (long long)(s->nb_ptr[i + 1] - s->nb_ptr[i]) * mu;
      dir.push_back(s->nb_rank[i]);
      dir.push_back(cnt);
      dir.push_back((long long)data.size());
      data.insert(data.end(), s->d_send + (size_t)s->nb_ptr[i] * mu, s->d_send + (size_t)s->nb_ptr[i] * mu + cnt);
    }
  }
  long long sizes[2] = {(long long)dir.size(), (long long)data.size()};
  std::vector<long long> all_sizes((size_t)2 * c->nproc);
  HB_CHECK(host_gather(c, sizes, all_sizes.data(), sizeof(sizes)));
  long long dmax = 1, vmax = 1;
  for (int p = 0; p < c->nproc; ++p) {
    dmax = std::max(dmax, all_sizes[2 * p]);
    vmax = std::max(vmax, all_sizes[2 * p + 1]);
  }
  dir.resize(dmax, 0);
  data.resize(vmax, mk(0.0));
  std::vector<long long> all_dir((size_t)dmax * c->nproc);
  std::vector<K> all_data((size_t)vmax * c->nproc);
  HB_CHECK(host_gather(c, dir.data(), all_dir.data(), (size_t)dmax * sizeof(long long)));
  HB_CHECK(host_gather(c, data.data(), all_data.data(), (size_t)vmax * sizeof(K)));
  for (Sub *s : c->subs)
    for (size_t i = 0; i < s->nb_rank.size(); ++i) {
      const long long want = (long long)(s->nb_ptr[i + 1] - s->nb_ptr[i]) * mu;
      bool found = false;
      for (int p = 0; p < c->nproc && !found; ++p) {
        const long long *d = &all_dir[(size_t)p * dmax];
        size_t at = 1;
        for (long long q = 0; q < d[0] && !found; ++q) {
          const long long grank = d[at], nnb = d[at + 1];
          at += 2;
          for (long long j = 0; j < nnb; ++j, at += 3)
            if (grank == s->nb_rank[i] && d[at] == s->grank) {
              if (d[at + 1] != want) {
                set_error("device stand-in: subdomains %d and %d disagree on the size of their interface", s->grank, (int)grank);
                return HPDDM_B200_ERR_STATE;
              }
              memcpy(s->d_recv + (size_t)s->nb_ptr[i] * mu, &all_data[(size_t)p * vmax + d[at + 2]], (size_t)want * sizeof(K));
              found = true;
            }
        }
      }
      if (!found) {
        set_error("device stand-in: no process holds neighbour %d of subdomain %d", s->nb_rank[i], s->grank);
        return HPDDM_B200_ERR_STATE;
      }
    }
  li = 0;
  for (Sub *s : c->subs) HB_CHECK(k_unpack(c, s, mu, x[li++]));
  return 1;
}
int p2p_check(Ctx *) { return 0; }
void p2p_free(Ctx *) {}
const K *p2p_last_halo_window(Ctx *) { return nullptr; }

}  // namespace hb
