// Device-resident Krylov driver: right-preconditioned restarted GMRES with the reference's
// defaults (include/HPDDM_GMRES.hpp:31-158; include/HPDDM_iterative.hpp:197-212: classical
// Gram-Schmidt, D-weighted inner products, |s_i| / ||b||_D <= tol).  The Krylov basis, the
// Gram-Schmidt products and the solution update stay in HBM; per iteration only i+2 scalars cross
// PCIe.  The small Hessenberg least-squares problem (Givens rotations,
// iterative.hpp:669-710) is solved on the host exactly as the reference does.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "hb_gcrodr.h"
#include "hb_internal.h"

namespace hb {

namespace {

// D-weighted products of the first k basis vectors of every column with that column of w (conjugated on V:
// iterative.hpp:503,517), reduced over the local subdomains and all processes, returned on the host:
// out[nu * k + r] = sum_i d_i conj(V_r[i, nu]) w[i, nu].  V[q] is the block basis of subdomain q (vector r = n_q x mu
// block at V[q] + r * mu * n_q, reference layout v[r] + nu * n), so one column's vectors are mu * n apart.
int dots(Ctx *c, int k, int mu, const std::vector<K *> &V, const std::vector<K *> &w, K *d_T, std::vector<K> &out) {
  HB_CUDA(cudaMemsetAsync(d_T, 0, (size_t)k * mu * sizeof(K), c->stream));
  for (size_t q = 0; q < c->subs.size(); ++q) {
    const size_t n = c->subs[q]->n;
    for (int nu = 0; nu < mu; ++nu) HB_CHECK(k_vdots(c, c->subs[q], k, V[q] + nu * n, (int64_t)mu * n, w[q] + nu * n, d_T + (size_t)nu * k));
  }
  HB_CHECK(nccl_allreduce_sum(c, reinterpret_cast<double *>(d_T), k * mu * KD));
  out.resize((size_t)k * mu);
  HB_CUDA(cudaMemcpyAsync(out.data(), d_T, (size_t)k * mu * sizeof(K), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // namespace

// IterativeMethod::GMRES for mu right-hand sides advancing together (GMRES.hpp:31-158: every column has its own Krylov
// space, Hessenberg matrix and rotations, but the preconditioner and the operator are applied to the n x mu block at once,
// so the factor panels are streamed once per iteration for all columns).  b, x: device pointers per local subdomain,
// column-major n x mu (x holds the initial guess).  hasConverged semantics as in the reference: a converged column keeps
// its dimension for the next solution update and is frozen afterwards (GMRES.hpp:90-92, iterative.hpp:98-103,272-336).
int gmres_device(Ctx *c, const std::vector<const K *> &b, const std::vector<K *> &x, int mu, int correction, int restart, int max_it, double tol,
                 int *iterations, double *rel_residual) {
  const size_t L = c->subs.size();
  const int m = restart;
  std::vector<K *> V(L, nullptr), w(L), z(L), t(L);
  std::vector<const K *> cz(L), ct(L);
  K *d_T = nullptr, *d_h = nullptr;
  auto cleanup = [&]() {
    for (K *p : V) cudaFree(p);
    for (size_t i = 0; i < L; ++i) {
      cudaFree(w[i]);
      cudaFree(z[i]);
      cudaFree(t[i]);
    }
    cudaFree(d_T);
    cudaFree(d_h);
  };
#define KR(call)        \
  do {                  \
    int r__ = (call);   \
    if (r__ < 0) {      \
      cleanup();        \
      return r__;       \
    }                   \
  } while (0)
#define KRC(call)                                                                                 \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      cleanup();                                                                                  \
      return e__ == cudaErrorMemoryAllocation ? HPDDM_B200_ERR_NOMEM : HPDDM_B200_ERR_CUDA;        \
    }                                                                                             \
  } while (0)
  for (size_t i = 0; i < L; ++i) {
    const size_t len = std::max<size_t>((size_t)c->subs[i]->n * mu, 1);
    w[i] = z[i] = t[i] = nullptr;
    KRC(cudaMalloc(&V[i], len * (m + 1) * sizeof(K)));
    KRC(cudaMalloc(&w[i], len * sizeof(K)));
    KRC(cudaMalloc(&z[i], len * sizeof(K)));
    KRC(cudaMalloc(&t[i], len * sizeof(K)));
    cz[i] = z[i];
    ct[i] = t[i];
  }
  KRC(cudaMalloc(&d_T, (size_t)(m + 2) * mu * sizeof(K)));
  KRC(cudaMalloc(&d_h, (size_t)(m + 2) * mu * sizeof(K)));
  std::vector<K> hv;
  auto len_of = [&](size_t q) { return (int64_t)c->subs[q]->n * mu; };
  auto col = [&](K *base, size_t q, int nu) { return base + (size_t)nu * c->subs[q]->n; };
  auto vec = [&](size_t q, int r) { return V[q] + (size_t)r * mu * c->subs[q]->n; };
  // Schwarz::start (schwarz.hpp:496-514): penalised rows + exchange(x)
  for (size_t i = 0; i < L; ++i) {
    Sub *s = c->subs[i];
    KR(k_bc(c, s, mu, b[i], x[i]));
    KR(k_scale(c, s->n, mu, s->d_d, x[i], x[i]));
  }
  KR(halo(c, x.data(), mu));
  // ||b||_D per column, entries of penalised boundary rows divided by HPDDM_PEN (initializeNorm, iterative.hpp:455-468)
  std::vector<double> normb(mu);
  KR(rhs_norms(c, b, mu, normb));
  for (int nu = 0; nu < mu; ++nu)
    if (normb[nu] < 1e-12) normb[nu] = 1.0;  // GMRES.hpp:73
  // checkConvergence (iterative.hpp:98-103): tol > 0 relative to ||b||, tol < 0 absolute
  auto converged = [&](double r, int nu) { return tol > 0.0 ? r / normb[nu] <= tol : r <= -tol; };
  // per column: Hessenberg (m+1) x m, Givens rotations as the reference stores them (iterative.hpp:690-710: cosine in K,
  // sine real), rotated right-hand side sv, hasConverged
  const size_t hs = (size_t)(m + 1) * m;
  std::vector<K> H(hs * mu, mk(0.0)), cs((size_t)m * mu), sv((size_t)(m + 1) * mu), y(m);
  std::vector<double> sn((size_t)m * mu), res(mu, 0.0);
  std::vector<int> conv(mu, -m);
  int j = 1;
  bool done = false;
  while (j <= max_it) {
    // v0 = b - A x
    {
      std::vector<const K *> cx(L);
      for (size_t i = 0; i < L; ++i) cx[i] = x[i];
      KR(gmv_core(c, cx, w, mu));
    }
    for (size_t i = 0; i < L; ++i) {
      KR(k_scal_copy(c, len_of(i), -1.0, w[i], w[i]));
      KR(k_axpy(c, len_of(i), 1.0, b[i], w[i]));
    }
    KR(dots(c, 1, mu, w, w, d_T, hv));
    if (j == 1) {
      bool tiny = false;
      for (int nu = 0; nu < mu; ++nu) tiny = tiny || hb_real(hv[nu]) < 4.930380657631324e-32;  // eps^2 (GMRES.hpp:75)
      if (tiny) {
        j = 0;
        break;
      }
    }
    std::fill(sv.begin(), sv.end(), mk(0.0));
    for (int nu = 0; nu < mu; ++nu) {
      if (conv[nu] > 0) conv[nu] = 0;  // GMRES.hpp:91
      const double beta0 = std::sqrt(hb_real(hv[nu]));
      sv[(size_t)nu * (m + 1)] = mk(beta0);
      for (size_t q = 0; q < L; ++q) KR(k_scal_copy(c, c->subs[q]->n, 1.0 / beta0, col(w[q], q, nu), col(vec(q, 0), q, nu)));
    }
    std::fill(H.begin(), H.end(), mk(0.0));
    int i = 0;
    done = false;
    while (i < m && j <= max_it) {
      std::vector<const K *> vi(L);
      for (size_t q = 0; q < L; ++q) vi[q] = vec(q, i);
      KR(apply_core(c, vi, z, mu, correction));  // z = M^-1 v_i   (GMRES.hpp:116), all columns at once
      KR(gmv_core(c, cz, w, mu));                // w = A z        (GMRES.hpp:117)
      KR(dots(c, i + 1, mu, V, w, d_T, hv));     // classical Gram-Schmidt: all products first
      KRC(cudaMemcpyAsync(d_h, hv.data(), (size_t)(i + 1) * mu * sizeof(K), cudaMemcpyHostToDevice, c->stream));
      for (size_t q = 0; q < L; ++q)
        for (int nu = 0; nu < mu; ++nu)
          KR(k_vupdate(c, c->subs[q], i + 1, col(V[q], q, nu), (int64_t)mu * c->subs[q]->n, d_h + (size_t)nu * (i + 1), -1.0, col(w[q], q, nu)));
      const std::vector<K> hcol(hv);
      KR(dots(c, 1, mu, w, w, d_T, hv));
      for (int nu = 0; nu < mu; ++nu) {
        const double hn = std::sqrt(hb_real(hv[nu]));
        K *Hc = &H[hs * nu + (size_t)i * (m + 1)];
        K *cc = &cs[(size_t)m * nu];
        double *ss = &sn[(size_t)m * nu];
        K *sq = &sv[(size_t)nu * (m + 1)];
        for (int k = 0; k <= i; ++k) Hc[k] = hcol[(size_t)nu * (i + 1) + k];
        Hc[i + 1] = mk(hn);
        if (i < m - 1)
          for (size_t q = 0; q < L; ++q) KR(k_scal_copy(c, c->subs[q]->n, hn == 0.0 ? 1.0 : 1.0 / hn, col(w[q], q, nu), col(vec(q, i + 1), q, nu)));
        for (int k = 0; k < i; ++k) {  // previous rotations (iterative.hpp:690-697)
          const K g = hb_conj(cc[k]) * Hc[k] + ss[k] * Hc[k + 1];
          Hc[k + 1] = cc[k] * Hc[k + 1] - ss[k] * Hc[k];
          Hc[k] = g;
        }
        const double delta = std::hypot(hb_abs(Hc[i]), hb_abs(Hc[i + 1]));  // nrm2 of the two entries (iterative.hpp:701)
        ss[i] = hb_real(Hc[i + 1]) / delta;
        cc[i] = Hc[i] / delta;
        Hc[i] = mk(delta);
        Hc[i + 1] = mk(0.0);
        sq[i + 1] = -ss[i] * sq[i];
        sq[i] = sq[i] * hb_conj(cc[i]);
      }
      ++i;
      bool all = true;
      for (int nu = 0; nu < mu; ++nu) {
        res[nu] = hb_abs(sv[(size_t)nu * (m + 1) + i]);
        if (conv[nu] == -m && converged(res[nu], nu)) conv[nu] = i;  // checkConvergence (iterative.hpp:98-103)
        all = all && conv[nu] != -m;
      }
      if (all) {
        done = true;
        break;
      }
      ++j;
    }
    // updateSol (iterative.hpp:272-336): per column y = H^-1 s over its own dimension, x += M^-1 (V y)
    bool any = false;
    for (size_t q = 0; q < L; ++q) KRC(cudaMemsetAsync(t[q], 0, (size_t)len_of(q) * sizeof(K), c->stream));
    for (int nu = 0; nu < mu; ++nu) {
      int dim = conv[nu] != 0 ? std::abs(conv[nu]) : 0;
      if (!done && conv[nu] == -m) dim = i;  // restart (i == m) or iteration limit (i = iterations of this cycle)
      if (dim == 0) continue;
      any = true;
      const K *Hn = &H[hs * nu];
      const K *sq = &sv[(size_t)nu * (m + 1)];
      for (int k = dim - 1; k >= 0; --k) {
        K acc = sq[k];
        for (int l = k + 1; l < dim; ++l) acc -= Hn[k + (size_t)l * (m + 1)] * y[l];
        y[k] = acc / Hn[k + (size_t)k * (m + 1)];
      }
      KRC(cudaMemcpyAsync(d_h + (size_t)nu * (m + 1), y.data(), dim * sizeof(K), cudaMemcpyHostToDevice, c->stream));
      KRC(cudaStreamSynchronize(c->stream));  // y is reused by the next column
      for (size_t q = 0; q < L; ++q) KR(k_vupdate(c, c->subs[q], dim, col(V[q], q, nu), (int64_t)mu * c->subs[q]->n, d_h + (size_t)nu * (m + 1), 1.0, col(t[q], q, nu)));
    }
    if (any) {
      KR(apply_core(c, ct, z, mu, correction));
      for (size_t q = 0; q < L; ++q)
        for (int nu = 0; nu < mu; ++nu)
          if (conv[nu] != 0) KR(k_axpy(c, c->subs[q]->n, 1.0, col(z[q], q, nu), col(x[q], q, nu)));
    }
    if (done || j > max_it) break;
  }
  KRC(cudaStreamSynchronize(c->stream));
  cleanup();
  *iterations = std::min(j, max_it);
  if (rel_residual)
    for (int nu = 0; nu < mu; ++nu) rel_residual[nu] = res[nu] / normb[nu];
  return 0;
#undef KR
#undef KRC
}


// IterativeMethod::CG (include/HPDDM_CG.hpp:31-168, non-flexible variant): preconditioned conjugate gradient with the
// reference's D-weighted inner products sum_i d_i conj(x_i) y_i, all mu right-hand sides advancing together, convergence
// when ||M^-1 r||_D / ||M^-1 r_0||_D <= tol per column (CG.hpp:60-64,139-140).  r, p, z stay in HBM; 2 mu scalars cross PCIe
// twice per iteration.
int cg_device(Ctx *c, const std::vector<const K *> &b, const std::vector<K *> &x, int mu, int correction, int max_it, double tol, int *iterations,
              double *rel_residual) {
  const size_t L = c->subs.size();
  std::vector<K *> r(L, nullptr), p(L, nullptr), z(L, nullptr);
  std::vector<const K *> cr(L), cp(L), cx(L);
  K *d_dir = nullptr;
  auto cleanup = [&]() {
    for (size_t i = 0; i < L; ++i) {
      cudaFree(r[i]);
      cudaFree(p[i]);
      cudaFree(z[i]);
    }
    cudaFree(d_dir);
  };
#define KR(call)        \
  do {                  \
    int r__ = (call);   \
    if (r__ < 0) {      \
      cleanup();        \
      return r__;       \
    }                   \
  } while (0)
#define KRC(call)                                                                                 \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      cleanup();                                                                                  \
      return e__ == cudaErrorMemoryAllocation ? HPDDM_B200_ERR_NOMEM : HPDDM_B200_ERR_CUDA;        \
    }                                                                                             \
  } while (0)
  auto len_of = [&](size_t q) { return (int64_t)c->subs[q]->n * mu; };
  auto col = [&](K *base, size_t q, int nu) { return base + (size_t)nu * c->subs[q]->n; };
  for (size_t i = 0; i < L; ++i) {
    const size_t len = std::max<size_t>((size_t)len_of(i), 1);
    KRC(cudaMalloc(&r[i], len * sizeof(K)));
    KRC(cudaMalloc(&p[i], len * sizeof(K)));
    KRC(cudaMalloc(&z[i], len * sizeof(K)));
    cr[i] = r[i];
    cp[i] = p[i];
    cx[i] = x[i];
  }
  KRC(cudaMalloc(&d_dir, (size_t)2 * mu * sizeof(K)));
  std::vector<K> hd((size_t)2 * mu);
  // two families of D-weighted products, one reduction (CG.hpp:103,120: Allreduce over 2 mu values)
  auto dots2 = [&](const std::vector<K *> &x0, const std::vector<K *> &y0, const std::vector<K *> *x1, const std::vector<K *> *y1) -> int {
    HB_CUDA(cudaMemsetAsync(d_dir, 0, (size_t)2 * mu * sizeof(K), c->stream));
    for (size_t q = 0; q < L; ++q) {
      HB_CHECK(k_dot(c, c->subs[q], mu, x0[q], y0[q], d_dir));
      if (x1) HB_CHECK(k_dot(c, c->subs[q], mu, (*x1)[q], (*y1)[q], d_dir + mu));
    }
    HB_CHECK(nccl_allreduce_sum(c, reinterpret_cast<double *>(d_dir), 2 * mu * KD));
    HB_CUDA(cudaMemcpyAsync(hd.data(), d_dir, (size_t)2 * mu * sizeof(K), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  };
  // Schwarz::start, r = b - A x, p = M^-1 r (CG.hpp:60-65)
  for (size_t i = 0; i < L; ++i) {
    Sub *s = c->subs[i];
    KR(k_bc(c, s, mu, b[i], x[i]));
    KR(k_scale(c, s->n, mu, s->d_d, x[i], x[i]));
  }
  KR(halo(c, x.data(), mu));
  KR(gmv_core(c, cx, z, mu));
  for (size_t i = 0; i < L; ++i) {
    KR(k_copy(c, len_of(i), b[i], r[i]));
    KR(k_axpy(c, len_of(i), -1.0, z[i], r[i]));
  }
  KR(apply_core(c, cr, p, mu, correction));
  KR(dots2(r, p, &p, &p));                            // (r, D p) and ||p||_D^2       (CG.hpp:66-69,87 at i = 0)
  // rz = (r, D M^-1 r): the reference recomputes it at the top of every iteration (CG.hpp:87) from vectors that have not
  // changed since the end of the previous one (CG.hpp:110), where it is available already -- kept instead of recomputed
  std::vector<double> rz(mu), pAp(mu), res(mu), last(mu, 0.0);
  std::vector<int> conv(mu, -max_it);
  bool tiny = false;
  for (int nu = 0; nu < mu; ++nu) {
    rz[nu] = hb_real(hd[nu]);
    res[nu] = std::sqrt(hb_real(hd[mu + nu]));
    tiny = tiny || hb_real(hd[mu + nu]) < 4.930380657631324e-32;  // eps^2 (CG.hpp:85)
  }
  int i = 0;
  if (!tiny) {
    while (i < max_it) {
      KR(gmv_core(c, cp, z, mu));                     // z = A p                     (CG.hpp:88)
      KR(dots2(z, p, nullptr, nullptr));              // (A p, D p)                  (CG.hpp:90)
      for (int nu = 0; nu < mu; ++nu) pAp[nu] = hb_real(hd[nu]);
      ++i;
      for (int nu = 0; nu < mu; ++nu)
        if (conv[nu] == -max_it) {                    // converged columns are frozen (CG.hpp:99-105)
          const double alpha = rz[nu] / pAp[nu];
          for (size_t q = 0; q < L; ++q) {
            KR(k_axpy(c, c->subs[q]->n, alpha, col(p[q], q, nu), col(x[q], q, nu)));
            KR(k_axpy(c, c->subs[q]->n, -alpha, col(z[q], q, nu), col(r[q], q, nu)));
          }
        }
      KR(apply_core(c, cr, z, mu, correction));       // z = M^-1 r                  (CG.hpp:107)
      KR(dots2(r, z, &z, &z));                        // (r, D z) and (z, D z)       (CG.hpp:110-111)
      for (int nu = 0; nu < mu; ++nu) {
        const double beta = hb_real(hd[nu]) / rz[nu];
        rz[nu] = hb_real(hd[nu]);
        for (size_t q = 0; q < L; ++q) {              // p = z + beta p              (CG.hpp:115)
          KR(k_scal_copy(c, c->subs[q]->n, beta, col(p[q], q, nu), col(p[q], q, nu)));
          KR(k_axpy(c, c->subs[q]->n, 1.0, col(z[q], q, nu), col(p[q], q, nu)));
        }
        last[nu] = std::sqrt(hb_real(hd[mu + nu]));
      }
      bool all = true;
      for (int nu = 0; nu < mu; ++nu) {
        if (conv[nu] == -max_it && last[nu] / res[nu] <= tol) conv[nu] = i;  // checkConvergence<2> (iterative.hpp:98-103)
        all = all && conv[nu] != -max_it;
      }
      if (all) {
        --i;
        break;
      }
    }
  } else
    i = -1;
  ++i;
  KRC(cudaStreamSynchronize(c->stream));
  cleanup();
  *iterations = std::min(i, max_it);
  if (rel_residual)
    for (int nu = 0; nu < mu; ++nu) rel_residual[nu] = res[nu] > 0.0 ? last[nu] / res[nu] : 0.0;
  return 0;
#undef KR
#undef KRC
}


// ------------------------------------------------------------------ block GMRES
namespace {

// small dense helpers on the host, column-major, written against LAPACK's conventions (zpotf2, zlarfg, zgeqr2, zunm2r) because the
// reference's convergence test reads individual entries of the Householder-transformed block residual (see bgmres_device)
bool potrf_upper(int n, K *A, int lda) {  // A = R^H R, R upper in place; false if not positive definite
  for (int j = 0; j < n; ++j) {
    double ajj = hb_real(A[j + (size_t)j * lda]);
    for (int k = 0; k < j; ++k) ajj -= hb_norm(A[k + (size_t)j * lda]);
    if (!(ajj > 0.0)) return false;
    ajj = std::sqrt(ajj);
    A[j + (size_t)j * lda] = mk(ajj);
    for (int i = j + 1; i < n; ++i) {
      K acc = A[j + (size_t)i * lda];
      for (int k = 0; k < j; ++k) acc -= hb_conj(A[k + (size_t)j * lda]) * A[k + (size_t)i * lda];
      A[j + (size_t)i * lda] = acc / ajj;
    }
  }
  return true;
}
void trtri_upper(int n, const K *R, int ldr, K *X) {  // X = R^-1 (n x n, ld n)
  for (int c = 0; c < n; ++c) {
    for (int r = 0; r < n; ++r) X[r + (size_t)c * n] = mk(0.0);
    X[c + (size_t)c * n] = mk(1.0) / R[c + (size_t)c * ldr];
    for (int r = c - 1; r >= 0; --r) {
      K acc = mk(0.0);
      for (int k = r + 1; k <= c; ++k) acc -= R[r + (size_t)k * ldr] * X[k + (size_t)c * n];
      X[r + (size_t)c * n] = acc / R[r + (size_t)r * ldr];
    }
  }
}
void larfg(int n, K &alpha, K *x, K &tau) {  // elementary reflector H = I - tau v v^H, v(0) = 1, H^H (alpha; x) = (beta; 0)
  double xn = 0.0;
  for (int i = 0; i < n - 1; ++i) xn += hb_norm(x[i]);
  xn = std::sqrt(xn);
  const double ar = hb_real(alpha), ai = hb_imag(alpha);
  if (n <= 0 || (xn == 0.0 && ai == 0.0)) {
    tau = mk(0.0);
    return;
  }
  const double beta = -std::copysign(std::sqrt(ar * ar + ai * ai + xn * xn), ar);
  tau = mk((beta - ar) / beta, -ai / beta);
  const K scal = mk(1.0) / (alpha - mk(beta));
  for (int i = 0; i < n - 1; ++i) x[i] = x[i] * scal;
  alpha = mk(beta);
}
// C(0:m, 0:nc) <- (I - t v v^H) C, v = (1; v1)
void larf_left(int m, int nc, const K *v1, K t, K *C, int ldc) {
  if (t == mk(0.0)) return;
  for (int j = 0; j < nc; ++j) {
    K *cj = C + (size_t)j * ldc;
    K w = cj[0];
    for (int r = 1; r < m; ++r) w += hb_conj(v1[r - 1]) * cj[r];
    w = t * w;
    cj[0] -= w;
    for (int r = 1; r < m; ++r) cj[r] -= v1[r - 1] * w;
  }
}
void geqr2(int m, int n, K *A, int lda, K *tau) {
  for (int i = 0; i < std::min(m, n); ++i) {
    larfg(m - i, A[i + (size_t)i * lda], A + i + 1 + (size_t)i * lda, tau[i]);
    if (i < n - 1) larf_left(m - i, n - i - 1, A + i + 1 + (size_t)i * lda, hb_conj(tau[i]), A + i + (size_t)(i + 1) * lda, lda);
  }
}
// C <- Q^H C with the k reflectors stored below the diagonal of A (unm2r "L", "C")
void unm2r_left_c(int m, int nc, int k, const K *A, int lda, const K *tau, K *C, int ldc) {
  for (int i = 0; i < k; ++i) larf_left(m - i, nc, A + i + 1 + (size_t)i * lda, hb_conj(tau[i]), C + i, ldc);
}

}  // namespace

// IterativeMethod::BGMRES (include/HPDDM_GMRES.hpp:160-313) with the reference defaults (iterative.hpp:192-212): right
// preconditioning, block classical Gram-Schmidt (blockOrthogonalization, iterative.hpp:523-556), CholQR of every new block (QR / VR,
// iterative.hpp:560-583,623-640), no deflation of right-hand sides, block Hessenberg reduced by Householder QRs of 2 mu x mu blocks
// (BlockArnoldi, iterative.hpp:714-737), convergence when for every column nu the norm of the first nu + 1 entries of column nu of the
// trailing block of the transformed residual, over ||b_nu||_D, is <= tol (checkBlockConvergence, iterative.hpp:139-146: a partial norm,
// kept as is).  The block basis, the Gram products and the updates stay in HBM: per iteration one block apply, one block SpMV, two
// tall-skinny products, two small reductions to the host.  On a rank-deficient block (CholQR breakdown) the reference restarts with
// GMRES (GMRES.hpp:307-312); so does this.
int bgmres_device(Ctx *c, const std::vector<const K *> &b, const std::vector<K *> &x, int mu, int correction, int restart, int max_it, double tol,
                  int *iterations, double *rel_residual) {
  const size_t L = c->subs.size();
  const int m = restart, ldh = mu * (m + 1);
  if (mu > 8) {
    set_error("solve_bgmres: at most 8 right-hand sides per block (got %d)", mu);
    return HPDDM_B200_ERR_ARG;
  }
  std::vector<K *> V(L, nullptr), z(L, nullptr), t(L, nullptr);
  std::vector<const K *> cz(L), ct(L), cx(L);
  K *d_T = nullptr, *d_R = nullptr;
  auto cleanup = [&]() {
    for (size_t i = 0; i < L; ++i) {
      cudaFree(V[i]);
      cudaFree(z[i]);
      cudaFree(t[i]);
    }
    cudaFree(d_T);
    cudaFree(d_R);
  };
#define KR(call)        \
  do {                  \
    int r__ = (call);   \
    if (r__ < 0) {      \
      cleanup();        \
      return r__;       \
    }                   \
  } while (0)
#define KRC(call)                                                                                 \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      cleanup();                                                                                  \
      return e__ == cudaErrorMemoryAllocation ? HPDDM_B200_ERR_NOMEM : HPDDM_B200_ERR_CUDA;        \
    }                                                                                             \
  } while (0)
  auto len_of = [&](size_t q) { return (int64_t)c->subs[q]->n * mu; };
  auto vec = [&](size_t q, int r) { return V[q] + (size_t)r * mu * c->subs[q]->n; };
  for (size_t i = 0; i < L; ++i) {
    const size_t len = std::max<size_t>((size_t)len_of(i), 1);
    KRC(cudaMalloc(&V[i], len * (m + 1) * sizeof(K)));
    KRC(cudaMalloc(&z[i], len * sizeof(K)));
    KRC(cudaMalloc(&t[i], len * sizeof(K)));
    cz[i] = z[i];
    ct[i] = t[i];
    cx[i] = x[i];
  }
  KRC(cudaMalloc(&d_T, (size_t)ldh * mu * sizeof(K)));
  KRC(cudaMalloc(&d_R, (size_t)64 * sizeof(K)));
  std::vector<K> hbuf;
  // T (k x mu, ld k) = sum over subdomains / processes of X^H D W, X = n x k block (ld n); left on the device in d_T and copied to the host
  auto gram = [&](int k, const std::vector<K *> &X, const std::vector<K *> &W) -> int {
    HB_CUDA(cudaMemsetAsync(d_T, 0, (size_t)k * mu * sizeof(K), c->stream));
    for (size_t q = 0; q < L; ++q) HB_CHECK(k_zt_raw(c, c->subs[q]->n, k, X[q], c->subs[q]->d_d, mu, W[q], d_T, k));
    HB_CHECK(nccl_allreduce_sum(c, reinterpret_cast<double *>(d_T), k * mu * KD));
    hbuf.resize((size_t)k * mu);
    HB_CUDA(cudaMemcpyAsync(hbuf.data(), d_T, (size_t)k * mu * sizeof(K), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  };
  // CholQR of the block W (IterativeMethod::QR, HPDDM_QR_CHOLQR): R (mu x mu upper, ld mu) on the host; W <- W R^-1 if update.
  // returns 1 on breakdown (Gram matrix not positive definite)
  std::vector<K> R((size_t)mu * mu), Rinv((size_t)mu * mu);
  auto cholqr = [&](const std::vector<K *> &W, bool update) -> int {
    HB_CHECK(gram(mu, W, W));
    R = hbuf;
    if (!potrf_upper(mu, R.data(), mu)) return 1;
    for (int cc = 0; cc < mu; ++cc)
      for (int r = cc + 1; r < mu; ++r) R[r + (size_t)cc * mu] = mk(0.0);
    if (update) {
      trtri_upper(mu, R.data(), mu, Rinv.data());
      HB_CUDA(cudaMemcpyAsync(d_R, Rinv.data(), (size_t)mu * mu * sizeof(K), cudaMemcpyHostToDevice, c->stream));
      for (size_t q = 0; q < L; ++q) HB_CHECK(k_rmul_upper(c, c->subs[q]->n, mu, d_R, W[q]));
    }
    return 0;
  };
  // Schwarz::start
  for (size_t i = 0; i < L; ++i) {
    Sub *s = c->subs[i];
    KR(k_bc(c, s, mu, b[i], x[i]));
    KR(k_scale(c, s->n, mu, s->d_d, x[i], x[i]));
  }
  KR(halo(c, x.data(), mu));
  std::vector<double> normb(mu), last(mu, 0.0);
  KR(rhs_norms(c, b, mu, normb));  // penalised boundary rows / HPDDM_PEN (initializeNorm, iterative.hpp:455-468)
  for (int nu = 0; nu < mu; ++nu)
    if (normb[nu] < 1e-12) normb[nu] = 1.0;
  std::vector<K> H((size_t)ldh * m * mu), tau((size_t)m * mu), sv((size_t)ldh * mu), Hc((size_t)ldh * mu), Y((size_t)ldh * mu);
  int j = 1, dim = 0;
  bool done = false, breakdown = false;
  while (j <= max_it) {
    // block residual V_0 R_0 = b - A x
    std::vector<K *> v0(L), w(L);
    for (size_t q = 0; q < L; ++q) v0[q] = vec(q, 0);
    KR(gmv_core(c, cx, v0, mu));
    for (size_t q = 0; q < L; ++q) {
      KR(k_scal_copy(c, len_of(q), -1.0, v0[q], v0[q]));
      KR(k_axpy(c, len_of(q), 1.0, b[q], v0[q]));
    }
    int rc = cholqr(v0, true);
    if (rc < 0) KR(rc);
    if (rc == 1) {
      breakdown = true;
      break;
    }
    std::fill(H.begin(), H.end(), mk(0.0));
    std::fill(tau.begin(), tau.end(), mk(0.0));
    std::fill(sv.begin(), sv.end(), mk(0.0));
    for (int cc = 0; cc < mu; ++cc)
      for (int r = 0; r <= cc; ++r) sv[r + (size_t)cc * ldh] = R[r + (size_t)cc * mu];
    int i = 0;
    done = false;
    while (i < m && j <= max_it) {
      std::vector<const K *> vi(L);
      for (size_t q = 0; q < L; ++q) {
        vi[q] = vec(q, i);
        w[q] = vec(q, i + 1);
      }
      KR(apply_core(c, vi, z, mu, correction));  // GMRES.hpp:246
      KR(gmv_core(c, cz, w, mu));                // GMRES.hpp:247
      // block classical Gram-Schmidt: all products first, then one update (iterative.hpp:547-555)
      const int k = (i + 1) * mu;
      KR(gram(k, V, w));
      std::fill(Hc.begin(), Hc.end(), mk(0.0));
      for (int cc = 0; cc < mu; ++cc)
        for (int r = 0; r < k; ++r) Hc[r + (size_t)cc * ldh] = hbuf[r + (size_t)cc * k];
      for (size_t q = 0; q < L; ++q) KR(k_vupdate_blk(c, c->subs[q]->n, k, mu, V[q], d_T, k, -1.0, w[q]));
      rc = cholqr(w, i < m - 1);
      if (rc < 0) KR(rc);
      if (rc == 1) {
        breakdown = true;
        break;
      }
      for (int cc = 0; cc < mu; ++cc)
        for (int r = 0; r <= cc; ++r) Hc[k + r + (size_t)cc * ldh] = R[r + (size_t)cc * mu];
      // Householder reduction of the block Hessenberg matrix (BlockArnoldi, iterative.hpp:727-729)
      for (int p = 0; p < i; ++p) unm2r_left_c(2 * mu, mu, mu, &H[(size_t)p * mu + (size_t)p * mu * ldh], ldh, &tau[(size_t)p * mu], &Hc[(size_t)p * mu], ldh);
      geqr2(2 * mu, mu, &Hc[(size_t)i * mu], ldh, &tau[(size_t)i * mu]);
      std::copy(Hc.begin(), Hc.end(), H.begin() + (size_t)i * mu * ldh);
      unm2r_left_c(2 * mu, mu, mu, &H[(size_t)i * mu + (size_t)i * mu * ldh], ldh, &tau[(size_t)i * mu], &sv[(size_t)i * mu], ldh);
      ++i;
      bool all = true;
      for (int nu = 0; nu < mu; ++nu) {  // checkBlockConvergence<1>, t <= 1 branch (iterative.hpp:139-146)
        double nrm = 0.0;
        for (int r = 0; r <= nu; ++r) nrm += hb_norm(sv[(size_t)i * mu + r + (size_t)nu * ldh]);
        last[nu] = std::sqrt(nrm);
        all = all && (tol > 0.0 ? last[nu] / normb[nu] <= tol : last[nu] <= -tol);  // checkBlockConvergence (iterative.hpp:139-146)
      }
      if (all) {
        dim = mu * i;
        done = true;
        break;
      }
      ++j;
    }
    if (breakdown) break;
    if (!done) dim = mu * i;  // restart (i == m) or iteration limit
    if (dim > 0) {            // updateSol (iterative.hpp:272-336): trtrs, gemm, one apply
      for (int cc = 0; cc < mu; ++cc)
        for (int r = dim - 1; r >= 0; --r) {
          K acc = sv[r + (size_t)cc * ldh];
          for (int l = r + 1; l < dim; ++l) acc -= H[r + (size_t)l * ldh] * Y[l + (size_t)cc * dim];
          Y[r + (size_t)cc * dim] = acc / H[r + (size_t)r * ldh];
        }
      KRC(cudaMemcpyAsync(d_T, Y.data(), (size_t)dim * mu * sizeof(K), cudaMemcpyHostToDevice, c->stream));
      for (size_t q = 0; q < L; ++q) {
        KRC(cudaMemsetAsync(t[q], 0, (size_t)len_of(q) * sizeof(K), c->stream));
        KR(k_vupdate_blk(c, c->subs[q]->n, dim, mu, V[q], d_T, dim, 1.0, t[q]));
      }
      KR(apply_core(c, ct, z, mu, correction));
      for (size_t q = 0; q < L; ++q) KR(k_axpy(c, len_of(q), 1.0, z[q], x[q]));
      KRC(cudaStreamSynchronize(c->stream));  // Y (pageable) is rewritten by the next cycle
    }
    if (done || j > max_it) break;
  }
  KRC(cudaStreamSynchronize(c->stream));
  cleanup();
  if (breakdown) {
    if (getenv("HPDDM_B200_DEBUG")) fprintf(stderr, "[hpddm_b200] BGMRES: rank-deficient block, continuing with GMRES (GMRES.hpp:307-312)\n");
    return gmres_device(c, b, x, mu, correction, restart, max_it, tol, iterations, rel_residual);
  }
  *iterations = std::min(j, max_it);
  if (rel_residual)
    for (int nu = 0; nu < mu; ++nu) rel_residual[nu] = last[nu] / normb[nu];
  return 0;
#undef KR
#undef KRC
}


// ------------------------------------------------------------------ GCRO-DR
// IterativeMethod::GCRODR (include/HPDDM_GCRODR.hpp:35-444).  The driver is hb_gcrodr.cpp (host logic: Hessenberg matrices, rotations,
// harmonic Ritz problems -- checked on the CPU against goldens of the unmodified reference through tests/native/gcrodr_host.cpp); this is
// its vector space: Krylov basis, recycled pair (U, C) and every product stay in HBM, built from the kernels of the drivers above.
namespace {

struct RecycledDev {  // Ctx::recycled: the pair survives between solves like the reference's A.storage() (HPDDM_option.hpp:445-454)
  gcro::Recycled r;
  std::vector<int> n;  // subdomain sizes it was built for
};

struct DeviceBackend : gcro::Backend {
  Ctx *c;
  int mu, correction, cap;
  K *d_T = nullptr, *d_h = nullptr;
  RecycledDev *rec;
  DeviceBackend(Ctx *c_, int mu_, int correction_, int cap_, RecycledDev *rec_) : c(c_), mu(mu_), correction(correction_), cap(cap_), rec(rec_) {}
  ~DeviceBackend() override {
    cudaFree(d_T);
    cudaFree(d_h);
  }
  int init() {  // staging for products / coefficients: cap basis vectors (non-block driver) or cap blocks of mu columns (block driver)
    HB_CUDA(cudaMalloc(&d_T, (size_t)cap * mu * mu * sizeof(K)));
    HB_CUDA(cudaMalloc(&d_h, (size_t)cap * mu * mu * sizeof(K)));
    return 0;
  }
  size_t subs() const override { return c->subs.size(); }
  int64_t rows(size_t q) const override { return c->subs[q]->n; }
  K *colp(const gcro::Vec &v, size_t q, int nu) const { return v[q] + (size_t)nu * c->subs[q]->n; }
  int alloc(gcro::Vec &v, int blocks) override {
    v.assign(c->subs.size(), nullptr);
    for (size_t q = 0; q < v.size(); ++q) {
      const size_t bytes = std::max<size_t>((size_t)c->subs[q]->n * mu * blocks, 1) * sizeof(K);
      cudaError_t e = cudaMalloc(&v[q], bytes);
      if (e == cudaSuccess) e = cudaMemsetAsync(v[q], 0, bytes, c->stream);
      if (e != cudaSuccess) {
        set_error("CUDA error %s while allocating %zu bytes of Krylov vectors (GCRO-DR)", cudaGetErrorString(e), bytes);
        release(v);
        return e == cudaErrorMemoryAllocation ? HPDDM_B200_ERR_NOMEM : HPDDM_B200_ERR_CUDA;
      }
    }
    return 0;
  }
  void release(gcro::Vec &v) override {
    if (v.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (K *p : v) cudaFree(p);
    v.clear();
  }
  gcro::Recycled &recycled() override { return rec->r; }
  int start(const gcro::Vec &b, const gcro::Vec &x) override {  // Schwarz::start (schwarz.hpp:496-514): penalised rows + exchange(x)
    for (size_t q = 0; q < c->subs.size(); ++q) {
      Sub *s = c->subs[q];
      HB_CHECK(k_bc(c, s, mu, b[q], x[q]));
      HB_CHECK(k_scale(c, s->n, mu, s->d_d, x[q], x[q]));
    }
    return halo(c, x.data(), mu);
  }
  int rhs_norms(const gcro::Vec &b, std::vector<double> &out) override {
    const std::vector<const K *> cb(b.begin(), b.end());
    out.resize(mu);
    return hb::rhs_norms(c, cb, mu, out);
  }
  int apply(const gcro::Vec &in, const gcro::Vec &out) override {
    const std::vector<const K *> cin(in.begin(), in.end());
    return apply_core(c, cin, out, mu, correction);
  }
  int gmv(const gcro::Vec &in, const gcro::Vec &out) override {
    const std::vector<const K *> cin(in.begin(), in.end());
    return gmv_core(c, cin, out, mu);
  }
  int dots(int count, const gcro::Vec &basis, const gcro::Vec &w, std::vector<K> &out) override {
    if (count > cap) {
      set_error("GCRO-DR: %d products exceed the staging capacity %d", count, cap);
      return HPDDM_B200_ERR_STATE;
    }
    return hb::dots(c, count, mu, basis, w, d_T, out);
  }
  int combine_col(int nu, int count, const gcro::Vec &basis, const K *coef, double alpha, const gcro::Vec &w) override {
    if (count <= 0) return 0;
    if (count > cap) {
      set_error("GCRO-DR: %d coefficients exceed the staging capacity %d", count, cap);
      return HPDDM_B200_ERR_STATE;
    }
    // the driver reuses `coef` right after this call: wait for the copy (a pageable source is staged before cudaMemcpyAsync returns, but a
    // heap block that shares a page with a range the caller pinned would make the copy truly asynchronous); the kernels that read d_h are
    // ordered behind it on the context's stream.  A few microseconds next to a preconditioner apply.
    HB_CUDA(cudaMemcpyAsync(d_h, coef, (size_t)count * sizeof(K), cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t q = 0; q < c->subs.size(); ++q)
      HB_CHECK(k_vupdate(c, c->subs[q], count, colp(basis, q, nu), (int64_t)mu * c->subs[q]->n, d_h, alpha, colp(w, q, nu)));
    return 0;
  }
  int scal_col(int nu, double a, const gcro::Vec &in, const gcro::Vec &out) override {
    for (size_t q = 0; q < c->subs.size(); ++q) HB_CHECK(k_scal_copy(c, c->subs[q]->n, a, colp(in, q, nu), colp(out, q, nu)));
    return 0;
  }
  int axpy_col(int nu, double a, const gcro::Vec &in, const gcro::Vec &out) override {
    for (size_t q = 0; q < c->subs.size(); ++q) HB_CHECK(k_axpy(c, c->subs[q]->n, a, colp(in, q, nu), colp(out, q, nu)));
    return 0;
  }
  int zero_col(int nu, const gcro::Vec &out) override {
    for (size_t q = 0; q < c->subs.size(); ++q)
      if (c->subs[q]->n) HB_CUDA(cudaMemsetAsync(colp(out, q, nu), 0, (size_t)c->subs[q]->n * sizeof(K), c->stream));
    return 0;
  }
  // block products (BGCRO-DR): the tall-skinny kernels of the BGMRES driver; a basis of `count` blocks is an n x (count mu) matrix with ld n
  int gram(int count, const gcro::Vec &basis, const gcro::Vec &w, std::vector<K> &out) override {
    const int rows = count * mu;
    if (count > cap) {
      set_error("BGCRO-DR: %d blocks exceed the staging capacity %d", count, cap);
      return HPDDM_B200_ERR_STATE;
    }
    HB_CUDA(cudaMemsetAsync(d_T, 0, (size_t)rows * mu * sizeof(K), c->stream));
    for (size_t q = 0; q < c->subs.size(); ++q) HB_CHECK(k_zt_raw(c, c->subs[q]->n, rows, basis[q], c->subs[q]->d_d, mu, w[q], d_T, rows));
    HB_CHECK(nccl_allreduce_sum(c, reinterpret_cast<double *>(d_T), rows * mu * KD));
    out.resize((size_t)rows * mu);
    HB_CUDA(cudaMemcpyAsync(out.data(), d_T, (size_t)rows * mu * sizeof(K), cudaMemcpyDeviceToHost, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  }
  int combine_blk(int count, const gcro::Vec &basis, const K *coef, double alpha, const gcro::Vec &w) override {
    if (count <= 0) return 0;
    const int rows = count * mu;
    if (count > cap) {
      set_error("BGCRO-DR: %d blocks exceed the staging capacity %d", count, cap);
      return HPDDM_B200_ERR_STATE;
    }
    HB_CUDA(cudaMemcpyAsync(d_h, coef, (size_t)rows * mu * sizeof(K), cudaMemcpyHostToDevice, c->stream));
    HB_CUDA(cudaStreamSynchronize(c->stream));  // `coef` is reused by the driver (see combine_col)
    for (size_t q = 0; q < c->subs.size(); ++q) HB_CHECK(k_vupdate_blk(c, c->subs[q]->n, rows, mu, basis[q], d_h, rows, alpha, w[q]));
    return 0;
  }
};

}  // namespace

void gcrodr_release(Ctx *c) {
  RecycledDev *rd = static_cast<RecycledDev *>(c->recycled);
  if (!rd) return;
  cudaStreamSynchronize(c->stream);
  for (K *p : rd->r.U) cudaFree(p);
  for (K *p : rd->r.C) cudaFree(p);
  delete rd;
  c->recycled = nullptr;
}

// block = false: IterativeMethod::GCRODR (one Krylov space and one recycled pair per right-hand side); block = true: IterativeMethod::BGCRODR
// (GCRODR.hpp:445-907: one block Krylov space and one pair of mu k columns for all right-hand sides)
int gcrodr_device(Ctx *c, bool block, const std::vector<const K *> &b, const std::vector<K *> &x, int mu, int correction, int restart, int recycle, int target,
                  int strategy, int same_system, int max_it, double tol, int *iterations, double *rel_residual) {
  if (recycle <= 0)  // GCRODR.hpp:50-55, 460-465
    return block ? bgmres_device(c, b, x, mu, correction, std::min(restart, max_it), max_it, tol, iterations, rel_residual)
                 : gmres_device(c, b, x, mu, correction, restart, max_it, tol, iterations, rel_residual);
  if (block && (size_t)(std::min(restart, max_it) + 1) * mu * 4 * sizeof(K) > 48 * 1024) {
    set_error("solve_bgcrodr: (restart + 1) x mu = %d x %d exceeds the shared-memory staging of the block update kernel", std::min(restart, max_it) + 1, mu);
    return HPDDM_B200_ERR_ARG;
  }
  RecycledDev *rd = static_cast<RecycledDev *>(c->recycled);
  std::vector<int> sizes;
  for (Sub *s : c->subs) sizes.push_back(s->n);
  if (rd && rd->n != sizes) {  // the decomposition changed under the stored pair
    gcrodr_release(c);
    rd = nullptr;
  }
  if (!rd) {
    rd = new RecycledDev();
    rd->n = sizes;
    c->recycled = rd;
  }
  const int m = std::min(restart, max_it);
  DeviceBackend be(c, mu, correction, m + 2, rd);
  HB_CHECK(be.init());
  gcro::Params p;
  p.mu = mu;
  p.restart = restart;
  p.recycle = recycle;
  p.max_it = max_it;
  p.tol = tol;
  p.target = target;
  p.strategy = strategy;
  p.same_system = same_system;
  gcro::Vec bv(b.size());
  for (size_t q = 0; q < b.size(); ++q) bv[q] = const_cast<K *>(b[q]);  // the driver only reads b
  const int rc = block ? gcro::run_block(be, bv, x, p, iterations, rel_residual) : gcro::run(be, bv, x, p, iterations, rel_residual);
  cudaStreamSynchronize(c->stream);
  if (rc == gcro::ERR_EIGENSOLVER) {
    set_error("solve_gcrodr: the harmonic Ritz eigenproblem of a cycle could not be solved (QR iteration did not converge or singular pencil)");
    return HPDDM_B200_ERR_NUMERIC;
  }
  return rc;
}

}  // namespace hb

using namespace hb;

// stage host vectors (if any), run `method`, copy the solution back
template <class Method>
static int krylov_entry(Ctx *c, const K *const *b, K *const *x, int mu, int where, Method method) {
  HB_CHECK(check_ready(c, std::max(mu, 1)));
  const size_t L = c->subs.size();
  // vectors live in private device buffers for the whole solve (d_in / d_out are used by nothing else here)
  std::vector<K *> bd(L, nullptr), xd(L, nullptr);
  auto release = [&]() {
    if (where != HPDDM_B200_HOST) return;
    for (size_t i = 0; i < L; ++i) {
      cudaFree(bd[i]);
      cudaFree(xd[i]);
    }
  };
  for (size_t i = 0; i < L; ++i) {
    const size_t bytes = (size_t)c->subs[i]->n * mu * sizeof(K);
    if (where == HPDDM_B200_HOST) {
      cudaError_t e = cudaMalloc(&bd[i], std::max<size_t>(bytes, sizeof(K)));
      if (e == cudaSuccess) e = cudaMalloc(&xd[i], std::max<size_t>(bytes, sizeof(K)));
      if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(bd[i], b[i], bytes, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(xd[i], x[i], bytes, cudaMemcpyHostToDevice, c->stream);
      if (e != cudaSuccess) {
        set_error("CUDA error %s while staging the right-hand sides of a Krylov solve", cudaGetErrorString(e));
        cudaStreamSynchronize(c->stream);
        release();
        return e == cudaErrorMemoryAllocation ? HPDDM_B200_ERR_NOMEM : HPDDM_B200_ERR_CUDA;
      }
    } else {
      bd[i] = const_cast<K *>(b[i]);
      xd[i] = x[i];
    }
  }
  std::vector<const K *> bc(bd.begin(), bd.end());
  const int rc = method(bc, xd);  // all columns advance together
  if (where == HPDDM_B200_HOST) {
    if (rc == 0)
      for (size_t i = 0; i < L; ++i) cudaMemcpyAsync(x[i], xd[i], (size_t)c->subs[i]->n * mu * sizeof(K), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    release();
  }
  if (rc == 0) {  // DEVICE-pointer callers included: a peer-memory collective or a persistent sweep that gave up waiting must not pass silently
    cudaStreamSynchronize(c->stream);
    for (Sub *s : c->subs) HB_CHECK(sptrsv_check(s));
    HB_CHECK(p2p_check(c));
  }
  return rc;
}

extern "C" int HB_API(solve)(hb_ctx_t *ctx, const K *const *b, K *const *x, int mu, int correction, int restart, int max_it, double tol,
                                int where, int *iterations, double *rel_residual) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !iterations || restart < 1 || max_it < 1 || mu < 1) {
    set_error("solve: bad arguments");
    return HPDDM_B200_ERR_ARG;
  }
  *iterations = 0;
  return krylov_entry(c, b, x, mu, where, [&](const std::vector<const K *> &bd, const std::vector<K *> &xd) {
    return gmres_device(c, bd, xd, mu, correction, restart, max_it, tol, iterations, rel_residual);
  });
}

extern "C" int HB_API(solve_cg)(hb_ctx_t *ctx, const K *const *b, K *const *x, int mu, int correction, int max_it, double tol, int where, int *iterations,
                                   double *rel_residual) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !iterations || max_it < 1 || mu < 1) {
    set_error("solve_cg: bad arguments");
    return HPDDM_B200_ERR_ARG;
  }
  *iterations = 0;
  // the reference runs CG only for symmetric preconditioners -- SORAS / ASM / none without a deflated correction -- and
  // falls back to GMRES otherwise (CG.hpp:41-44)
  bool symmetric = correction != HPDDM_B200_CORRECTION_DEFLATED;
  for (Sub *s : c->subs) symmetric = symmetric && (s->prcndtnr == HPDDM_B200_PRCNDTNR_OS || s->prcndtnr == HPDDM_B200_PRCNDTNR_SY || s->prcndtnr == HPDDM_B200_PRCNDTNR_NO);
  return krylov_entry(c, b, x, mu, where, [&](const std::vector<const K *> &bd, const std::vector<K *> &xd) {
    if (!symmetric) return gmres_device(c, bd, xd, mu, correction, 40, max_it, tol, iterations, rel_residual);
    return cg_device(c, bd, xd, mu, correction, max_it, tol, iterations, rel_residual);
  });
}

extern "C" int HB_API(solve_bgmres)(hb_ctx_t *ctx, const K *const *b, K *const *x, int mu, int correction, int restart, int max_it, double tol, int where,
                                       int *iterations, double *rel_residual) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !iterations || restart < 1 || max_it < 1 || mu < 1) {
    set_error("solve_bgmres: bad arguments");
    return HPDDM_B200_ERR_ARG;
  }
  *iterations = 0;
  return krylov_entry(c, b, x, mu, where, [&](const std::vector<const K *> &bd, const std::vector<K *> &xd) {
    return bgmres_device(c, bd, xd, mu, correction, std::min(restart, max_it), max_it, tol, iterations, rel_residual);
  });
}

extern "C" int HB_API(solve_gcrodr)(hb_ctx_t *ctx, const K *const *b, K *const *x, int mu, int correction, int restart, int recycle, int recycle_target,
                                       int recycle_strategy, int recycle_same_system, int max_it, double tol, int where, int *iterations,
                                       double *rel_residual) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !iterations || restart < 1 || max_it < 1 || mu < 1 || recycle_target < 0 || recycle_target > 5 || recycle_strategy < 0 || recycle_strategy > 1 ||
      recycle_same_system < 0 || recycle_same_system > 2) {
    set_error("solve_gcrodr: bad arguments");
    return HPDDM_B200_ERR_ARG;
  }
  *iterations = 0;
  return krylov_entry(c, b, x, mu, where, [&](const std::vector<const K *> &bd, const std::vector<K *> &xd) {
    return gcrodr_device(c, false, bd, xd, mu, correction, restart, recycle, recycle_target, recycle_strategy, recycle_same_system, max_it, tol, iterations, rel_residual);
  });
}

extern "C" int HB_API(solve_bgcrodr)(hb_ctx_t *ctx, const K *const *b, K *const *x, int mu, int correction, int restart, int recycle, int recycle_target,
                                        int recycle_strategy, int recycle_same_system, int max_it, double tol, int where, int *iterations,
                                        double *rel_residual) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c || !iterations || restart < 1 || max_it < 1 || mu < 1 || recycle_target < 0 || recycle_target > 5 || recycle_strategy < 0 || recycle_strategy > 1 ||
      recycle_same_system < 0 || recycle_same_system > 2) {
    set_error("solve_bgcrodr: bad arguments");
    return HPDDM_B200_ERR_ARG;
  }
  *iterations = 0;
  return krylov_entry(c, b, x, mu, where, [&](const std::vector<const K *> &bd, const std::vector<K *> &xd) {
    return gcrodr_device(c, true, bd, xd, mu, correction, restart, recycle, recycle_target, recycle_strategy, recycle_same_system, max_it, tol, iterations, rel_residual);
  });
}

extern "C" int HB_API(recycle_dim)(hb_ctx_t *ctx) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  return c && c->recycled ? static_cast<RecycledDev *>(c->recycled)->r.k : 0;
}

extern "C" int HB_API(recycle_destroy)(hb_ctx_t *ctx) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!c) return HPDDM_B200_ERR_ARG;
  cudaSetDevice(c->device);
  gcrodr_release(c);
  return 0;
}
