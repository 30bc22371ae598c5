// Device-resident Krylov driver: right-preconditioned restarted GMRES with the reference's
// defaults (include/HPDDM_GMRES.hpp:31-158; include/HPDDM_iterative.hpp:197-212: classical
// Gram-Schmidt, D-weighted inner products, |s_i| / ||b||_D <= tol).  The Krylov basis, the
// Gram-Schmidt products and the solution update stay in HBM; per iteration only i+2 scalars cross
// PCIe.  The small Hessenberg least-squares problem (Givens rotations,
// iterative.hpp:669-710) is solved on the host exactly as the reference does.
#include <cmath>
#include <vector>

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
  // ||b||_D per column (iterative.hpp:455-468)
  std::vector<double> normb(mu);
  {
    std::vector<K *> bb(L);
    for (size_t i = 0; i < L; ++i) bb[i] = const_cast<K *>(b[i]);
    KR(dots(c, 1, mu, bb, bb, d_T, hv));  // (a single "vector" per column: stride irrelevant)
    for (int nu = 0; nu < mu; ++nu) {
      normb[nu] = std::sqrt(hb_real(hv[nu]));
      if (normb[nu] < 1e-12) normb[nu] = 1.0;
    }
  }
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
        if (conv[nu] == -m && res[nu] / normb[nu] <= tol) conv[nu] = i;  // checkConvergence (iterative.hpp:98-103)
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

}  // namespace hb

using namespace hb;

// stage host vectors (if any), run `method`, copy the solution back
template <class Method>
static int krylov_entry(Ctx *c, const K *const *b, K *const *x, int mu, int where, Method method) {
  HB_CHECK(check_ready(c, std::max(mu, 1)));
  const size_t L = c->subs.size();
  // vectors live in private device buffers for the whole solve (d_in / d_out are used by nothing else here)
  std::vector<K *> bd(L, nullptr), xd(L, nullptr);
  for (size_t i = 0; i < L; ++i) {
    const size_t bytes = std::max<size_t>((size_t)c->subs[i]->n * mu, 1) * sizeof(K);
    if (where == HPDDM_B200_HOST) {
      HB_CUDA(cudaMalloc(&bd[i], bytes));
      HB_CUDA(cudaMalloc(&xd[i], bytes));
      HB_CUDA(cudaMemcpyAsync(bd[i], b[i], bytes, cudaMemcpyHostToDevice, c->stream));
      HB_CUDA(cudaMemcpyAsync(xd[i], x[i], bytes, cudaMemcpyHostToDevice, c->stream));
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
    for (size_t i = 0; i < L; ++i) {
      cudaFree(bd[i]);
      cudaFree(xd[i]);
    }
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
