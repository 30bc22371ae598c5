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

// D-weighted dots of k basis vectors (per subdomain V_s, n_s x k) with w, reduced over the local
// subdomains and all processes, returned on the host
// (conjugated on V: iterative.hpp:503,517)
int dots(Ctx *c, int k, const std::vector<K *> &V, const std::vector<K *> &w, K *d_T, std::vector<K> &out) {
  HB_CUDA(cudaMemsetAsync(d_T, 0, k * sizeof(K), c->stream));
  for (size_t i = 0; i < c->subs.size(); ++i) HB_CHECK(k_vdots(c, c->subs[i], k, V[i], w[i], d_T));
  HB_CHECK(nccl_allreduce_sum(c, reinterpret_cast<double *>(d_T), k * KD));
  out.resize(k);
  HB_CUDA(cudaMemcpyAsync(out.data(), d_T, k * sizeof(K), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

}  // namespace

// one right-hand side; b, x: device pointers per local subdomain (x holds the initial guess)
int gmres_device(Ctx *c, const std::vector<const K *> &b, const std::vector<K *> &x, int correction, int restart, int max_it, double tol,
                 int *iterations, double *rel_residual) {
  const size_t L = c->subs.size();
  const int m = restart;
  std::vector<K *> V(L, nullptr), w(L), z(L), t(L);
  std::vector<const K *> cz(L), cw(L);
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
      return HPDDM_B200_ERR_CUDA;                                                                 \
    }                                                                                             \
  } while (0)
  for (size_t i = 0; i < L; ++i) {
    const size_t n = std::max<size_t>(c->subs[i]->n, 1);
    w[i] = z[i] = t[i] = nullptr;
    KRC(cudaMalloc(&V[i], n * (m + 1) * sizeof(K)));
    KRC(cudaMalloc(&w[i], n * sizeof(K)));
    KRC(cudaMalloc(&z[i], n * sizeof(K)));
    KRC(cudaMalloc(&t[i], n * sizeof(K)));
    cz[i] = z[i];
    cw[i] = w[i];
  }
  KRC(cudaMalloc(&d_T, (m + 2) * sizeof(K)));
  KRC(cudaMalloc(&d_h, (m + 2) * sizeof(K)));
  std::vector<K> hv;
  // Schwarz::start (schwarz.hpp:496-514): penalised rows + exchange(x)
  for (size_t i = 0; i < L; ++i) {
    Sub *s = c->subs[i];
    KR(k_bc(c, s, 1, b[i], x[i]));
    KR(k_scale(c, s->n, 1, s->d_d, x[i], x[i]));
  }
  KR(halo(c, x.data(), 1));
  // ||b||_D (iterative.hpp:455-468)
  {
    std::vector<K *> bb(L);
    for (size_t i = 0; i < L; ++i) bb[i] = const_cast<K *>(b[i]);
    KR(dots(c, 1, bb, bb, d_T, hv));
  }
  double normb = std::sqrt(hb_real(hv[0]));
  if (normb < 1e-12) normb = 1.0;
  // Givens rotations as the reference stores them (iterative.hpp:690-710): cosine in K, sine real
  std::vector<K> H((size_t)(m + 1) * m, mk(0.0)), cs(m), sv(m + 1), y(m);
  std::vector<double> sn(m);
  int j = 1;
  double res = 0.0;
  bool done = false;
  while (j <= max_it) {
    // v0 = b - A x
    {
      std::vector<const K *> cx(L);
      for (size_t i = 0; i < L; ++i) cx[i] = x[i];
      KR(gmv_core(c, cx, w, 1));
    }
    for (size_t i = 0; i < L; ++i) {
      KR(k_scal_copy(c, c->subs[i]->n, -1.0, w[i], w[i]));
      KR(k_axpy(c, c->subs[i]->n, 1.0, b[i], w[i]));
    }
    KR(dots(c, 1, w, w, d_T, hv));
    if (j == 1 && hb_real(hv[0]) < 4.930380657631324e-32) {  // eps^2 (GMRES.hpp:75)
      j = 0;
      break;
    }
    sv.assign(m + 1, mk(0.0));
    const double beta0 = std::sqrt(hb_real(hv[0]));
    sv[0] = mk(beta0);
    for (size_t i = 0; i < L; ++i) KR(k_scal_copy(c, c->subs[i]->n, 1.0 / beta0, w[i], V[i]));
    std::fill(H.begin(), H.end(), mk(0.0));
    int i = 0;
    done = false;
    while (i < m && j <= max_it) {
      std::vector<const K *> vi(L);
      for (size_t q = 0; q < L; ++q) vi[q] = V[q] + (size_t)i * c->subs[q]->n;
      KR(apply_core(c, vi, z, 1, correction));  // z = M^-1 v_i   (GMRES.hpp:116)
      KR(gmv_core(c, cz, w, 1));                // w = A z        (GMRES.hpp:117)
      KR(dots(c, i + 1, V, w, d_T, hv));        // classical Gram-Schmidt: all products first
      KRC(cudaMemcpyAsync(d_h, hv.data(), (i + 1) * sizeof(K), cudaMemcpyHostToDevice, c->stream));
      for (size_t q = 0; q < L; ++q) KR(k_vupdate(c, c->subs[q], i + 1, V[q], d_h, -1.0, w[q]));
      std::vector<K> hcol(hv);
      KR(dots(c, 1, w, w, d_T, hv));
      const double hn = std::sqrt(hb_real(hv[0]));
      for (int k = 0; k <= i; ++k) H[k + (size_t)i * (m + 1)] = hcol[k];
      H[i + 1 + (size_t)i * (m + 1)] = mk(hn);
      if (i < m - 1)
        for (size_t q = 0; q < L; ++q) KR(k_scal_copy(c, c->subs[q]->n, hn == 0.0 ? 1.0 : 1.0 / hn, w[q], V[q] + (size_t)(i + 1) * c->subs[q]->n));
      K *Hc = &H[(size_t)i * (m + 1)];
      for (int k = 0; k < i; ++k) {  // previous rotations (iterative.hpp:690-697)
        const K g = hb_conj(cs[k]) * Hc[k] + sn[k] * Hc[k + 1];
        Hc[k + 1] = cs[k] * Hc[k + 1] - sn[k] * Hc[k];
        Hc[k] = g;
      }
      const double delta = std::hypot(hb_abs(Hc[i]), hb_abs(Hc[i + 1]));  // nrm2 of the two entries (iterative.hpp:701)
      sn[i] = hb_real(Hc[i + 1]) / delta;
      cs[i] = Hc[i] / delta;
      Hc[i] = mk(delta);
      Hc[i + 1] = mk(0.0);
      sv[i + 1] = -sn[i] * sv[i];
      sv[i] = sv[i] * hb_conj(cs[i]);
      ++i;
      res = hb_abs(sv[i]);
      if (res / normb <= tol) {  // checkConvergence (iterative.hpp:98-103)
        done = true;
        break;
      }
      ++j;
    }
    // updateSol (iterative.hpp:272-336): y = H^-1 s, x += M^-1 (V y)
    const int dim = i;
    for (int k = dim - 1; k >= 0; --k) {
      K acc = sv[k];
      for (int l = k + 1; l < dim; ++l) acc -= H[k + (size_t)l * (m + 1)] * y[l];
      y[k] = acc / H[k + (size_t)k * (m + 1)];
    }
    if (dim > 0) {
      KRC(cudaMemcpyAsync(d_h, y.data(), dim * sizeof(K), cudaMemcpyHostToDevice, c->stream));
      for (size_t q = 0; q < L; ++q) {
        KRC(cudaMemsetAsync(t[q], 0, (size_t)c->subs[q]->n * sizeof(K), c->stream));
        KR(k_vupdate(c, c->subs[q], dim, V[q], d_h, 1.0, t[q]));
      }
      std::vector<const K *> ct(L);
      for (size_t q = 0; q < L; ++q) ct[q] = t[q];
      KR(apply_core(c, ct, z, 1, correction));
      for (size_t q = 0; q < L; ++q) KR(k_axpy(c, c->subs[q]->n, 1.0, z[q], x[q]));
    }
    if (done || j > max_it) break;
  }
  KRC(cudaStreamSynchronize(c->stream));
  cleanup();
  *iterations = std::min(j, max_it);
  if (rel_residual) *rel_residual = res / normb;
  return 0;
#undef KR
#undef KRC
}

}  // namespace hb

using namespace hb;

extern "C" int HB_API(solve)(hb_ctx_t *ctx, const K *const *b, K *const *x, int mu, int correction, int restart, int max_it, double tol,
                                int where, int *iterations, double *rel_residual) {
  Ctx *c = reinterpret_cast<Ctx *>(ctx);
  if (!iterations || restart < 1 || max_it < 1) {
    set_error("solve: bad arguments");
    return HPDDM_B200_ERR_ARG;
  }
  HB_CHECK(check_ready(c, std::max(mu, 1)));
  const size_t L = c->subs.size();
  // vectors live in private device buffers for the whole solve (d_in / d_out are used by nothing else here)
  std::vector<K *> bd(L), xd(L);
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
  int itmax = 0, rc = 0;
  for (int col = 0; col < mu && rc == 0; ++col) {  // every column runs its own Krylov space (pseudo-block, like the reference's non-block GMRES)
    std::vector<const K *> bc(L);
    std::vector<K *> xc(L);
    for (size_t i = 0; i < L; ++i) {
      bc[i] = bd[i] + (size_t)col * c->subs[i]->n;
      xc[i] = xd[i] + (size_t)col * c->subs[i]->n;
    }
    int it = 0;
    double rr = 0.0;
    rc = gmres_device(c, bc, xc, correction, restart, max_it, tol, &it, &rr);
    itmax = std::max(itmax, it);
    if (rel_residual) rel_residual[col] = rr;
  }
  if (where == HPDDM_B200_HOST) {
    if (rc == 0)
      for (size_t i = 0; i < L; ++i) cudaMemcpyAsync(x[i], xd[i], (size_t)c->subs[i]->n * mu * sizeof(K), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    for (size_t i = 0; i < L; ++i) {
      cudaFree(bd[i]);
      cudaFree(xd[i]);
    }
  }
  *iterations = itmax;
  return rc;
}
