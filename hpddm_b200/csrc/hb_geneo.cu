// GenEO deflation vectors on the GPU ("next" row f-4 of SURVEY.md section 8f).
//
// Replaces Schwarz::solveGEVP<EIGENSOLVER> (include/HPDDM_schwarz.hpp:665-715): smallest eigenpairs of
//     A_Neu x = lambda B x ,   B = restriction to the overlap of D A_Neu D   (scaleIntoOverlap, schwarz.hpp:622-657)
// which the reference hands to ARPACK in shift-invert mode (sigma = 0, which = "LM", tol 1e-6:
// include/HPDDM_ARPACK.hpp:84-178, include/HPDDM_eigensolver.hpp:69).  Here: block subspace iteration with
// Rayleigh-Ritz on  theta = 1 / lambda  of  A_Neu^-1 B, every heavy step being a hot-path kernel:
//     Y = A_Neu^-1 W      block SpTRSV (4 right-hand sides per pass over the Neumann factor)
//     W' = B Y            CSR SpMM
//     G_A = Y^T W, G_B = Y^T W'        tall-skinny products (kk_zt)
//     small k x k Rayleigh-Ritz on the host (Cholesky + Jacobi), W <- W' C, finally Z = Y C[:, :nu].
// (A Y = W by construction, and B (Y C) = W' C: no product is ever recomputed.)
#include <algorithm>
#include <cmath>
#include <random>
#include <vector>

#include "hb_internal.h"

#ifdef HB_COMPLEX
// Complex scalars (BASELINE.json config 5: Helmholtz + DtN coarse space): the reference has no in-tree pencil for that case
// either -- the driver supplies its coarse vectors through set_vectors.
using namespace hb;
extern "C" int HB_API(sub_solve_gevp)(hb_sub_t *, int, int, const int *, const int *, const K *, int, char, int, double, int, double *) {
  set_error("solve_gevp: the GPU GenEO eigensolver is implemented for real scalars only; pass the coarse vectors with " HB_PREFIX "_sub_set_vectors");
  return HPDDM_B200_ERR_STATE;
}
#else
namespace hb {

namespace {

// cyclic Jacobi for a symmetric k x k matrix (column-major); eigenvalues in w, eigenvectors in V
void jacobi_eig(int k, std::vector<double> &M, std::vector<double> &w, std::vector<double> &V) {
  V.assign((size_t)k * k, 0.0);
  for (int i = 0; i < k; ++i) V[i + (size_t)i * k] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int j = 0; j < k; ++j)
      for (int i = 0; i < k; ++i) (i == j ? diag : off) += M[i + (size_t)j * k] * M[i + (size_t)j * k];
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < k - 1; ++p)
      for (int q = p + 1; q < k; ++q) {
        const double apq = M[p + (size_t)q * k];
        if (std::fabs(apq) < 1e-300) continue;
        const double app = M[p + (size_t)p * k], aqq = M[q + (size_t)q * k];
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = t * cs;
        for (int i = 0; i < k; ++i) {  // columns p, q
          const double mp = M[i + (size_t)p * k], mq = M[i + (size_t)q * k];
          M[i + (size_t)p * k] = cs * mp - sn * mq;
          M[i + (size_t)q * k] = sn * mp + cs * mq;
        }
        for (int i = 0; i < k; ++i) {  // rows p, q
          const double mp = M[p + (size_t)i * k], mq = M[q + (size_t)i * k];
          M[p + (size_t)i * k] = cs * mp - sn * mq;
          M[q + (size_t)i * k] = sn * mp + cs * mq;
        }
        for (int i = 0; i < k; ++i) {
          const double vp = V[i + (size_t)p * k], vq = V[i + (size_t)q * k];
          V[i + (size_t)p * k] = cs * vp - sn * vq;
          V[i + (size_t)q * k] = sn * vp + cs * vq;
        }
      }
  }
  w.resize(k);
  for (int i = 0; i < k; ++i) w[i] = M[i + (size_t)i * k];
}

// Rayleigh-Ritz for the pencil (GB, GA), GA SPD: returns theta (descending) and C (k x k) with C^T GA C = I
int rayleigh_ritz(int k, std::vector<double> GA, std::vector<double> GB, std::vector<double> &theta, std::vector<double> &C) {
  // GA = R^T R (upper R, column-major)
  std::vector<double> R((size_t)k * k, 0.0);
  for (int j = 0; j < k; ++j) {
    for (int i = 0; i <= j; ++i) {
      double s = GA[i + (size_t)j * k];
      for (int l = 0; l < i; ++l) s -= R[l + (size_t)i * k] * R[l + (size_t)j * k];
      if (i == j) {
        if (!(s > 0.0)) return -1;
        R[j + (size_t)j * k] = std::sqrt(s);
      } else
        R[i + (size_t)j * k] = s / R[i + (size_t)i * k];
    }
  }
  // Rinv (upper)
  std::vector<double> Ri((size_t)k * k, 0.0);
  for (int j = 0; j < k; ++j) {
    Ri[j + (size_t)j * k] = 1.0 / R[j + (size_t)j * k];
    for (int i = j - 1; i >= 0; --i) {
      double s = 0.0;
      for (int l = i + 1; l <= j; ++l) s += R[i + (size_t)l * k] * Ri[l + (size_t)j * k];
      Ri[i + (size_t)j * k] = -s / R[i + (size_t)i * k];
    }
  }
  // M = Ri^T GB Ri
  std::vector<double> T((size_t)k * k, 0.0), M((size_t)k * k, 0.0);
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int l = 0; l <= j; ++l) s += GB[i + (size_t)l * k] * Ri[l + (size_t)j * k];
      T[i + (size_t)j * k] = s;
    }
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int l = 0; l <= i; ++l) s += Ri[l + (size_t)i * k] * T[l + (size_t)j * k];
      M[i + (size_t)j * k] = s;
    }
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < j; ++i) M[i + (size_t)j * k] = M[j + (size_t)i * k] = 0.5 * (M[i + (size_t)j * k] + M[j + (size_t)i * k]);
  std::vector<double> w, V;
  jacobi_eig(k, M, w, V);
  std::vector<int> ord(k);
  for (int i = 0; i < k; ++i) ord[i] = i;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) { return w[a] > w[b]; });
  theta.resize(k);
  C.assign((size_t)k * k, 0.0);
  for (int j = 0; j < k; ++j) {
    theta[j] = w[ord[j]];
    for (int i = 0; i < k; ++i) {
      double s = 0.0;
      for (int l = i; l < k; ++l) s += Ri[i + (size_t)l * k] * V[l + (size_t)ord[j] * k];
      C[i + (size_t)j * k] = s;
    }
  }
  return 0;
}

}  // namespace
}  // namespace hb

using namespace hb;

extern "C" int HB_API(sub_solve_gevp)(hb_sub_t *sub, int n, int nnz, const int *ia, const int *ja, const double *a, int sym, char numbering, int nu,
                                         double tol, int max_it, double *eigenvalues) {
  Sub *s = reinterpret_cast<Sub *>(sub);
  if (!s || nu < 1 || n != s->n) {
    set_error("solve_gevp: bad arguments (n = %d, subdomain order %d, nu = %d)", n, s ? s->n : -1, nu);
    return HPDDM_B200_ERR_ARG;
  }
  if ((int)s->d_host.size() != n) {
    set_error("solve_gevp: set the partition of unity first (hpddm_b200_sub_set_scaling)");
    return HPDDM_B200_ERR_STATE;
  }
  Ctx *c = s->ctx;
  HB_CUDA(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  if (tol <= 0.0) tol = 1e-6;     // eigensolver_tol default (include/HPDDM_eigensolver.hpp:69)
  if (max_it <= 0) max_it = 100;
  HostCSR AN;
  HB_CHECK(to_host_csr(n, nnz, ia, ja, a, sym, numbering, AN));
  // ---- B = D A_Neu D on the overlap dofs with d > EPS (schwarz.hpp:626-641)
  std::vector<char> into(n, 0);
  for (int v : s->nb_idx)
    if (s->d_host[v] > 1e-12) into[v] = 1;
  std::vector<int> bia(n + 1, 0), bja;
  std::vector<double> ba;
  for (int i = 0; i < n; ++i) {
    if (into[i])
      for (int k = AN.ia[i]; k < AN.ia[i + 1]; ++k) {
        const int j = AN.ja[k];
        const double v = s->d_host[i] * s->d_host[j] * AN.a[k];
        if (into[j] && std::fabs(v) > 1e-12) {
          bja.push_back(j);
          ba.push_back(v);
        }
      }
    bia[i + 1] = (int)bja.size();
  }
  const int k = std::min(n, std::max(2 * nu, nu + 8));
  // ---- factor A_Neu (a private solver object; retried with a tiny shift when A_Neu is singular: floating subdomain)
  Sub ns;
  ns.ctx = c;
  ns.grank = s->grank;
  ns.n = n;
  ns.gx = s->gx;
  ns.gy = s->gy;
  ns.gz = s->gz;
  ns.gdof = s->gdof;
  int rc = numfact_device(&ns, AN);
  if (rc == HPDDM_B200_ERR_NUMERIC) {
    double dmax = 0.0;
    for (int i = 0; i < n; ++i)
      for (int q = AN.ia[i]; q < AN.ia[i + 1]; ++q)
        if (AN.ja[q] == i) dmax = std::max(dmax, std::fabs(AN.a[q]));
    for (int i = 0; i < n; ++i)
      for (int q = AN.ia[i]; q < AN.ia[i + 1]; ++q)
        if (AN.ja[q] == i) AN.a[q] += 1e-10 * dmax;
    rc = numfact_device(&ns, AN);
  }
  if (rc < 0) return rc;
  int *d_bia = nullptr, *d_bja = nullptr;
  double *d_ba = nullptr, *Y = nullptr, *W = nullptr, *W2 = nullptr, *ones = nullptr, *d_G = nullptr, *d_C = nullptr;
  auto cleanup = [&]() {
    free_factor(ns.fac);
    for (void *p : {(void *)d_bia, (void *)d_bja, (void *)d_ba, (void *)Y, (void *)W, (void *)W2, (void *)ones, (void *)d_G, (void *)d_C}) cudaFree(p);
  };
#define GV(call)      \
  do {                \
    int r__ = (call); \
    if (r__ < 0) {    \
      cleanup();      \
      return r__;     \
    }                 \
  } while (0)
#define GVC(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      cleanup();                                                                                   \
      return HPDDM_B200_ERR_CUDA;                                                                  \
    }                                                                                              \
  } while (0)
  const size_t nk = (size_t)n * k;
  GVC(cudaMalloc(&d_bia, (n + 1) * sizeof(int)));
  GVC(cudaMalloc(&d_bja, std::max<size_t>(bja.size(), 1) * sizeof(int)));
  GVC(cudaMalloc(&d_ba, std::max<size_t>(ba.size(), 1) * sizeof(double)));
  GVC(cudaMalloc(&Y, nk * sizeof(double)));
  GVC(cudaMalloc(&W, nk * sizeof(double)));
  GVC(cudaMalloc(&W2, nk * sizeof(double)));
  GVC(cudaMalloc(&ones, (size_t)n * sizeof(double)));
  GVC(cudaMalloc(&d_G, (size_t)2 * k * k * sizeof(double)));
  GVC(cudaMalloc(&d_C, (size_t)k * k * sizeof(double)));
  GVC(cudaMemcpyAsync(d_bia, bia.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (!bja.empty()) {
    GVC(cudaMemcpyAsync(d_bja, bja.data(), bja.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    GVC(cudaMemcpyAsync(d_ba, ba.data(), ba.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  GV(k_fill(c, n, 1.0, ones));
  {
    // deterministic start block (ARPACK starts from a random residual; include/HPDDM_ARPACK.hpp:101)
    std::vector<double> X0(nk);
    std::mt19937 gen(1234 + s->grank);
    std::uniform_real_distribution<double> dis(-1.0, 1.0);
    for (double &v : X0) v = dis(gen);
    GVC(cudaMemcpyAsync(Y, X0.data(), nk * sizeof(double), cudaMemcpyHostToDevice, st));
    GVC(cudaStreamSynchronize(st));
  }
  GV(k_spmv_raw(c, n, (int64_t)bja.size(), d_bia, d_bja, d_ba, k, 1.0, Y, 0.0, nullptr, W, nullptr));  // W = B X0
  std::vector<double> GA((size_t)k * k), GB((size_t)k * k), theta, prev, C;
  bool converged = false;
  int it = 0;
  for (it = 1; it <= max_it; ++it) {
    GV(solve_cols(&ns, W, Y, k, nullptr, false));                                                          // Y = A_Neu^-1 W
    GV(k_spmv_raw(c, n, (int64_t)bja.size(), d_bia, d_bja, d_ba, k, 1.0, Y, 0.0, nullptr, W2, nullptr));   // W' = B Y
    GVC(cudaMemsetAsync(d_G, 0, (size_t)2 * k * k * sizeof(double), st));
    GV(k_zt_raw(c, n, k, Y, ones, k, W, d_G, k));                                                           // G_A = Y^T W  (= Y^T A Y)
    GV(k_zt_raw(c, n, k, Y, ones, k, W2, d_G + (size_t)k * k, k));                                          // G_B = Y^T B Y
    GVC(cudaMemcpyAsync(GA.data(), d_G, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToHost, st));
    GVC(cudaMemcpyAsync(GB.data(), d_G + (size_t)k * k, (size_t)k * k * sizeof(double), cudaMemcpyDeviceToHost, st));
    GVC(cudaStreamSynchronize(st));
    for (int j = 0; j < k; ++j)
      for (int i = 0; i < j; ++i) {
        GA[i + (size_t)j * k] = GA[j + (size_t)i * k] = 0.5 * (GA[i + (size_t)j * k] + GA[j + (size_t)i * k]);
        GB[i + (size_t)j * k] = GB[j + (size_t)i * k] = 0.5 * (GB[i + (size_t)j * k] + GB[j + (size_t)i * k]);
      }
    if (rayleigh_ritz(k, GA, GB, theta, C) < 0) {
      set_error("solve_gevp: Rayleigh-Ritz breakdown at iteration %d (subspace lost rank; fewer than %d overlap dofs?)", it, k);
      cleanup();
      return HPDDM_B200_ERR_NUMERIC;
    }
    if (!prev.empty()) {
      double worst = 0.0;
      for (int j = 0; j < nu; ++j) worst = std::max(worst, std::fabs(theta[j] - prev[j]) / std::max(std::fabs(theta[j]), 1e-300));
      converged = worst < tol * 1e-2;  // eigenvalue change ~ (eigenvector error)^2: 1e-2 * tol keeps the vectors near ARPACK's accuracy
    }
    prev = theta;
    GVC(cudaMemcpyAsync(d_C, C.data(), (size_t)k * k * sizeof(double), cudaMemcpyHostToDevice, st));
    if (converged || it == max_it) break;
    GV(k_zexp_raw(c, n, k, W2, ones, k, d_C, k, W));  // W = W' C  ( = B (Y C) )
  }
  // Z = Y C[:, :nu]
  if (s->d_Z) cudaFree(s->d_Z);
  s->d_Z = nullptr;
  GVC(cudaMalloc(&s->d_Z, (size_t)n * nu * sizeof(double)));
  GV(k_zexp_raw(c, n, k, Y, ones, nu, d_C, k, s->d_Z));
  GV(k_flush_tiny(c, (int64_t)n * nu, 1.0e-18, s->d_Z));  // schwarz.hpp:713
  s->nu = nu;
  s->ctx->epoch++;
  if (eigenvalues)
    for (int j = 0; j < nu; ++j) eigenvalues[j] = 1.0 / theta[j];
  GVC(cudaStreamSynchronize(st));
  cleanup();
  if (!converged && getenv("HPDDM_B200_DEBUG")) fprintf(stderr, "[hpddm_b200] solve_gevp: stopped after %d iterations without meeting tol %g\n", it, tol);
  return it;
#undef GV
#undef GVC
}

#endif  // HB_COMPLEX

extern "C" int HB_API(sub_get_vectors)(hb_sub_t *sub, K *Z, int *nu) {
  Sub *s = reinterpret_cast<Sub *>(sub);
  if (!s || !nu) return HPDDM_B200_ERR_ARG;
  *nu = s->nu;
  if (Z && s->nu > 0) {
    HB_CUDA(cudaSetDevice(s->ctx->device));
    HB_CUDA(cudaMemcpyAsync(Z, s->d_Z, (size_t)s->n * s->nu * sizeof(K), cudaMemcpyDeviceToHost, s->ctx->stream));
    HB_CUDA(cudaStreamSynchronize(s->ctx->stream));
  }
  return 0;
}
