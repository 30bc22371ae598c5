// GCRO-DR (IterativeMethod::GCRODR, include/HPDDM_GCRODR.hpp:35-444) with the reference defaults (iterative.hpp:197-218): right
// preconditioning, classical Gram-Schmidt, CholQR, harmonic Ritz values selected by `recycle_target`, `recycle_strategy` A or B,
// recycle_same_system = 0.  Every right-hand side has its own Krylov space, Hessenberg matrix and recycled pair (U, C), the
// preconditioner and the operator are applied to all columns at once -- the non-block driver of the reference.  Host code only:
// vectors are touched through gcro::Backend (hb_gcrodr.h), the rest is dense algebra of order restart + 1 in complex arithmetic
// (real scalars: zero imaginary parts; products of reals stay exact).
//
//   first cycle of a first solve  = GMRES(m); its k harmonic Ritz vectors give the first pair      (GCRODR.hpp:179-214, 242-316)
//   later cycles                   = Arnoldi on (I - C C^H D) A M^-1 with m - k new vectors,
//                                    x += M^-1 (U (C^H D r - B y) + V y)                             (GCRODR.hpp:187-196, iterative.hpp:338-393)
//                                    new pair from  G^H G z = theta G^H W^H D [U~ V] z               (GCRODR.hpp:317-430)
//   a later solve                  = C = A M^-1 U re-orthonormalised by CholQR, residual projected  (GCRODR.hpp:94-130)
//
// Only the span of the selected eigenvectors enters the iteration (another basis of the same span changes U and C by one and
// the same unitary diagonal factor), so the eigen-solver below (complex Schur form + back-substitution) stands in for the
// reference's hseqr + hsein / ggev; the generalised problem G^H G z = theta G^H W^H D [U~ V] z is solved through the thin QR of G as the
// standard problem R^-1 (Q^H W^H D [U~ V]) z = z / theta (harmonic_pencil below).
#include "hb_gcrodr.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <limits>

namespace hb {
namespace gcro {

namespace {

typedef std::complex<double> zc;
inline zc to_z(K a) { return zc(hb_real(a), hb_imag(a)); }
inline K from_z(zc a) { return mk(a.real(), a.imag()); }

struct M {  // dense column-major matrix
  int r = 0, c = 0;
  std::vector<zc> a;
  M() {}
  M(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, zc(0.0)) {}
  zc &operator()(int i, int j) { return a[i + (size_t)j * r]; }
  const zc &operator()(int i, int j) const { return a[i + (size_t)j * r]; }
};

M mul(const M &A, const M &B, bool conjA = false) {  // A B or A^H B
  const int rows = conjA ? A.c : A.r, inner = conjA ? A.r : A.c;
  M C(rows, B.c);
  for (int j = 0; j < B.c; ++j)
    for (int l = 0; l < inner; ++l) {
      const zc b = B(l, j);
      if (b == zc(0.0)) continue;
      for (int i = 0; i < rows; ++i) C(i, j) += (conjA ? std::conj(A(l, i)) : A(i, l)) * b;
    }
  return C;
}

bool eig(const M &A0, std::vector<zc> &w, M &X) {
  const int n = A0.r;
  const double eps = std::numeric_limits<double>::epsilon();
  M H = A0, Z(n, n);
  w.assign(n, zc(0.0));
  for (int i = 0; i < n; ++i) Z(i, i) = 1.0;
  std::vector<zc> v(n);
  for (int k = 0; k + 2 < n; ++k) {  // Householder reduction to Hessenberg form, H <- P H P, Z <- Z P
    double alpha = 0.0;
    for (int i = k + 1; i < n; ++i) alpha += std::norm(H(i, k));
    alpha = std::sqrt(alpha);
    if (alpha == 0.0) continue;
    const zc x0 = H(k + 1, k);
    const zc phase = std::abs(x0) > 0.0 ? x0 / std::abs(x0) : zc(1.0);
    double vn = 0.0;
    for (int i = k + 1; i < n; ++i) {
      v[i] = H(i, k);
      if (i == k + 1) v[i] += phase * alpha;
      vn += std::norm(v[i]);
    }
    if (vn == 0.0) continue;
    const double beta = 2.0 / vn;
    for (int j = 0; j < n; ++j) {
      zc s(0.0);
      for (int i = k + 1; i < n; ++i) s += std::conj(v[i]) * H(i, j);
      s *= beta;
      for (int i = k + 1; i < n; ++i) H(i, j) -= v[i] * s;
    }
    for (int i = 0; i < n; ++i) {
      zc s(0.0), t(0.0);
      for (int j = k + 1; j < n; ++j) {
        s += H(i, j) * v[j];
        t += Z(i, j) * v[j];
      }
      s *= beta;
      t *= beta;
      for (int j = k + 1; j < n; ++j) {
        H(i, j) -= s * std::conj(v[j]);
        Z(i, j) -= t * std::conj(v[j]);
      }
    }
  }
  double hnorm = 0.0;
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      if (i > j + 1) H(i, j) = 0.0;
      hnorm = std::max(hnorm, std::abs(H(i, j)));
    }
  if (hnorm == 0.0) hnorm = 1.0;
  std::vector<zc> cs(n), sn(n);
  int hi = n - 1, iter = 0;
  while (hi >= 0) {  // explicitly shifted QR sweeps on the active block l .. hi, deflation from the bottom
    int l = hi;
    while (l > 0) {
      double sd = std::abs(H(l - 1, l - 1)) + std::abs(H(l, l));
      if (sd == 0.0) sd = hnorm;
      if (std::abs(H(l, l - 1)) <= eps * sd) {
        H(l, l - 1) = 0.0;
        break;
      }
      --l;
    }
    if (l == hi) {
      w[hi] = H(hi, hi);
      --hi;
      iter = 0;
      continue;
    }
    if (++iter > 300) return false;
    const zc a = H(hi - 1, hi - 1), b = H(hi - 1, hi), c = H(hi, hi - 1), d = H(hi, hi);
    const zc tr = 0.5 * (a + d), disc = std::sqrt(tr * tr - (a * d - b * c));
    zc sigma = std::abs(tr + disc - d) < std::abs(tr - disc - d) ? tr + disc : tr - disc;  // Wilkinson shift
    if (iter % 10 == 0) sigma = d + zc(0.75 * std::abs(c), 0.0);                            // exceptional shift
    for (int i = l; i <= hi; ++i) H(i, i) -= sigma;
    for (int k = l; k < hi; ++k) {  // (H - sigma) = Q R
      const zc f = H(k, k), g = H(k + 1, k);
      const double r = std::sqrt(std::norm(f) + std::norm(g));
      cs[k] = r == 0.0 ? zc(1.0) : f / r;
      sn[k] = r == 0.0 ? zc(0.0) : g / r;
      for (int j = k; j < n; ++j) {
        const zc t1 = std::conj(cs[k]) * H(k, j) + std::conj(sn[k]) * H(k + 1, j);
        H(k + 1, j) = -sn[k] * H(k, j) + cs[k] * H(k + 1, j);
        H(k, j) = t1;
      }
      H(k + 1, k) = 0.0;
    }
    for (int k = l; k < hi; ++k) {  // R Q + sigma, Z <- Z Q
      for (int i = 0; i <= std::min(k + 1, hi); ++i) {
        const zc t1 = H(i, k) * cs[k] + H(i, k + 1) * sn[k];
        H(i, k + 1) = -H(i, k) * std::conj(sn[k]) + H(i, k + 1) * std::conj(cs[k]);
        H(i, k) = t1;
      }
      for (int i = 0; i < n; ++i) {
        const zc t1 = Z(i, k) * cs[k] + Z(i, k + 1) * sn[k];
        Z(i, k + 1) = -Z(i, k) * std::conj(sn[k]) + Z(i, k + 1) * std::conj(cs[k]);
        Z(i, k) = t1;
      }
    }
    for (int i = l; i <= hi; ++i) H(i, i) += sigma;
  }
  X = M(n, n);  // eigenvectors of the triangular factor, then X = Z Y
  const double small = eps * hnorm;
  std::vector<zc> y(n);
  for (int j = 0; j < n; ++j) {
    y[j] = 1.0;
    for (int i = j - 1; i >= 0; --i) {
      zc s(0.0);
      for (int q = i + 1; q <= j; ++q) s += H(i, q) * y[q];
      zc den = H(i, i) - H(j, j);
      if (std::abs(den) < small) den = small;
      y[i] = -s / den;
    }
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) {
      zc s(0.0);
      for (int q = 0; q <= j; ++q) s += Z(i, q) * y[q];
      X(i, j) = s;
      nrm += std::norm(s);
    }
    nrm = std::sqrt(nrm);
    if (nrm > 0.0)
      for (int i = 0; i < n; ++i) X(i, j) /= nrm;
  }
  return true;
}

// P = Q R by Householder reflections: Q (rows x k, orthonormal columns), R (k x k upper triangular)
void qr(const M &P, M &Q, M &R) {
  const int rows = P.r, k = P.c;
  M A = P;
  std::vector<std::vector<zc>> vs(k);
  std::vector<double> betas(k, 0.0);
  for (int j = 0; j < k && j < rows; ++j) {
    double alpha = 0.0;
    for (int i = j; i < rows; ++i) alpha += std::norm(A(i, j));
    alpha = std::sqrt(alpha);
    std::vector<zc> &v = vs[j];
    v.assign(rows, zc(0.0));
    if (alpha == 0.0) continue;
    const zc x0 = A(j, j);
    const zc phase = std::abs(x0) > 0.0 ? x0 / std::abs(x0) : zc(1.0);
    double vn = 0.0;
    for (int i = j; i < rows; ++i) {
      v[i] = A(i, j);
      if (i == j) v[i] += phase * alpha;
      vn += std::norm(v[i]);
    }
    if (vn == 0.0) continue;
    betas[j] = 2.0 / vn;
    for (int c = j; c < k; ++c) {
      zc s(0.0);
      for (int i = j; i < rows; ++i) s += std::conj(v[i]) * A(i, c);
      s *= betas[j];
      for (int i = j; i < rows; ++i) A(i, c) -= v[i] * s;
    }
  }
  R = M(k, k);
  for (int c = 0; c < k; ++c)
    for (int i = 0; i <= c && i < rows; ++i) R(i, c) = A(i, c);
  Q = M(rows, k);
  for (int c = 0; c < k && c < rows; ++c) Q(c, c) = 1.0;
  for (int j = std::min(k, rows) - 1; j >= 0; --j) {
    if (betas[j] == 0.0) continue;
    const std::vector<zc> &v = vs[j];
    for (int c = 0; c < k; ++c) {
      zc s(0.0);
      for (int i = j; i < rows; ++i) s += std::conj(v[i]) * Q(i, c);
      s *= betas[j];
      for (int i = j; i < rows; ++i) Q(i, c) -= v[i] * s;
    }
  }
}

// Y R = V for Y (R upper triangular k x k): the trsm("R", "U", "N", "N") of GCRODR.hpp:312,420
M solve_right_upper(const M &V, const M &R) {
  M Y(V.r, V.c);
  for (int c = 0; c < V.c; ++c)
    for (int i = 0; i < V.r; ++i) {
      zc s = V(i, c);
      for (int l = 0; l < c; ++l) s -= Y(i, l) * R(l, c);
      Y(i, c) = s / R(c, c);
    }
  return Y;
}

// G = R^H R (potrf "U"); false if G is not positive definite
bool chol_upper(const M &G, M &R) {
  const int n = G.r;
  R = M(n, n);
  for (int j = 0; j < n; ++j) {
    double d = G(j, j).real();
    for (int l = 0; l < j; ++l) d -= std::norm(R(l, j));
    if (!(d > 0.0)) return false;
    d = std::sqrt(d);
    R(j, j) = d;
    for (int i = j + 1; i < n; ++i) {
      zc s = G(j, i);
      for (int l = 0; l < j; ++l) s -= std::conj(R(l, j)) * R(l, i);
      R(j, i) = s / d;
    }
  }
  return true;
}

// Generalised harmonic Ritz problem of a cycle with a recycled pair (GCRODR.hpp:317-430, 761-884):  G^H G z = theta G^H What z  with
// What = W^H D [U~ V] (strategy A) or its idealisation [I 0; 0 I; 0 0] (strategy B).  The reference hands the pencil (G^H G, G^H What) to
// ggev; here G = Q R (thin) turns it into the standard problem  R^-1 (Q^H What) z = z / theta  -- no squared condition number, which
// keeps the selected subspace the reference's one also for large recycled dimensions (checked: GCRO-DR(100, 50) on the 40X sequence).
bool harmonic_pencil(const M &G, const M &What, std::vector<zc> &th, M &X) {
  const int dim = G.c;
  M Q, R;
  qr(G, Q, R);
  M T = mul(Q, What, true);  // dim x dim
  for (int c = 0; c < dim; ++c)
    for (int r = dim - 1; r >= 0; --r) {
      if (R(r, r) == zc(0.0)) return false;
      zc acc = T(r, c);
      for (int l = r + 1; l < dim; ++l) acc -= R(r, l) * T(l, c);
      T(r, c) = acc / R(r, r);
    }
  std::vector<zc> muv;
  if (!eig(T, muv, X)) return false;
  th.resize(dim);
  for (int q = 0; q < dim; ++q) th[q] = std::abs(muv[q]) > 0.0 ? zc(1.0) / muv[q] : zc(std::numeric_limits<double>::infinity(), 0.0);
  return true;
}

// selectNu (include/HPDDM_specifications.hpp:90-123): indices ordered by the recycle target
std::vector<int> order(const std::vector<zc> &th, int target) {
  std::vector<double> key(th.size());
  for (size_t i = 0; i < th.size(); ++i) {
    const zc z = th[i];
    switch (target) {
      case 1: key[i] = -std::norm(z); break;  // LM
      case 2: key[i] = z.real(); break;       // SR
      case 3: key[i] = -z.real(); break;      // LR
      case 4: key[i] = z.imag(); break;       // SI
      case 5: key[i] = -z.imag(); break;      // LI
      default: key[i] = std::norm(z);         // SM
    }
    if (key[i] != key[i]) key[i] = std::numeric_limits<double>::infinity();
  }
  std::vector<int> idx(th.size());
  for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int)i;
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return key[a] < key[b]; });
  return idx;
}

// The k columns handed to the products of GCRODR.hpp:307-308 / 403-419.  Complex scalars: the eigenvectors of the first k eigenvalues
// of `ord`.  Real scalars: LAPACK's real storage of the same invariant subspace -- a real eigenvalue contributes its (real)
// eigenvector, a conjugate pair the real and the imaginary part of one of its eigenvectors; a pair cut by the k-th position
// contributes the real part only (the reference keeps one column of it as well, GCRODR.hpp:283-296).
M select_columns(const std::vector<zc> &th, const M &X, const std::vector<int> &ord, int k) {
  const int n = X.r;
  M V(n, k);
  if (IS_COMPLEX) {
    for (int c = 0; c < k; ++c)
      for (int i = 0; i < n; ++i) V(i, c) = X(i, ord[c]);
    return V;
  }
  std::vector<char> used(th.size(), 0);
  int cols = 0;
  for (size_t o = 0; o < ord.size() && cols < k; ++o) {
    const int j = ord[o];
    if (used[j]) continue;
    used[j] = 1;
    const zc z = th[j];
    const bool real_ev = !(std::abs(z.imag()) > 1e-10 * std::max(1.0, std::abs(z)));
    if (real_ev) {  // rotate the arbitrary phase of the complex arithmetic away
      int big = 0;
      for (int i = 1; i < n; ++i)
        if (std::abs(X(i, j)) > std::abs(X(big, j))) big = i;
      const zc ph = std::abs(X(big, j)) > 0.0 ? std::conj(X(big, j)) / std::abs(X(big, j)) : zc(1.0);
      for (int i = 0; i < n; ++i) V(i, cols) = (X(i, j) * ph).real();
      ++cols;
    } else {
      int partner = -1;
      double best = 0.0;
      for (size_t q = 0; q < th.size(); ++q) {
        if (used[q]) continue;
        const double dist = std::abs(th[q] - std::conj(z));
        if (partner < 0 || dist < best) {
          partner = (int)q;
          best = dist;
        }
      }
      if (partner >= 0 && best <= 1e-6 * std::max(1.0, std::abs(z))) used[partner] = 1;
      for (int i = 0; i < n; ++i) V(i, cols) = X(i, j).real();
      ++cols;
      if (cols < k) {
        for (int i = 0; i < n; ++i) V(i, cols) = X(i, j).imag();
        ++cols;
      }
    }
  }
  return V;
}

// LAPACK-convention Householder tools (zlarfg, zlarf, zgeqr2, zunm2r) on column-major zc data: the block driver keeps the reflectors of
// its 2 mu x mu reductions the way geqrf does, because the reference's convergence test reads single entries of Q^H applied to the
// block residual (checkBlockConvergence, iterative.hpp:139-146)
void larfg(int n, zc &alpha, zc *x, zc &tau) {
  double xn = 0.0;
  for (int i = 0; i < n - 1; ++i) xn += std::norm(x[i]);
  xn = std::sqrt(xn);
  if (n <= 0 || (xn == 0.0 && alpha.imag() == 0.0)) {
    tau = 0.0;
    return;
  }
  const double beta = -std::copysign(std::sqrt(std::norm(alpha) + xn * xn), alpha.real());
  tau = zc((beta - alpha.real()) / beta, -alpha.imag() / beta);
  const zc scal = zc(1.0) / (alpha - beta);
  for (int i = 0; i < n - 1; ++i) x[i] *= scal;
  alpha = beta;
}
void larf_left(int m, int nc, const zc *v1, zc t, zc *Cm, int ldc) {  // C <- (I - t v v^H) C, v = (1; v1)
  if (t == zc(0.0)) return;
  for (int j = 0; j < nc; ++j) {
    zc *cj = Cm + (size_t)j * ldc;
    zc w = cj[0];
    for (int r = 1; r < m; ++r) w += std::conj(v1[r - 1]) * cj[r];
    w *= t;
    cj[0] -= w;
    for (int r = 1; r < m; ++r) cj[r] -= v1[r - 1] * w;
  }
}
void geqr2(int m, int n, zc *A, int lda, zc *tau) {
  for (int i = 0; i < std::min(m, n); ++i) {
    larfg(m - i, A[i + (size_t)i * lda], A + i + 1 + (size_t)i * lda, tau[i]);
    if (i < n - 1) larf_left(m - i, n - i - 1, A + i + 1 + (size_t)i * lda, std::conj(tau[i]), A + i + (size_t)(i + 1) * lda, lda);
  }
}
void unm2r_left(bool conj_trans, int m, int nc, int k, const zc *A, int lda, const zc *tau, zc *Cm, int ldc) {  // C <- Q^H C or Q C
  if (conj_trans)
    for (int i = 0; i < k; ++i) larf_left(m - i, nc, A + i + 1 + (size_t)i * lda, std::conj(tau[i]), Cm + i, ldc);
  else
    for (int i = k - 1; i >= 0; --i) larf_left(m - i, nc, A + i + 1 + (size_t)i * lda, tau[i], Cm + i, ldc);
}

Vec block(const Backend &be, const Vec &base, int r, int mu) {
  Vec out(base.size());
  for (size_t q = 0; q < base.size(); ++q) out[q] = base[q] + (size_t)r * mu * be.rows(q);
  return out;
}

struct Column {  // per right-hand side state of one cycle
  M R;               // rotated Hessenberg matrix, absolute row / column indices (rows < shift unused)
  M save;            // unrotated Hessenberg matrix of this cycle, column i - shift
  M B;               // C^H D A M^-1 v_i, column i
  std::vector<zc> cs, s;
  std::vector<double> sn;
};

}  // namespace

bool eig_general(int n, const double *a, double *w, double *x) {
  M A(n, n), X;
  for (size_t i = 0; i < (size_t)n * n; ++i) A.a[i] = zc(a[2 * i], a[2 * i + 1]);
  std::vector<zc> ev;
  if (!eig(A, ev, X)) return false;
  for (int i = 0; i < n; ++i) {
    w[2 * i] = ev[i].real();
    w[2 * i + 1] = ev[i].imag();
  }
  for (size_t i = 0; i < (size_t)n * n; ++i) {
    x[2 * i] = X.a[i].real();
    x[2 * i + 1] = X.a[i].imag();
  }
  return true;
}

#define HB_GC_RET(call)      \
  do {                       \
    const int r__ = (call);  \
    if (r__ < 0) return r__; \
  } while (0)
#define GC(call)             \
  do {                       \
    const int r__ = (call);  \
    if (r__ < 0) {           \
      cleanup();             \
      return r__;            \
    }                        \
  } while (0)

int run(Backend &be, const Vec &b, const Vec &x, const Params &p, int *iterations, double *rel_residual) {
  const int mu = p.mu, max_it = p.max_it;
  const int m = std::min(p.restart, max_it);  // iterative.hpp:210
  int k = std::min(m - 1, p.recycle);         // iterative.hpp:215
  Recycled &rec = be.recycled();
  if (rec.k > 0 && (rec.mu != mu || rec.block)) {  // the reference re-interprets a pair stored for another number of right-hand sides
    be.release(rec.U);              // (GCRODR.hpp:66-67); here such a pair is dropped and rebuilt
    be.release(rec.C);
    rec.k = rec.mu = 0;
  }
  bool haveU = rec.k > 0;
  if (haveU) k = rec.k;
  Vec V, z, work, pt, scratch;
  auto cleanup = [&]() {
    be.release(V);
    be.release(z);
    be.release(work);
    be.release(pt);
    be.release(scratch);
  };
  GC(be.alloc(V, m + 1));
  GC(be.alloc(z, 1));
  GC(be.alloc(work, 1));
  std::vector<K> hv, coef;
  std::vector<double> norm(mu), res(mu, 0.0), resnorm(mu, 0.0);
  std::vector<int> conv(mu, -m);
  std::vector<Column> col(mu);
  auto v = [&](int r) { return block(be, V, r, mu); };
  auto tok = [&](const std::vector<zc> &src, int n0, int cnt) {  // zc -> K coefficients
    coef.resize(cnt);
    for (int i = 0; i < cnt; ++i) coef[i] = from_z(src[n0 + i]);
    return coef.data();
  };
  GC(be.start(b, x));
  GC(be.rhs_norms(b, norm));
  for (int nu = 0; nu < mu; ++nu)
    if (norm[nu] < 1e-12) norm[nu] = 1.0;  // HPDDM_EPS (GCRODR.hpp:142)
  auto converged = [&](double r, int nu) { return p.tol > 0.0 ? r / norm[nu] <= p.tol : r <= -p.tol; };  // iterative.hpp:98-103
  int j = 1;
  while (j <= max_it) {
    const int shift = haveU ? k : 0;
    int i = shift;
    Vec vi = v(i);
    GC(be.gmv(x, vi));  // v_i = b - A x
    for (int nu = 0; nu < mu; ++nu) {
      GC(be.scal_col(nu, -1.0, vi, vi));
      GC(be.axpy_col(nu, 1.0, b, vi));
    }
    if (j == 1 && haveU && p.same_system) {  // GCRODR.hpp:115-129 with id[4] / 4 != 0: C still is A M^-1 U; r <- (I - C C^H D) r, x += M^-1 (U C^H D r)
      GC(be.dots(k, rec.C, vi, hv));
      for (int nu = 0; nu < mu; ++nu) {
        GC(be.zero_col(nu, work));
        GC(be.combine_col(nu, k, rec.C, &hv[(size_t)nu * k], -1.0, vi));
        GC(be.combine_col(nu, k, rec.U, &hv[(size_t)nu * k], 1.0, work));
      }
      GC(be.apply(work, z));
      for (int nu = 0; nu < mu; ++nu) GC(be.axpy_col(nu, 1.0, z, x));
    } else if (j == 1 && haveU) {  // GCRODR.hpp:94-130: C = A M^-1 U, CholQR, r <- (I - C C^H D) r, x += M^-1 U C^H D r
      GC(be.alloc(pt, k));
      for (int c = 0; c < k; ++c) {
        Vec uc = block(be, rec.U, c, mu), pc = block(be, pt, c, mu), cc = block(be, rec.C, c, mu);
        GC(be.apply(uc, pc));
        GC(be.gmv(pc, cc));
      }
      std::vector<M> G(mu, M(k, k));
      for (int c = 0; c < k; ++c) {
        Vec cc = block(be, rec.C, c, mu);
        GC(be.dots(k, rec.C, cc, hv));
        for (int nu = 0; nu < mu; ++nu)
          for (int a = 0; a < k; ++a) G[nu](a, c) = to_z(hv[(size_t)nu * k + a]);
      }
      for (int nu = 0; nu < mu; ++nu) {
        M R, I(k, k);
        if (!chol_upper(G[nu], R)) continue;  // rank-deficient A M^-1 U: the reference would trsm with a partial factor; kept as is
        for (int a = 0; a < k; ++a) I(a, a) = 1.0;
        const M Rinv = solve_right_upper(I, R);
        Vec *blks[3] = {&rec.C, &pt, &rec.U};
        for (Vec *blk : blks)
          for (int c = k - 1; c >= 0; --c) {  // in place: column c only needs the old columns l <= c
            Vec bc = block(be, *blk, c, mu);
            GC(be.scal_col(nu, Rinv(c, c).real(), bc, bc));
            if (c > 0) GC(be.combine_col(nu, c, *blk, tok(Rinv.a, (size_t)c * k, c), 1.0, bc));
          }
      }
      GC(be.dots(k, rec.C, vi, hv));
      for (int nu = 0; nu < mu; ++nu) {
        GC(be.combine_col(nu, k, rec.C, &hv[(size_t)nu * k], -1.0, vi));
        GC(be.combine_col(nu, k, pt, &hv[(size_t)nu * k], 1.0, x));
      }
      be.release(pt);
    }
    GC(be.dots(1, vi, vi, hv));
    if (j == 1) {
      bool tiny = false;
      for (int nu = 0; nu < mu; ++nu) tiny = tiny || hb_real(hv[nu]) < 4.930380657631324e-32;  // eps^2 (GCRODR.hpp:143)
      if (tiny) {
        j = 0;
        break;
      }
    }
    for (int nu = 0; nu < mu; ++nu) {
      if (conv[nu] > 0) conv[nu] = 0;  // GCRODR.hpp:159
      Column &cn = col[nu];
      cn.R = M(m + 1, m);
      cn.save = M(m + 1, m);
      cn.B = M(std::max(k, 1), m);
      cn.cs.assign(m, zc(0.0));
      cn.sn.assign(m, 0.0);
      cn.s.assign(m + 1, zc(0.0));
      resnorm[nu] = std::sqrt(hb_real(hv[nu]));
      cn.s[i] = resnorm[nu];
      GC(be.scal_col(nu, 1.0 / resnorm[nu], vi, vi));
    }
    while (i < m && j <= max_it) {
      Vec cur = v(i), nxt = v(i + 1);
      GC(be.apply(cur, z));    // GCRODR.hpp:184
      GC(be.gmv(z, nxt));      // GCRODR.hpp:185
      if (haveU) {             // orthogonalization against C (GCRODR.hpp:191)
        GC(be.dots(k, rec.C, nxt, hv));
        for (int nu = 0; nu < mu; ++nu) {
          for (int c = 0; c < k; ++c) col[nu].B(c, i) = to_z(hv[(size_t)nu * k + c]);
          GC(be.combine_col(nu, k, rec.C, &hv[(size_t)nu * k], -1.0, nxt));
        }
      }
      const int cnt = i + 1 - shift;  // Arnoldi with `shift` (iterative.hpp:669-710), classical Gram-Schmidt: all products first
      Vec vs = v(shift);
      GC(be.dots(cnt, vs, nxt, hv));
      const std::vector<K> hcol(hv);
      for (int nu = 0; nu < mu; ++nu) GC(be.combine_col(nu, cnt, vs, &hcol[(size_t)nu * cnt], -1.0, nxt));
      GC(be.dots(1, nxt, nxt, hv));
      for (int nu = 0; nu < mu; ++nu) {
        Column &cn = col[nu];
        const double hn = std::sqrt(hb_real(hv[nu]));
        if (i < m - 1) GC(be.scal_col(nu, 1.0 / hn, nxt, nxt));
        for (int l = 0; l < cnt; ++l) {
          cn.save(l, i - shift) = to_z(hcol[(size_t)nu * cnt + l]);
          cn.R(shift + l, i) = cn.save(l, i - shift);
        }
        cn.save(cnt, i - shift) = hn;
        cn.R(i + 1, i) = hn;
        for (int l = shift; l < i; ++l) {  // previous rotations (iterative.hpp:690-696)
          const zc g = std::conj(cn.cs[l]) * cn.R(l, i) + cn.sn[l] * cn.R(l + 1, i);
          cn.R(l + 1, i) = -cn.sn[l] * cn.R(l, i) + cn.cs[l] * cn.R(l + 1, i);
          cn.R(l, i) = g;
        }
        const double delta = std::hypot(std::abs(cn.R(i, i)), std::abs(cn.R(i + 1, i)));
        cn.sn[i] = cn.R(i + 1, i).real() / delta;
        cn.cs[i] = cn.R(i, i) / delta;
        cn.R(i, i) = delta;
        cn.R(i + 1, i) = 0.0;
        cn.s[i + 1] = -cn.sn[i] * cn.s[i];
        cn.s[i] *= std::conj(cn.cs[i]);
      }
      ++i;
      bool all = true;
      for (int nu = 0; nu < mu; ++nu) {
        res[nu] = std::abs(col[nu].s[i]);
        if (conv[nu] == -m && converged(res[nu], nu)) conv[nu] = i;
        all = all && conv[nu] != -m;
      }
      if (all) {
        i += haveU ? m - k : m;  // GCRODR.hpp:209-212
        break;
      }
      ++j;
    }
    bool done;
    if (j != max_it + 1 && i == m)
      done = false;  // restart
    else {
      done = true;
      if (j == max_it + 1) {  // GCRODR.hpp:223-231
        int rem = haveU ? (max_it - m) % (m - k) : max_it % m;
        if (rem) {
          if (haveU) rem += k;
          for (int nu = 0; nu < mu; ++nu)
            if (conv[nu] < 0) conv[nu] = rem;
        }
      }
    }
    // updateSolRecycling (iterative.hpp:338-393): y = R^-1 s, x += M^-1 (U (C^H D r - B y) + V y)
    std::vector<K> cr;
    if (haveU && !p.same_system) {
      GC(be.dots(k, rec.C, v(shift), hv));
      cr = hv;
    }
    bool any = false;
    for (int nu = 0; nu < mu; ++nu) {
      GC(be.zero_col(nu, work));
      const int dim = std::abs(conv[nu]);
      if (dim == 0) continue;
      any = true;
      Column &cn = col[nu];
      std::vector<zc> y(std::max(dim - shift, 0));
      for (int r = dim - 1; r >= shift; --r) {
        zc acc = cn.s[r];
        for (int l = r + 1; l < dim; ++l) acc -= cn.R(r, l) * y[l - shift];
        y[r - shift] = acc / cn.R(r, r);
      }
      if (haveU) {
        std::vector<zc> su(k);
        for (int c = 0; c < k; ++c) {
          su[c] = p.same_system ? zc(0.0) : resnorm[nu] * to_z(cr[(size_t)nu * k + c]);  // iterative.hpp:351: `same` drops C^H D r
          for (int l = shift; l < dim; ++l) su[c] -= cn.B(c, l) * y[l - shift];
        }
        GC(be.combine_col(nu, k, rec.U, tok(su, 0, k), 1.0, work));
      }
      if (dim > shift) GC(be.combine_col(nu, dim - shift, v(shift), tok(y, 0, dim - shift), 1.0, work));
    }
    if (any) {
      GC(be.apply(work, z));
      for (int nu = 0; nu < mu; ++nu)
        if (conv[nu] != 0) GC(be.axpy_col(nu, 1.0, z, x));
    }
    if (i == m) {  // GCRODR.hpp:234-237: the last basis vector is normalised here (Arnoldi leaves it as is when i == m - 1)
      if (haveU) i -= k;
      Vec last = v(m);
      for (int nu = 0; nu < mu; ++nu) GC(be.scal_col(nu, 1.0 / col[nu].save(i, i - 1).real(), last, last));
    }
    if (p.same_system > 1) {
      // GCRODR.hpp:239: id[4] / 4 <= 1 guards both branches below -- the stored pair is used as is
    } else if (!haveU) {  // GCRODR.hpp:242-316: the first pair, from the harmonic Ritz vectors of GMRES(m)
      int dim = 0;
      bool first = true;
      for (int nu = 0; nu < mu; ++nu)
        if (conv[nu] != 0 && (first || conv[nu] < dim)) {
          dim = conv[nu];
          first = false;
        }
      dim = std::abs(dim);
      if (j < k || dim < k) k = dim;
      if (k > 0 && dim > 0) {
        GC(be.alloc(rec.U, k));
        const int rc_c = be.alloc(rec.C, k);
        if (rc_c < 0) {  // never leave half a pair behind
          be.release(rec.U);
          GC(rc_c);
        }
        rec.k = k;
        rec.mu = mu;
        rec.block = false;
        for (int nu = 0; nu < mu; ++nu) {
          Column &cn = col[nu];
          // f = H_m^-H e_m through the stored rotations (GCRODR.hpp:250-255), last column of H_m += h_{m+1,m}^2 f
          std::vector<zc> f(dim);
          zc hq = cn.cs[dim - 1] / cn.R(dim - 1, dim - 1);
          for (int l = dim - 1; l > 0; --l) {
            f[l] = cn.cs[l - 1] * hq;
            hq *= -cn.sn[l - 1];
          }
          f[0] = hq;
          M Hbar(dim + 1, dim), Hm(dim, dim);
          for (int c = 0; c < dim; ++c)
            for (int r = 0; r <= std::min(c + 1, dim); ++r) Hbar(r, c) = cn.save(r, c);
          for (int c = 0; c < dim; ++c)
            for (int r = 0; r < dim; ++r) Hm(r, c) = Hbar(r, c);
          const zc h2 = Hbar(dim, dim - 1) * Hbar(dim, dim - 1);
          for (int r = 0; r < dim; ++r) Hm(r, dim - 1) += h2 * f[r];
          std::vector<zc> th;
          M X;
          if (!eig(Hm, th, X)) {
            be.release(rec.U);  // no pair rather than a partly built one
            be.release(rec.C);
            rec.k = rec.mu = 0;
            cleanup();
            return ERR_EIGENSOLVER;
          }
          const M vr = select_columns(th, X, order(th, p.target), k);
          M Q, Rr;
          qr(mul(Hbar, vr), Q, Rr);
          const M Y = solve_right_upper(vr, Rr);
          for (int c = 0; c < k; ++c) {
            Vec uc = block(be, rec.U, c, mu), cc = block(be, rec.C, c, mu);
            GC(be.combine_col(nu, dim, v(0), tok(Y.a, (size_t)c * dim, dim), 1.0, uc));
            GC(be.combine_col(nu, dim + 1, v(0), tok(Q.a, (size_t)c * (dim + 1), dim + 1), 1.0, cc));
          }
        }
        haveU = true;
      }
    } else if (j > m - k) {  // GCRODR.hpp:317-430: new pair from [U, V]
      if (scratch.empty()) GC(be.alloc(scratch, k));
      for (int nu = 0; nu < mu; ++nu) {
        if (conv[nu] == 0) continue;
        Column &cn = col[nu];
        const int dim = std::abs(conv[nu]), diff = dim - k;
        if (diff < 1) continue;
        std::vector<double> Du(k, 1.0);
        M Wt(dim + 1, k);
        if (p.strategy == 0) {  // strategy A: U~ = U Du of unit D-norm columns, W^H D U~ with W = [C, v_k .. v_dim]
          for (int c = 0; c < k; ++c) {
            Vec uc = block(be, rec.U, c, mu);
            GC(be.dots(1, uc, uc, hv));
            Du[c] = 1.0 / std::sqrt(hb_real(hv[nu]));
            GC(be.dots(k, rec.C, uc, hv));
            for (int a = 0; a < k; ++a) Wt(a, c) = Du[c] * to_z(hv[(size_t)nu * k + a]);
            GC(be.dots(diff + 1, v(k), uc, hv));
            for (int a = 0; a <= diff; ++a) Wt(k + a, c) = Du[c] * to_z(hv[(size_t)nu * (diff + 1) + a]);
          }
        }
        M G(dim + 1, dim), Hbar(diff + 1, diff);
        for (int c = 0; c < diff; ++c)
          for (int r = 0; r <= std::min(c + 1, diff); ++r) Hbar(r, c) = cn.save(r, c);
        for (int c = 0; c < k; ++c) G(c, c) = Du[c];
        for (int c = 0; c < diff; ++c) {
          for (int r = 0; r < k; ++r) G(r, k + c) = cn.B(r, k + c);
          for (int r = 0; r <= diff; ++r) G(k + r, k + c) = Hbar(r, c);
        }
        M What(dim + 1, dim);  // W^H D [U~ V]: first k columns measured (strategy A) or taken as [I; 0] (strategy B, GCRODR.hpp:376-382); V_k.. are orthonormal
        for (int c = 0; c < k; ++c)
          for (int r = 0; r <= dim; ++r) What(r, c) = p.strategy == 0 ? Wt(r, c) : zc(r == c ? 1.0 : 0.0);
        for (int c = 0; c < diff; ++c) What(k + c, k + c) = 1.0;
        std::vector<zc> th;
        M X;
        if (!harmonic_pencil(G, What, th, X)) {
          cleanup();
          return ERR_EIGENSOLVER;
        }
        const M vr = select_columns(th, X, order(th, p.target), k);
        M Q, Rr;
        qr(mul(G, vr), Q, Rr);
        M Y = solve_right_upper(vr, Rr);
        for (int c = 0; c < k; ++c)
          for (int r = 0; r < k; ++r) Y(r, c) *= Du[r];
        // U <- [U, v_k .. v_{dim-1}] Y, C <- [C, v_k .. v_dim] Q   (through the scratch blocks: the outputs alias the inputs)
        for (int pass = 0; pass < 2; ++pass) {
          Vec &dst = pass == 0 ? rec.U : rec.C;
          const M &F = pass == 0 ? Y : Q;
          const int extra = pass == 0 ? diff : diff + 1;
          for (int c = 0; c < k; ++c) {
            Vec sc = block(be, scratch, c, mu);
            GC(be.zero_col(nu, sc));
            GC(be.combine_col(nu, k, dst, tok(F.a, (size_t)c * F.r, k), 1.0, sc));
            GC(be.combine_col(nu, extra, v(k), tok(F.a, (size_t)c * F.r + k, extra), 1.0, sc));
          }
          for (int c = 0; c < k; ++c) GC(be.scal_col(nu, 1.0, block(be, scratch, c, mu), block(be, dst, c, mu)));
        }
      }
    }
    if (done) break;
  }
  cleanup();
  *iterations = std::min(j, max_it);
  if (rel_residual)
    for (int nu = 0; nu < mu; ++nu) rel_residual[nu] = p.tol > 0.0 ? res[nu] / norm[nu] : res[nu];
  return 0;
}

// ------------------------------------------------------------------ block driver
int run_block(Backend &be, const Vec &b, const Vec &x, const Params &p, int *iterations, double *rel_residual) {
  const int mu = p.mu, max_it = p.max_it;
  const int m = std::min(p.restart, max_it);
  int k = std::min(m - 1, p.recycle);
  const int ldh = mu * (m + 1);
  Recycled &rec = be.recycled();
  if (rec.k > 0 && (rec.mu != mu || !rec.block)) {  // a pair of another shape (other mu, or one pair per column): dropped and rebuilt
    be.release(rec.U);
    be.release(rec.C);
    rec.k = rec.mu = 0;
  }
  bool haveU = rec.k > 0;
  if (haveU) k = rec.k;
  Vec V, z, work, pt, scratch;
  auto cleanup = [&]() {
    be.release(V);
    be.release(z);
    be.release(work);
    be.release(pt);
    be.release(scratch);
  };
  GC(be.alloc(V, m + 1));
  GC(be.alloc(z, 1));
  GC(be.alloc(work, 1));
  GC(be.alloc(scratch, 1));
  std::vector<K> hv, coef;
  std::vector<double> norm(mu), last(mu, 0.0);
  auto v = [&](int r) { return block(be, V, r, mu); };
  // rows [r0, r0 + nr) x columns [c0, c0 + nc) of a dense matrix as a contiguous K array (ld nr)
  auto sub = [&](const M &A, int r0, int nr, int c0, int nc) {
    coef.resize((size_t)nr * nc);
    for (int c = 0; c < nc; ++c)
      for (int r = 0; r < nr; ++r) coef[r + (size_t)c * nr] = from_z(A(r0 + r, c0 + c));
    return coef.data();
  };
  auto to_m = [&](int nr, int nc) {  // hv (nr x nc, ld nr) -> M
    M A(nr, nc);
    for (size_t i = 0; i < (size_t)nr * nc; ++i) A.a[i] = to_z(hv[i]);
    return A;
  };
  auto zero_blk = [&](const Vec &w) -> int {
    for (int nu = 0; nu < mu; ++nu) HB_GC_RET(be.zero_col(nu, w));
    return 0;
  };
  auto copy_blk = [&](const Vec &in, const Vec &out) -> int {
    for (int nu = 0; nu < mu; ++nu) HB_GC_RET(be.scal_col(nu, 1.0, in, out));
    return 0;
  };
  // w <- w T for a mu x mu matrix T (through the scratch block)
  auto rmul = [&](const Vec &w, const M &T) -> int {
    HB_GC_RET(zero_blk(scratch));
    HB_GC_RET(be.combine_blk(1, w, sub(T, 0, mu, 0, mu), 1.0, scratch));
    return copy_blk(scratch, w);
  };
  // CholQR of one block (IterativeMethod::QR, HPDDM_QR_CHOLQR): R upper; w <- w R^-1 if update.  1 = not positive definite
  auto cholqr = [&](const Vec &w, bool update, M &R) -> int {
    HB_GC_RET(be.gram(1, w, w, hv));
    if (!chol_upper(to_m(mu, mu), R)) return 1;
    if (update) {
      M I(mu, mu);
      for (int a = 0; a < mu; ++a) I(a, a) = 1.0;
      HB_GC_RET(rmul(w, solve_right_upper(I, R)));
    }
    return 0;
  };
  auto fallback = [&]() {  // GCRODR.hpp:896-906: rank-deficient block -> the non-block driver, from the current iterate
    cleanup();
    return run(be, b, x, p, iterations, rel_residual);
  };
  GC(be.start(b, x));
  GC(be.rhs_norms(b, norm));
  for (int nu = 0; nu < mu; ++nu)
    if (norm[nu] < 1e-12) norm[nu] = 1.0;
  int j = 1, dim = mu * m;
  while (j <= max_it) {
    const int shift = haveU ? k : 0;
    Vec v0 = v(shift);
    GC(be.gmv(x, v0));
    for (int nu = 0; nu < mu; ++nu) {
      GC(be.scal_col(nu, -1.0, v0, v0));
      GC(be.axpy_col(nu, 1.0, b, v0));
    }
    if (j == 1 && haveU) {  // GCRODR.hpp:515-556
      const int bK = mu * k;
      if (!p.same_system) {
        GC(be.alloc(pt, k));
        for (int c = 0; c < k; ++c) {
          GC(be.apply(block(be, rec.U, c, mu), block(be, pt, c, mu)));
          GC(be.gmv(block(be, pt, c, mu), block(be, rec.C, c, mu)));
        }
        M G(bK, bK), R;
        for (int c = 0; c < k; ++c) {
          GC(be.gram(k, rec.C, block(be, rec.C, c, mu), hv));
          for (int cc = 0; cc < mu; ++cc)
            for (int r = 0; r < bK; ++r) G(r, c * mu + cc) = to_z(hv[r + (size_t)cc * bK]);
        }
        if (chol_upper(G, R)) {
          M I(bK, bK);
          for (int a = 0; a < bK; ++a) I(a, a) = 1.0;
          const M Rinv = solve_right_upper(I, R);
          Vec *blks[3] = {&rec.C, &pt, &rec.U};
          for (Vec *blk : blks)
            for (int c = k - 1; c >= 0; --c) {  // block upper triangular: column block c only needs the old blocks l <= c
              GC(zero_blk(scratch));
              GC(be.combine_blk(c + 1, *blk, sub(Rinv, 0, (c + 1) * mu, c * mu, mu), 1.0, scratch));
              GC(copy_blk(scratch, block(be, *blk, c, mu)));
            }
        }
      }
      GC(be.gram(k, rec.C, v0, hv));
      const std::vector<K> hc(hv);
      GC(be.combine_blk(k, rec.C, hc.data(), -1.0, v0));
      if (!p.same_system) {
        GC(be.combine_blk(k, pt, hc.data(), 1.0, x));
        be.release(pt);
      } else {
        GC(zero_blk(work));
        GC(be.combine_blk(k, rec.U, hc.data(), 1.0, work));
        GC(be.apply(work, z));
        for (int nu = 0; nu < mu; ++nu) GC(be.axpy_col(nu, 1.0, z, x));
      }
    }
    M R0;
    {
      const int rc = cholqr(v0, true, R0);
      GC(rc);
      if (rc == 1) return fallback();
    }
    M H(ldh, m * mu), s(ldh, mu), save(ldh, m * mu), Bm(std::max(k, 1) * mu, m * mu);
    std::vector<zc> tau((size_t)m * mu, zc(0.0));
    for (int c = 0; c < mu; ++c)
      for (int r = 0; r <= c; ++r) s(shift * mu + r, c) = R0(r, c);
    int i = shift;
    bool conv_now = false;
    while (i < m && j <= max_it) {
      Vec w = v(i + 1);
      GC(be.apply(v(i), z));
      GC(be.gmv(z, w));
      if (haveU) {  // orthogonalisation against C (GCRODR.hpp:616)
        GC(be.gram(k, rec.C, w, hv));
        for (int c = 0; c < mu; ++c)
          for (int r = 0; r < k * mu; ++r) Bm(r, i * mu + c) = to_z(hv[r + (size_t)c * k * mu]);
        const std::vector<K> hb(hv);
        GC(be.combine_blk(k, rec.C, hb.data(), -1.0, w));
      }
      M Hc(ldh, mu);
      const int cnt = i + 1 - shift;  // BlockArnoldi (iterative.hpp:714-737): block classical Gram-Schmidt, all products first
      GC(be.gram(cnt, v(shift), w, hv));
      for (int c = 0; c < mu; ++c)
        for (int r = 0; r < cnt * mu; ++r) Hc(shift * mu + r, c) = to_z(hv[r + (size_t)c * cnt * mu]);
      {
        const std::vector<K> hp(hv);
        GC(be.combine_blk(cnt, v(shift), hp.data(), -1.0, w));
      }
      M R;
      {
        const int rc = cholqr(w, i < m - 1, R);
        GC(rc);
        if (rc == 1) return fallback();
      }
      for (int c = 0; c < mu; ++c)
        for (int r = 0; r <= c; ++r) Hc((i + 1) * mu + r, c) = R(r, c);
      for (int c = 0; c < mu; ++c)
        for (int r = 0; r < (i + 2 - shift) * mu; ++r) save(r, (i - shift) * mu + c) = Hc(shift * mu + r, c);
      for (int kk = shift; kk < i; ++kk) unm2r_left(true, 2 * mu, mu, mu, &H(kk * mu, kk * mu), ldh, &tau[(size_t)kk * mu], &Hc(kk * mu, 0), ldh);
      geqr2(2 * mu, mu, &Hc(i * mu, 0), ldh, &tau[(size_t)i * mu]);
      for (int c = 0; c < mu; ++c)
        for (int r = 0; r < ldh; ++r) H(r, i * mu + c) = Hc(r, c);
      unm2r_left(true, 2 * mu, mu, mu, &H(i * mu, i * mu), ldh, &tau[(size_t)i * mu], &s(i * mu, 0), ldh);
      ++i;
      bool all = true;
      for (int nu = 0; nu < mu; ++nu) {  // checkBlockConvergence<5>, t <= 1 (iterative.hpp:139-146): a partial norm, kept
        double nrm = 0.0;
        for (int r = 0; r <= nu; ++r) nrm += std::norm(s(i * mu + r, nu));
        last[nu] = std::sqrt(nrm);
        all = all && (p.tol > 0.0 ? last[nu] / norm[nu] <= p.tol : last[nu] <= -p.tol);
      }
      if (all) {
        dim = mu * i;
        conv_now = true;
        break;
      }
      ++j;
    }
    bool done;
    if (!conv_now && j != max_it + 1 && i == m)
      done = false;
    else {
      done = true;
      if (j == max_it + 1) {
        const int rem = haveU ? (max_it - m) % (m - k) : max_it % m;
        if (rem) dim = mu * (rem + (haveU ? k : 0));
      }
    }
    // updateSolRecycling, block form (iterative.hpp:372-391): Y = R^-1 S, x += M^-1 ([U V] [C^H D r - B Y; Y])
    const int da = dim - mu * shift;
    M Y(std::max(da, 0), mu);
    for (int c = 0; c < mu; ++c)
      for (int r = da - 1; r >= 0; --r) {
        zc acc = s(shift * mu + r, c);
        for (int l = r + 1; l < da; ++l) acc -= H(shift * mu + r, shift * mu + l) * Y(l, c);
        Y(r, c) = acc / H(shift * mu + r, shift * mu + r);
      }
    GC(zero_blk(work));
    if (haveU) {
      const int bK = mu * k;
      M top(bK, mu);
      if (!p.same_system) {  // C^H D (V_k R_0): the block residual at the start of the cycle
        GC(copy_blk(v(shift), z));
        GC(rmul(z, R0));
        GC(be.gram(k, rec.C, z, hv));
        top = to_m(bK, mu);
      }
      for (int c = 0; c < mu; ++c)
        for (int r = 0; r < bK; ++r)
          for (int l = 0; l < da; ++l) top(r, c) -= Bm(r, shift * mu + l) * Y(l, c);
      GC(be.combine_blk(k, rec.U, sub(top, 0, bK, 0, mu), 1.0, work));
    }
    if (da > 0) GC(be.combine_blk(da / mu, v(shift), sub(Y, 0, da, 0, mu), 1.0, work));
    GC(be.apply(work, z));
    for (int nu = 0; nu < mu; ++nu) GC(be.axpy_col(nu, 1.0, z, x));
    if (!conv_now && i == m) {  // GCRODR.hpp:658-661: the last block is normalised here (BlockArnoldi leaves it as is when i == m - 1)
      const int ir = m - shift;
      M Rl(mu, mu), I(mu, mu);
      for (int c = 0; c < mu; ++c) {
        I(c, c) = 1.0;
        for (int r = 0; r <= c; ++r) Rl(r, c) = save(ir * mu + r, (ir - 1) * mu + c);
      }
      GC(rmul(v(m), solve_right_upper(I, Rl)));
    }
    if (p.same_system > 1) {
      // GCRODR.hpp:663: id[4] / 4 <= 1 guards both branches below
    } else if (!haveU) {  // GCRODR.hpp:674-760: first pair
      const int db = std::min(j, m);
      if (db < k) k = db;
      const int bK = mu * k, df = db * mu;
      if (k > 0) {
        // last block column of H_m += first df rows of Q [R^-H E; 0], E = e_db (R_db^H R_x)   (GCRODR.hpp:682-692)
        M sb(ldh, mu);
        for (int c = 0; c < mu; ++c)
          for (int r = 0; r < mu; ++r) {
            zc acc(0.0);
            for (int l = 0; l < mu; ++l) acc += std::conj(save(db * mu + l, (db - 1) * mu + r)) * save(db * mu + l, (m - 1) * mu + c);
            sb((db - 1) * mu + r, c) = acc;
          }
        for (int c = 0; c < mu; ++c)  // trtrs("U", transc): R^H X = sb, forward substitution
          for (int r = 0; r < df; ++r) {
            zc acc = sb(r, c);
            for (int l = 0; l < r; ++l) acc -= std::conj(H(l, r)) * sb(l, c);
            sb(r, c) = acc / std::conj(H(r, r));
          }
        for (int kk = db - 1; kk >= 0; --kk) unm2r_left(false, 2 * mu, mu, mu, &H(kk * mu, kk * mu), ldh, &tau[(size_t)kk * mu], &sb(kk * mu, 0), ldh);
        M Hbar(df + mu, df), Mh(df, df);
        for (int c = 0; c < df; ++c)
          for (int r = 0; r < df + mu; ++r) Hbar(r, c) = save(r, c);
        for (int c = 0; c < df; ++c)
          for (int r = 0; r < df; ++r) Mh(r, c) = Hbar(r, c);
        for (int c = 0; c < mu; ++c)
          for (int r = 0; r < df; ++r) Mh(r, df - mu + c) += sb(r, c);
        std::vector<zc> th;
        M X;
        if (!eig(Mh, th, X)) {
          cleanup();
          return ERR_EIGENSOLVER;
        }
        const M vr = select_columns(th, X, order(th, p.target), bK);
        M Q, Rr;
        qr(mul(Hbar, vr), Q, Rr);
        const M Yc = solve_right_upper(vr, Rr);
        GC(be.alloc(rec.U, k));
        const int rc_c = be.alloc(rec.C, k);
        if (rc_c < 0) {
          be.release(rec.U);
          GC(rc_c);
        }
        rec.k = k;
        rec.mu = mu;
        rec.block = true;
        for (int c = 0; c < k; ++c) {
          GC(be.combine_blk(db, v(0), sub(Yc, 0, df, c * mu, mu), 1.0, block(be, rec.U, c, mu)));
          GC(be.combine_blk(db + 1, v(0), sub(Q, 0, df + mu, c * mu, mu), 1.0, block(be, rec.C, c, mu)));
        }
        haveU = true;
      }
    } else if (j > m - k) {  // GCRODR.hpp:761-884: new pair from [U, V]
      const int bK = mu * k, diff = dim - bK, nb = diff / mu;
      if (nb >= 1) {
        std::vector<double> Du(bK, 1.0);
        M Wt(dim + mu, bK);
        if (p.strategy == 0) {
          for (int c = 0; c < k; ++c) {
            Vec uc = block(be, rec.U, c, mu);
            GC(be.gram(1, uc, uc, hv));
            for (int cc = 0; cc < mu; ++cc) Du[c * mu + cc] = 1.0 / std::sqrt(hb_real(hv[cc + (size_t)cc * mu]));
            GC(be.gram(k, rec.C, uc, hv));
            for (int cc = 0; cc < mu; ++cc)
              for (int r = 0; r < bK; ++r) Wt(r, c * mu + cc) = Du[c * mu + cc] * to_z(hv[r + (size_t)cc * bK]);
            GC(be.gram(nb + 1, v(k), uc, hv));
            for (int cc = 0; cc < mu; ++cc)
              for (int r = 0; r < diff + mu; ++r) Wt(bK + r, c * mu + cc) = Du[c * mu + cc] * to_z(hv[r + (size_t)cc * (diff + mu)]);
          }
        }
        M G(dim + mu, dim), Hbar(diff + mu, diff);
        for (int c = 0; c < diff; ++c)
          for (int r = 0; r < diff + mu; ++r) Hbar(r, c) = save(r, c);
        for (int c = 0; c < bK; ++c) G(c, c) = Du[c];
        for (int c = 0; c < diff; ++c) {
          for (int r = 0; r < bK; ++r) G(r, bK + c) = Bm(r, k * mu + c);
          for (int r = 0; r < diff + mu; ++r) G(bK + r, bK + c) = Hbar(r, c);
        }
        M What(dim + mu, dim);
        for (int c = 0; c < bK; ++c)
          for (int r = 0; r < dim + mu; ++r) What(r, c) = p.strategy == 0 ? Wt(r, c) : zc(r == c ? 1.0 : 0.0);
        for (int c = 0; c < diff; ++c) What(bK + c, bK + c) = 1.0;
        std::vector<zc> th;
        M X;
        if (!harmonic_pencil(G, What, th, X)) {
          cleanup();
          return ERR_EIGENSOLVER;
        }
        const M vr = select_columns(th, X, order(th, p.target), bK);
        M Q, Rr;
        qr(mul(G, vr), Q, Rr);
        M Yc = solve_right_upper(vr, Rr);
        for (int c = 0; c < bK; ++c)
          for (int r = 0; r < bK; ++r) Yc(r, c) *= Du[r];
        // U <- [U, v_k ..] Yc, C <- [C, v_k .. v_{k+nb}] Q, through fresh blocks (the outputs alias the inputs)
        Vec Un, Cn;
        GC(be.alloc(Un, k));
        const int rc_c = be.alloc(Cn, k);
        if (rc_c < 0) {
          be.release(Un);
          GC(rc_c);
        }
        int rc = 0;
        for (int c = 0; c < k && rc >= 0; ++c) {
          rc = be.combine_blk(k, rec.U, sub(Yc, 0, bK, c * mu, mu), 1.0, block(be, Un, c, mu));
          if (rc >= 0) rc = be.combine_blk(nb, v(k), sub(Yc, bK, diff, c * mu, mu), 1.0, block(be, Un, c, mu));
          if (rc >= 0) rc = be.combine_blk(k, rec.C, sub(Q, 0, bK, c * mu, mu), 1.0, block(be, Cn, c, mu));
          if (rc >= 0) rc = be.combine_blk(nb + 1, v(k), sub(Q, bK, diff + mu, c * mu, mu), 1.0, block(be, Cn, c, mu));
        }
        if (rc < 0) {
          be.release(Un);
          be.release(Cn);
          GC(rc);
        }
        be.release(rec.U);
        be.release(rec.C);
        rec.U = Un;
        rec.C = Cn;
      }
    }
    if (done) break;
  }
  cleanup();
  *iterations = std::min(j, max_it);
  if (rel_residual)
    for (int nu = 0; nu < mu; ++nu) rel_residual[nu] = p.tol > 0.0 ? last[nu] / norm[nu] : last[nu];
  return 0;
}

}  // namespace gcro
}  // namespace hb
