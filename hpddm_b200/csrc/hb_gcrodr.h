// GCRO-DR driver of libhpddm_b200.so (IterativeMethod::GCRODR, include/HPDDM_GCRODR.hpp:35-444; the Krylov method of BASELINE
// config 5).  The driver (hb_gcrodr.cpp) is host code written against the small vector-space interface below: everything that
// touches a vector of length n goes through a Backend, everything it computes itself is small dense algebra of order
// restart + 1 (Hessenberg matrices, Givens rotations, the harmonic Ritz eigenproblems).  libhpddm_b200.so instantiates the
// interface with the kernels of the device Krylov drivers (hb_krylov.cu: basis, recycled pair and every product stay in HBM);
// tests/native/gcrodr_host.cpp instantiates it with plain host arrays and the oracle's operator so that the driver logic is
// checked on a machine without a GPU against goldens of the unmodified reference.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "hb_scalar.h"

namespace hb {
namespace gcro {

// A block vector: per local subdomain q a column-major n_q x mu array (the reference's layout, GCRODR.hpp:63: ldv = mu * n).
// A *basis* is a Vec pointing at the first of several block vectors stored back to back: block r starts at + r * mu * n_q.
typedef std::vector<K *> Vec;

// recycled pair C = A M^-1 U, C^H D C = I, kept between solves (the reference keeps it in A.storage(), HPDDM_option.hpp:445-454)
struct Recycled {
  Vec U, C;  // k blocks each
  int k = 0, mu = 0;
  bool block = false;  // built by the block driver: ONE pair of mu k columns for all right-hand sides (BGCRODR), not one pair per column
};

struct Backend {
  virtual ~Backend() {}
  virtual size_t subs() const = 0;
  virtual int64_t rows(size_t q) const = 0;
  virtual int alloc(Vec &v, int blocks) = 0;  // zero-filled
  virtual void release(Vec &v) = 0;
  virtual Recycled &recycled() = 0;
  virtual int start(const Vec &b, const Vec &x) = 0;                               // Schwarz::start (schwarz.hpp:496-514)
  virtual int rhs_norms(const Vec &b, std::vector<double> &out) = 0;         // ||b||_D per column (initializeNorm, iterative.hpp:455-468)
  virtual int apply(const Vec &in, const Vec &out) = 0;                            // out = M^-1 in, all mu columns
  virtual int gmv(const Vec &in, const Vec &out) = 0;                            // out = A in
  // out[nu * count + r] = sum over subdomains and processes of sum_i d_i conj(basis_r[i, nu]) w[i, nu]
  virtual int dots(int count, const Vec &basis, const Vec &w, std::vector<K> &out) = 0;
  // w[:, nu] += alpha * sum_r basis_r[:, nu] coef[r]     (coef: host, count entries)
  virtual int combine_col(int nu, int count, const Vec &basis, const K *coef, double alpha, const Vec &w) = 0;
  virtual int scal_col(int nu, double a, const Vec &in, const Vec &out) = 0;      // out[:, nu] = a in[:, nu]
  virtual int axpy_col(int nu, double a, const Vec &in, const Vec &out) = 0;      // out[:, nu] += a in[:, nu]
  virtual int zero_col(int nu, const Vec &out) = 0;
  // block products of the block driver: `basis` read as the n x (count mu) matrix X of its columns (ld n per subdomain)
  // out (count mu x mu, column-major, ld count mu) = sum over subdomains and processes of X^H D w
  virtual int gram(int count, const Vec &basis, const Vec &w, std::vector<K> &out) = 0;
  // w (n x mu) += alpha * X coef          (coef: host, count mu x mu, column-major, ld count mu)
  virtual int combine_blk(int count, const Vec &basis, const K *coef, double alpha, const Vec &w) = 0;
};

constexpr int ERR_EIGENSOLVER = -100;  // the QR iteration of a harmonic Ritz problem did not converge / singular pencil

struct Params {
  int mu = 1, restart = 40, recycle = 0, max_it = 100;
  double tol = 1e-6;
  int target = 0;    // HPDDM_RECYCLE_TARGET_SM .. LI (HPDDM_define.hpp:169-174)
  int strategy = 0;  // HPDDM_RECYCLE_STRATEGY_A / B (HPDDM_define.hpp:166-167)
  // -hpddm_recycle_same_system as IterativeMethod::options reads it (iterative.hpp:217; the reference raises the option from 1 to 2
  // after a converged solve, GCRODR.hpp:435): 0 = the operator may have changed, C = A M^-1 U is recomputed when a later solve starts;
  // 1 = same operator: the pair is built / updated, C^H D r is taken as zero in the solution update (iterative.hpp:351); 2 = same
  // operator, the stored pair is used as is (no product with A M^-1, no update)
  int same_system = 0;
};

// b, x: one block vector each (x holds the initial guess).  Returns 0 or a negative error code of the backend; iterations as the
// reference returns them (min(j, max_it)); rel_residual[nu] = last |s| / ||b_nu|| (absolute when tol < 0).  recycle <= 0 is the
// caller's business (the reference switches to GMRES, GCRODR.hpp:50-55).
int run(Backend &be, const Vec &b, const Vec &x, const Params &p, int *iterations, double *rel_residual);

// IterativeMethod::BGCRODR (GCRODR.hpp:445-907) with the reference defaults: one block Krylov space and one recycled pair of mu k columns
// for all right-hand sides, block classical Gram-Schmidt, CholQR, Householder-reduced block Hessenberg matrix (LAPACK conventions: the
// reference's convergence test reads individual entries of the transformed block residual, iterative.hpp:139-146), no deflation of
// right-hand sides.  A rank-deficient block continues with run(), as the reference continues with GCRODR (GCRODR.hpp:896-906).
int run_block(Backend &be, const Vec &b, const Vec &x, const Params &p, int *iterations, double *rel_residual);

// eigenvalues and unit-norm right eigenvectors of a general complex matrix (column-major n x n, interleaved re / im): complex
// Schur form by Householder reduction + shifted QR, back-substitution.  Exposed for the CPU tests.  Returns false if the QR
// iteration does not converge.
bool eig_general(int n, const double *a_interleaved, double *w_interleaved, double *x_interleaved);

}  // namespace gcro
}  // namespace hb
