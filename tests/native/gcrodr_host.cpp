// CPU harness of the product's GCRO-DR driver (hpddm_b200/csrc/hb_gcrodr.cpp): the same gcro::run() that libhpddm_b200.so drives
// with device kernels, here on plain host arrays.  The preconditioner apply, the operator product, Schwarz::start and the
// right-hand-side norms are callbacks (the tests pass the oracle's operator), the vector primitives are the loops below.
// TEST INFRASTRUCTURE ONLY -- not linked into the product.  Built by tests/test_cpu_gcrodr.py for both scalar types.
#include <cstring>

#include "../../hpddm_b200/csrc/hb_gcrodr.h"

using namespace hb;
using gcro::Vec;

extern "C" {
typedef int (*op_cb)(void *user, K *const *in, K *const *out);                 // out = op(in), P pointers of n_r x mu
typedef int (*norm_cb)(void *user, K *const *b, double *out);                  // ||b_nu||_D, mu values
}

namespace {

struct HostBackend : gcro::Backend {
  int P = 0, mu = 1;
  std::vector<int> n;
  std::vector<const double *> d;
  op_cb cb_apply = nullptr, cb_gmv = nullptr, cb_start = nullptr;
  norm_cb cb_norm = nullptr;
  void *user = nullptr;
  gcro::Recycled *rec = nullptr;
  long calls_apply = 0, calls_gmv = 0;

  size_t subs() const override { return (size_t)P; }
  int64_t rows(size_t q) const override { return n[q]; }
  int alloc(Vec &v, int blocks) override {
    v.assign(P, nullptr);
    for (int q = 0; q < P; ++q) {
      const size_t len = (size_t)n[q] * mu * blocks;
      v[q] = new K[len > 0 ? len : 1];
      for (size_t i = 0; i < len; ++i) v[q][i] = mk(0.0);
    }
    return 0;
  }
  void release(Vec &v) override {
    for (K *p : v) delete[] p;
    v.clear();
  }
  gcro::Recycled &recycled() override { return *rec; }
  int start(const Vec &b, const Vec &x) override {  // x <- Schwarz::start(b, x): in = b then x, out = x
    std::vector<K *> in(b.begin(), b.end());
    in.insert(in.end(), x.begin(), x.end());
    return cb_start(user, in.data(), x.data());
  }
  int rhs_norms(const Vec &b, std::vector<double> &out) override {
    out.assign(mu, 0.0);
    return cb_norm(user, b.data(), out.data());
  }
  int apply(const Vec &in, const Vec &out) override {
    ++calls_apply;
    return cb_apply(user, in.data(), out.data());
  }
  int gmv(const Vec &in, const Vec &out) override {
    ++calls_gmv;
    return cb_gmv(user, in.data(), out.data());
  }
  int dots(int count, const Vec &basis, const Vec &w, std::vector<K> &out) override {
    out.assign((size_t)count * mu, mk(0.0));
    for (int q = 0; q < P; ++q)
      for (int nu = 0; nu < mu; ++nu)
        for (int r = 0; r < count; ++r) {
          const K *v = basis[q] + (size_t)r * mu * n[q] + (size_t)nu * n[q];
          const K *x = w[q] + (size_t)nu * n[q];
          K acc = mk(0.0);
          for (int i = 0; i < n[q]; ++i) acc = acc + d[q][i] * (hb_conj(v[i]) * x[i]);
          out[(size_t)nu * count + r] = out[(size_t)nu * count + r] + acc;
        }
    return 0;
  }
  int combine_col(int nu, int count, const Vec &basis, const K *coef, double alpha, const Vec &w) override {
    for (int q = 0; q < P; ++q) {
      K *x = w[q] + (size_t)nu * n[q];
      for (int i = 0; i < n[q]; ++i) {
        K acc = mk(0.0);
        for (int r = 0; r < count; ++r) acc = acc + basis[q][(size_t)r * mu * n[q] + (size_t)nu * n[q] + i] * coef[r];
        x[i] = x[i] + alpha * acc;
      }
    }
    return 0;
  }
  int scal_col(int nu, double a, const Vec &in, const Vec &out) override {
    for (int q = 0; q < P; ++q)
      for (int i = 0; i < n[q]; ++i) out[q][(size_t)nu * n[q] + i] = a * in[q][(size_t)nu * n[q] + i];
    return 0;
  }
  int axpy_col(int nu, double a, const Vec &in, const Vec &out) override {
    for (int q = 0; q < P; ++q)
      for (int i = 0; i < n[q]; ++i) out[q][(size_t)nu * n[q] + i] = out[q][(size_t)nu * n[q] + i] + a * in[q][(size_t)nu * n[q] + i];
    return 0;
  }
  int zero_col(int nu, const Vec &out) override {
    for (int q = 0; q < P; ++q)
      for (int i = 0; i < n[q]; ++i) out[q][(size_t)nu * n[q] + i] = mk(0.0);
    return 0;
  }
  int gram(int count, const Vec &basis, const Vec &w, std::vector<K> &out) override {
    const int rows = count * mu;
    out.assign((size_t)rows * mu, mk(0.0));
    for (int q = 0; q < P; ++q)
      for (int col = 0; col < mu; ++col)
        for (int r = 0; r < rows; ++r) {
          K acc = mk(0.0);
          for (int i = 0; i < n[q]; ++i) acc = acc + d[q][i] * (hb_conj(basis[q][(size_t)r * n[q] + i]) * w[q][(size_t)col * n[q] + i]);
          out[r + (size_t)col * rows] = out[r + (size_t)col * rows] + acc;
        }
    return 0;
  }
  int combine_blk(int count, const Vec &basis, const K *coef, double alpha, const Vec &w) override {
    const int rows = count * mu;
    for (int q = 0; q < P; ++q)
      for (int col = 0; col < mu; ++col)
        for (int i = 0; i < n[q]; ++i) {
          K acc = mk(0.0);
          for (int r = 0; r < rows; ++r) acc = acc + basis[q][(size_t)r * n[q] + i] * coef[r + (size_t)col * rows];
          w[q][(size_t)col * n[q] + i] = w[q][(size_t)col * n[q] + i] + alpha * acc;
        }
    return 0;
  }
};

}  // namespace

extern "C" {

// state: in/out opaque handle of the recycled pair (nullptr on the first solve); counts[0 / 1] = apply / GMV calls of this solve
int gcrodr_host_run(int P, const int *n, const double *const *d, op_cb apply, op_cb gmv, op_cb start, norm_cb norms, void *user, K *const *b, K *const *x, int mu,
                    int restart, int recycle, int max_it, double tol, int target, int strategy, int same_system, int *iterations, double *rel_residual, void **state,
                    long *counts, int block) {
  HostBackend be;
  be.P = P;
  be.mu = mu;
  be.n.assign(n, n + P);
  be.d.assign(d, d + P);
  be.cb_apply = apply;
  be.cb_gmv = gmv;
  be.cb_start = start;
  be.cb_norm = norms;
  be.user = user;
  if (!*state) *state = new gcro::Recycled();
  be.rec = static_cast<gcro::Recycled *>(*state);
  gcro::Params p;
  p.mu = mu;
  p.restart = restart;
  p.recycle = recycle;
  p.max_it = max_it;
  p.tol = tol;
  p.target = target;
  p.strategy = strategy;
  p.same_system = same_system;
  const Vec bv(b, b + P), xv(x, x + P);
  const int rc = block ? gcro::run_block(be, bv, xv, p, iterations, rel_residual) : gcro::run(be, bv, xv, p, iterations, rel_residual);
  if (counts) {
    counts[0] = be.calls_apply;
    counts[1] = be.calls_gmv;
  }
  return rc;
}

int gcrodr_host_recycled_dim(void *state) { return state ? static_cast<gcro::Recycled *>(state)->k : 0; }

void gcrodr_host_free(void *state) {
  gcro::Recycled *r = static_cast<gcro::Recycled *>(state);
  if (!r) return;
  for (K *p : r->U) delete[] p;
  for (K *p : r->C) delete[] p;
  delete r;
}

int gcrodr_host_eig(int n, const double *a, double *w, double *x) { return gcro::eig_general(n, a, w, x) ? 0 : -1; }
}
