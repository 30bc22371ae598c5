/*
 * HPDDM_B200.hpp -- header-only C++ host layer over libhpddm_b200.so (include/hpddm_b200.h).
 *
 * Two seams, both keeping the reference's names and signatures:
 *
 *  1. HPDDM::B200Sub<K>  -- a SUBDOMAIN solver plugin (concept of include/HPDDM_SuiteSparse.hpp:224-424,
 *     include/HPDDM_MUMPS.hpp:206-318: numbering_, dtor, numfact<N>(MatrixCSR<K>*, bool, K*), inertia,
 *     deficiency, solve(K*, n), solve(const K*, K*, n)).  Define B200SUB (like MUMPSSUB / SUITESPARSESUB)
 *     and the UNMODIFIED reference (HPDDM::Schwarz, examples/schwarz.cpp, every Krylov driver) runs its
 *     local factorisations and triangular solves on the GPU.  See INTEGRATION.md for the two-line
 *     include a maintainer adds to include/HPDDM.hpp; oracle/ref_build builds exactly that.
 *
 *  2. HPDDM::B200Schwarz<K> -- mirror of HPDDM::Schwarz (include/HPDDM_schwarz.hpp) whose whole hot
 *     path (apply / deflation / exchange / GMV / start / end) runs on the GPU; one object = one
 *     subdomain = one process = one GPU, halo + coarse gather over NCCL.  It satisfies the duck-typed
 *     Operator concept of the Krylov drivers (include/HPDDM_GMRES.hpp:57-62,113-117) so
 *     IterativeMethod::solve(A, f, sol, mu, comm) works on it unchanged.
 *
 * K = double (C ABI hpddm_b200_*, include/hpddm_b200.h) and K = std::complex<double> (hpddm_b200z_*,
 * include/hpddm_b200z.h; the reference's FORCE_COMPLEX build, BASELINE config 5) are implemented; b200::Api<K>
 * selects the entry points of one scalar type.
 */
#ifndef HPDDM_B200_HPP_
#define HPDDM_B200_HPP_

#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "hpddm_b200.h"
#include "hpddm_b200z.h"

#ifndef HPDDM_NUMBERING
  #define HPDDM_NUMBERING 'C'
#endif

namespace HPDDM {
template <class K>
class MatrixCSR;  // include/HPDDM_matrix.hpp:156-165 (n_, m_, nnz_, ia_, ja_, a_, sym_)

namespace b200 {
/* Api<K>: the C entry points of one scalar type behind common names (K = double -> hpddm_b200_*, K = std::complex<double> -> hpddm_b200z_*) */
template <class K>
struct Api;
#define HPDDM_B200_DEFINE_API(K_, P_)                                                                                                              \
  template <>                                                                                                                                      \
  struct Api<K_> {                                                                                                                                 \
    typedef P_##_ctx ctx_t;                                                                                                                        \
    typedef P_##_sub sub_t;                                                                                                                        \
    static const char *last_error() { return P_##_last_error(); }                                                                                  \
    static int ctx_create(int dev, ctx_t **c) { return P_##_ctx_create(dev, c); }                                                                  \
    static int ctx_destroy(ctx_t *c) { return P_##_ctx_destroy(c); }                                                                               \
    static int nccl_unique_id(void *id) { return P_##_nccl_unique_id(id); }                                                                        \
    static int ctx_comm_init(ctx_t *c, const void *id, int r, int n) { return P_##_ctx_comm_init(c, id, r, n); }                                   \
    static int sub_create(ctx_t *c, int r, sub_t **s) { return P_##_sub_create(c, r, s); }                                                         \
    static int sub_destroy(sub_t *s) { return P_##_sub_destroy(s); }                                                                               \
    static int sub_set_matrix(sub_t *s, int n, int nnz, const int *ia, const int *ja, const K_ *a, int sym, char nb)                               \
    {                                                                                                                                              \
      return P_##_sub_set_matrix(s, n, nnz, ia, ja, a, sym, nb);                                                                                   \
    }                                                                                                                                              \
    static int sub_set_neighbors(sub_t *s, int c, const int *r, const int *sz, const int *idx) { return P_##_sub_set_neighbors(s, c, r, sz, idx); } \
    static int sub_set_scaling(sub_t *s, const double *d) { return P_##_sub_set_scaling(s, d); }                                                   \
    static int sub_set_grid_hint(sub_t *s, int nx, int ny, int nz, int dof) { return P_##_sub_set_grid_hint(s, nx, ny, nz, dof); }                 \
    static int multiplicity_scaling(ctx_t *c, double *const *d) { return P_##_multiplicity_scaling(c, d); }                                        \
    static int sub_numfact(sub_t *s, int t, int n, int nnz, const int *ia, const int *ja, const K_ *a, int sym, char nb)                           \
    {                                                                                                                                              \
      return P_##_sub_numfact(s, t, n, nnz, ia, ja, a, sym, nb);                                                                                   \
    }                                                                                                                                              \
    static int sub_set_vectors(sub_t *s, const K_ *Z, int nu) { return P_##_sub_set_vectors(s, Z, nu); }                                           \
    static int sub_solve_gevp(sub_t *s, int n, int nnz, const int *ia, const int *ja, const K_ *a, int sym, char nb, int nu, double tol, int it,   \
                              double *ev)                                                                                                          \
    {                                                                                                                                              \
      return P_##_sub_solve_gevp(s, n, nnz, ia, ja, a, sym, nb, nu, tol, it, ev);                                                                  \
    }                                                                                                                                              \
    static int sub_get_vectors(sub_t *s, K_ *Z, int *nu) { return P_##_sub_get_vectors(s, Z, nu); }                                                \
    static int sub_stats(sub_t *s, hpddm_b200_stats *st) { return P_##_sub_stats(s, st); }                                                         \
    static int build_coarse(ctx_t *c) { return P_##_build_coarse(c); }                                                                             \
    static int ctx_comm_init_host(ctx_t *c, int r, int n, P_##_allgather_fn f, void *u) { return P_##_ctx_comm_init_host(c, r, n, f, u); }         \
    static int ctx_transport(ctx_t *c) { return P_##_ctx_transport(c); }                                                                           \
    static long long launch_count(ctx_t *c) { return static_cast<long long>(P_##_ctx_launch_count(c)); }                                           \
    static int sub_boundary_conditions(sub_t *s, int *idx, K_ *val, int *cnt) { return P_##_sub_boundary_conditions(s, idx, val, cnt); }           \
    static int rhs_norm(ctx_t *c, const K_ *const *b, int mu, double *nrm, int w) { return P_##_rhs_norm(c, b, mu, nrm, w); }                      \
    static int compute_residual(ctx_t *c, const K_ *const *x, const K_ *const *f, double *st, int mu, int nrm, int w)                              \
    {                                                                                                                                              \
      return P_##_compute_residual(c, x, f, st, mu, nrm, w);                                                                                       \
    }                                                                                                                                              \
    static int start(ctx_t *c, const K_ *const *b, K_ *const *x, int mu, int w) { return P_##_start(c, b, x, mu, w); }                             \
    static int end(ctx_t *c) { return P_##_end(c); }                                                                                               \
    static int apply(ctx_t *c, const K_ *const *in, K_ *const *out, int mu, int corr, int w) { return P_##_apply(c, in, out, mu, corr, w); }        \
    static int deflation(ctx_t *c, const K_ *const *in, K_ *const *out, int mu, int w) { return P_##_deflation(c, in, out, mu, w); }                \
    static int exchange(ctx_t *c, K_ *const *x, int mu, int scaled, int w) { return P_##_exchange(c, x, mu, scaled, w); }                          \
    static int gmv(ctx_t *c, const K_ *const *in, K_ *const *out, int mu, int w) { return P_##_gmv(c, in, out, mu, w); }                            \
    static int sub_solve(sub_t *s, const K_ *b, K_ *x, int mu, int w) { return P_##_sub_solve(s, b, x, mu, w); }                                    \
    static int dot(ctx_t *c, const K_ *const *x, const K_ *const *y, int mu, K_ *res, int w) { return P_##_dot(c, x, y, mu, res, w); }              \
    static int solve(ctx_t *c, const K_ *const *b, K_ *const *x, int mu, int corr, int m, int it, double tol, int w, int *its, double *res)        \
    {                                                                                                                                              \
      return P_##_solve(c, b, x, mu, corr, m, it, tol, w, its, res);                                                                               \
    }                                                                                                                                              \
    static int solve_bgmres(ctx_t *c, const K_ *const *b, K_ *const *x, int mu, int corr, int m, int it, double tol, int w, int *its, double *res) \
    {                                                                                                                                              \
      return P_##_solve_bgmres(c, b, x, mu, corr, m, it, tol, w, its, res);                                                                        \
    }                                                                                                                                              \
    static int solve_cg(ctx_t *c, const K_ *const *b, K_ *const *x, int mu, int corr, int it, double tol, int w, int *its, double *res)            \
    {                                                                                                                                              \
      return P_##_solve_cg(c, b, x, mu, corr, it, tol, w, its, res);                                                                               \
    }                                                                                                                                              \
    static int solve_gcrodr(ctx_t *c, const K_ *const *b, K_ *const *x, int mu, int corr, int m, int k, int target, int strategy, int same,        \
                            int it, double tol, int w, int *its, double *res)                                                                      \
    {                                                                                                                                              \
      return P_##_solve_gcrodr(c, b, x, mu, corr, m, k, target, strategy, same, it, tol, w, its, res);                                             \
    }                                                                                                                                              \
    static int solve_bgcrodr(ctx_t *c, const K_ *const *b, K_ *const *x, int mu, int corr, int m, int k, int target, int strategy, int same,       \
                             int it, double tol, int w, int *its, double *res)                                                                     \
    {                                                                                                                                              \
      return P_##_solve_bgcrodr(c, b, x, mu, corr, m, k, target, strategy, same, it, tol, w, its, res);                                            \
    }                                                                                                                                              \
    static int recycle_destroy(ctx_t *c) { return P_##_recycle_destroy(c); }                                                                       \
  }
HPDDM_B200_DEFINE_API(double, hpddm_b200);
HPDDM_B200_DEFINE_API(std::complex<double>, hpddm_b200z);
#undef HPDDM_B200_DEFINE_API

template <class K>
inline void check(int rc, const char *what) {
  if (rc < 0) {
    std::fprintf(stderr, "[hpddm_b200] %s failed (%d): %s\n", what, rc, Api<K>::last_error());
    throw std::runtime_error(std::string(what) + ": " + Api<K>::last_error());
  }
}
/* one context per process and scalar type, created on first use (after MPI_Init / fork).  Device: HPDDM_B200_DEVICE,
 * else the local MPI rank exported by the launcher, else 0. */
template <class K>
inline typename Api<K>::ctx_t *context(bool fresh = false) {
  static typename Api<K>::ctx_t *ctx = nullptr;
  if (fresh || !ctx) {
    int dev = 0;
    for (const char *v : {"HPDDM_B200_DEVICE", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", "LOCAL_RANK"})
      if (const char *e = std::getenv(v)) {
        dev = std::atoi(e);
        break;
      }
    typename Api<K>::ctx_t *c = nullptr;
    check<K>(Api<K>::ctx_create(dev, &c), "ctx_create");
    if (fresh) return c;
    ctx = c;
  }
  return ctx;
}
template <class K>
inline long long launch_count(typename Api<K>::ctx_t *c) { return Api<K>::launch_count(c); }
inline double real_part(double v) { return v; }
inline double real_part(const std::complex<double> &v) { return v.real(); }
}  // namespace b200

/* ------------------------------------------------------------------ 1. SUBDOMAIN plugin */
template <class K>
class B200Sub {
  static_assert(std::is_same<K, double>::value || std::is_same<K, std::complex<double>>::value, "hpddm_b200: K must be double or std::complex<double>");
  typedef b200::Api<K> A_;

private:
  typename A_::ctx_t *ctx_;
  typename A_::sub_t *sub_;
  int                 n_;

public:
  B200Sub() : ctx_(), sub_(), n_() { }
  B200Sub(const B200Sub &) = delete;
  ~B200Sub() { dtor(); }
  static constexpr char numbering_ = 'C';
  void                  dtor()
  {
    if (sub_) A_::sub_destroy(sub_);
    sub_ = nullptr;
    if (ctx_) A_::ctx_destroy(ctx_);
    ctx_ = nullptr;
  }
  /* SUBDOMAIN::numfact (e.g. include/HPDDM_SuiteSparse.hpp:264-371) */
  template <char N = HPDDM_NUMBERING>
  void numfact(MatrixCSR<K> *const &A, bool = false, K *const & = nullptr)
  {
    dtor();
    ctx_ = b200::context<K>(true);  // private context: solver objects are independent of each other
    b200::check<K>(A_::sub_create(ctx_, 0, &sub_), "sub_create");
    n_ = A->n_;
    b200::check<K>(A_::sub_set_matrix(sub_, A->n_, A->nnz_, A->ia_, A->ja_, A->a_, A->sym_ ? 1 : 0, N), "sub_set_matrix");
    if (const char *g = std::getenv("HPDDM_B200_GRID")) {  // optional "nx,ny,nz[,dof]" ordering hint
      int nx = 0, ny = 0, nz = 1, dof = 1;
      if (std::sscanf(g, "%d,%d,%d,%d", &nx, &ny, &nz, &dof) >= 2 && (long long)nx * ny * nz * dof == A->n_) A_::sub_set_grid_hint(sub_, nx, ny, nz, dof);
    }
    b200::check<K>(A_::sub_numfact(sub_, HPDDM_B200_PRCNDTNR_GE, 0, 0, nullptr, nullptr, nullptr, 0, 'C'), "sub_numfact");
  }
  template <char = HPDDM_NUMBERING>
  int inertia(MatrixCSR<K> *const &)
  {
    return 0;
  }
  unsigned short deficiency() const { return 0; }
  /* SUBDOMAIN::solve (include/HPDDM_SuiteSparse.hpp:388-423): host pointers in/out */
  void solve(K *const x, const unsigned short &n = 1) const { b200::check<K>(A_::sub_solve(sub_, x, x, n, HPDDM_B200_HOST), "sub_solve"); }
  void solve(const K *const b, K *const x, const unsigned short &n = 1) const { b200::check<K>(A_::sub_solve(sub_, b, x, n, HPDDM_B200_HOST), "sub_solve"); }
};

/* ------------------------------------------------------------------ 2. Schwarz mirror */
namespace b200 {
/* stand-alone stand-in for HPDDM::OptionsPrefix<K> (include/HPDDM_option.hpp:389-460) */
template <class K>
class NoPrefix {
  std::string prefix_;

public:
  typedef K scalar_type;
  void        setPrefix(const std::string &p) { prefix_ = p; }
  std::string prefix() const { return prefix_; }
  std::string prefix(const std::string &opt) const { return prefix_ + opt; }
};
}  // namespace b200

/* Base: inside the reference pass HPDDM::OptionsPrefix<K> (what Subdomain<K> derives from,
 * include/HPDDM_subdomain.hpp:47) so that the recycling Krylov methods find storage()/k()/allocate(). */
template <class K, class Base = b200::NoPrefix<K>>
class B200Schwarz : public Base {
  static_assert(std::is_same<K, double>::value || std::is_same<K, std::complex<double>>::value, "hpddm_b200: K must be double or std::complex<double>");
  typedef b200::Api<K> A_;

public:
  typedef K scalar_type;
  /* Prcndtnr (include/HPDDM_enum.hpp) */
  enum class Prcndtnr : char { NO = HPDDM_B200_PRCNDTNR_NO, SY = HPDDM_B200_PRCNDTNR_SY, GE = HPDDM_B200_PRCNDTNR_GE, OS = HPDDM_B200_PRCNDTNR_OS, OG = HPDDM_B200_PRCNDTNR_OG };

private:
  typename A_::ctx_t *ctx_;
  typename A_::sub_t *sub_;
  MatrixCSR<K>   *a_;
  const double   *d_;
  int             dof_, rank_, size_, nu_, correction_;
  std::vector<std::pair<unsigned short, std::vector<int>>> map_;

public:
  B200Schwarz() : ctx_(b200::context<K>()), sub_(), a_(), d_(), dof_(), rank_(), size_(1), nu_(), correction_(HPDDM_B200_CORRECTION_NONE) { }
  B200Schwarz(const B200Schwarz &) = delete;
  ~B200Schwarz()
  {
    if (sub_) A_::sub_destroy(sub_);
  }
  /* NCCL bootstrap for the hot path (replaces Subdomain::communicator_ there): `bcast` broadcasts
   * 128 bytes from rank 0, e.g. [&](void* p){ MPI_Bcast(p, 128, MPI_BYTE, 0, MPI_COMM_WORLD); } */
  template <class Bcast>
  void setCommunicator(int rank, int size, Bcast bcast)
  {
    rank_ = rank;
    size_ = size;
    if (size > 1) {
      char id[128];
      if (rank == 0) b200::check<K>(A_::nccl_unique_id(id), "nccl_unique_id");
      bcast(static_cast<void *>(id));
      b200::check<K>(A_::ctx_comm_init(ctx_, id, rank, size), "ctx_comm_init");
    }
  }
  /* Host-bootstrapped communicator (hpddm_b200_ctx_comm_init_host): the control plane goes through the host program's own all-gather,
   * e.g. [](const void *s, void *r, size_t bytes, void *comm) { return MPI_Allgather(s, bytes, MPI_BYTE, r, bytes, MPI_BYTE, *(MPI_Comm *)comm); },
   * the hot-path collectives over the library's peer-memory fabric -- no NCCL needed (what the full seam does with Subdomain::communicator_) */
  void setCommunicatorHost(int rank, int size, int (*allgather)(const void *, void *, size_t, void *), void *user)
  {
    rank_ = rank;
    size_ = size;
    if (size > 1) b200::check<K>(A_::ctx_comm_init_host(ctx_, rank, size, allgather, user), "ctx_comm_init_host");
  }
  /* Subdomain::initialize(a, o, r) (include/HPDDM_subdomain.hpp:165-236) */
  template <class Neighbor, class Mapping>
  void initialize(MatrixCSR<K> *const &a, const Neighbor &o, const Mapping &r)
  {
    a_   = a;
    dof_ = a->n_;
    if (!sub_) b200::check<K>(A_::sub_create(ctx_, rank_, &sub_), "sub_create");
    b200::check<K>(A_::sub_set_matrix(sub_, a->n_, a->nnz_, a->ia_, a->ja_, a->a_, a->sym_ ? 1 : 0, HPDDM_NUMBERING), "sub_set_matrix");
    std::vector<int> ranks(o.begin(), o.end()), sizes, idx;
    unsigned short   i = 0;
    map_.clear();
    for (const auto &m : r) {
      sizes.push_back(static_cast<int>(m.size()));
      idx.insert(idx.end(), m.begin(), m.end());
      map_.emplace_back(static_cast<unsigned short>(ranks[i++]), std::vector<int>(m.begin(), m.end()));
    }
    b200::check<K>(A_::sub_set_neighbors(sub_, static_cast<int>(ranks.size()), ranks.data(), sizes.data(), idx.data()), "sub_set_neighbors");
  }
  void setGridHint(int nx, int ny, int nz = 1, int dof = 1) { b200::check<K>(A_::sub_set_grid_hint(sub_, nx, ny, nz, dof), "sub_set_grid_hint"); }
  /* Schwarz::multiplicityScaling (include/HPDDM_schwarz.hpp:381-404) */
  void multiplicityScaling(double *const d) const
  {
    double *arr[1] = {d};
    b200::check<K>(A_::multiplicity_scaling(ctx_, arr), "multiplicity_scaling");
  }
  /* Schwarz::initialize(d) (schwarz.hpp:178): d stays owned by the caller */
  void initialize(double *const &d)
  {
    d_ = d;
    b200::check<K>(A_::sub_set_scaling(sub_, d), "sub_set_scaling");
  }
  /* Schwarz::callNumfact (schwarz.hpp:337-368) */
  template <char N = HPDDM_NUMBERING>
  void callNumfact(MatrixCSR<K> *const &A = nullptr, int prcndtnr = HPDDM_B200_PRCNDTNR_GE)
  {
    if (A) b200::check<K>(A_::sub_numfact(sub_, prcndtnr, A->n_, A->nnz_, A->ia_, A->ja_, A->a_, A->sym_ ? 1 : 0, N), "sub_numfact");
    else b200::check<K>(A_::sub_numfact(sub_, prcndtnr, 0, 0, nullptr, nullptr, nullptr, 0, 'C'), "sub_numfact");
  }
  /* Preconditioner::setVectors (include/HPDDM_preconditioner.hpp:358-362): ev[0] contiguous n x nu; ownership stays with the caller here */
  void setVectors(K **const &ev, unsigned short nu)
  {
    nu_ = nu;
    b200::check<K>(A_::sub_set_vectors(sub_, *ev, nu), "sub_set_vectors");
  }
  /* Schwarz::solveGEVP<EIGENSOLVER>(A_Neumann) (schwarz.hpp:665-715): GenEO vectors computed on the GPU; the template
   * parameter of the reference (the eigensolver plugin) has no meaning here and is accepted for source compatibility */
  template <template <class> class Eps = B200Sub>
  void solveGEVP(MatrixCSR<K> *const &A, unsigned short nu = 20, double tol = 1.0e-6)
  {
    b200::check<K>(A_::sub_solve_gevp(sub_, A->n_, A->nnz_, A->ia_, A->ja_, A->a_, A->sym_ ? 1 : 0, HPDDM_NUMBERING, nu, tol, 0, nullptr), "sub_solve_gevp");
    nu_ = nu;
  }
  /* Schwarz::buildTwo (schwarz.hpp:440-495) */
  template <unsigned short excluded = 0, class Comm = int>
  int buildTwo(const Comm & = Comm(), int correction = HPDDM_B200_CORRECTION_DEFLATED)
  {
    correction_ = correction;
    b200::check<K>(A_::build_coarse(ctx_), "build_coarse");
    return 0;
  }
  void setCorrection(int correction) { correction_ = correction; }
  /* Schwarz::start / Subdomain::end (schwarz.hpp:496-514, subdomain.hpp:289) */
  template <bool excluded = false>
  bool start(const K *const b, K *const x, const unsigned short &mu = 1) const
  {
    const K *bb[1] = {b};
    K       *xx[1] = {x};
    b200::check<K>(A_::start(ctx_, bb, xx, mu, HPDDM_B200_HOST), "start");
    return false;
  }
  void end(const bool = true) const { A_::end(ctx_); }
  /* Schwarz::exchange<allocate> (schwarz.hpp:180-188) and Subdomain::exchange */
  template <bool allocate = false>
  void exchange(K *const x, const unsigned short &mu = 1) const
  {
    K *xx[1] = {x};
    b200::check<K>(A_::exchange(ctx_, xx, mu, 1, HPDDM_B200_HOST), "exchange");
  }
  /* Schwarz::apply<excluded>(in, out, mu, work) (schwarz.hpp:527-612); `in` is never clobbered */
  template <bool excluded = false>
  int apply(const K *const in, K *const out, const unsigned short &mu = 1, K * = nullptr) const
  {
    const K *ii[1] = {in};
    K       *oo[1] = {out};
    return A_::apply(ctx_, ii, oo, mu, correction_, HPDDM_B200_HOST);
  }
  /* Schwarz::deflation<excluded, transpose> (schwarz.hpp:1602-1622) */
  template <bool excluded, bool transpose = false>
  void deflation(const K *const in, K *const out, const unsigned short &mu) const
  {
    const K *ii[1] = {in};
    K       *oo[1] = {out};
    b200::check<K>(A_::deflation(ctx_, ii, oo, mu, HPDDM_B200_HOST), "deflation");
  }
  /* Schwarz::GMV (schwarz.hpp:726-747) */
  int GMV(const K *const in, K *const out, const int &mu = 1) const
  {
    const K *ii[1] = {in};
    K       *oo[1] = {out};
    return A_::gmv(ctx_, ii, oo, mu, HPDDM_B200_HOST);
  }
  /* Schwarz::computeResidual (schwarz.hpp:761-803): storage[2*nu] = ||f|| (entries larger than EPS * PEN divided by PEN),
   * storage[2*nu+1] = ||A x - f|| off the boundary-condition rows; norm = HPDDM_COMPUTE_RESIDUAL_L2 (0) / _L1 (1) / _LINFTY (2) */
  void computeResidual(const K *const x, const K *const f, double *const storage, const unsigned short mu = 1, const unsigned short norm = 0) const
  {
    const K *xx[1] = {x}, *ff[1] = {f};
    b200::check<K>(A_::compute_residual(ctx_, xx, ff, storage, mu, norm, HPDDM_B200_HOST), "compute_residual");
  }
  /* Device-resident counterpart of IterativeMethod::solve(A, f, sol, mu, comm) (include/HPDDM_iterative.hpp:1013-1111): the Krylov
   * vectors never leave HBM.  method: 0 = GMRES (HPDDM_KRYLOV_METHOD_GMRES), 1 = BGMRES, 2 = CG, 4 = GCRODR, 5 = BGCRODR (with `recycle` harmonic Ritz
   * vectors kept between calls, default target / strategy) -- the values of -hpddm_krylov_method (include/HPDDM_define.hpp).  Returns the
   * iteration count like the reference's drivers (negative = error). */
  int solve(const K *const f, K *const x, const unsigned short mu = 1, int method = 0, int restart = 40, int max_it = 100, double tol = 1.0e-6, int recycle = 0) const
  {
    const K *bb[1] = {f};
    K       *xx[1] = {x};
    int      it = 0, rc;
    if (method == 4) rc = A_::solve_gcrodr(ctx_, bb, xx, mu, correction_, restart, recycle, 0, 0, 0, max_it, tol, HPDDM_B200_HOST, &it, nullptr);
    else if (method == 5) rc = A_::solve_bgcrodr(ctx_, bb, xx, mu, correction_, restart, recycle, 0, 0, 0, max_it, tol, HPDDM_B200_HOST, &it, nullptr);
    else if (method == 1) rc = A_::solve_bgmres(ctx_, bb, xx, mu, correction_, restart, max_it, tol, HPDDM_B200_HOST, &it, nullptr);
    else if (method == 2) rc = A_::solve_cg(ctx_, bb, xx, mu, correction_, max_it, tol, HPDDM_B200_HOST, &it, nullptr);
    else rc = A_::solve(ctx_, bb, xx, mu, correction_, restart, max_it, tol, HPDDM_B200_HOST, &it, nullptr);
    return rc < 0 ? rc : it;
  }
  /* Preconditioner::getVectors (include/HPDDM_preconditioner.hpp:366): the deflation vectors as one contiguous n x nu block
   * (ev[0], ev[i] = ev[0] + i n), downloaded from the device (e.g. after solveGEVP); the caller owns the returned arrays */
  K **getVectors() const
  {
    int nu = 0;
    b200::check<K>(A_::sub_get_vectors(sub_, nullptr, &nu), "sub_get_vectors");
    if (nu == 0) return nullptr;
    K **ev = new K *[nu];
    *ev    = new K[static_cast<std::size_t>(nu) * dof_];
    for (int i = 1; i < nu; ++i) ev[i] = *ev + static_cast<std::size_t>(i) * dof_;
    b200::check<K>(A_::sub_get_vectors(sub_, *ev, &nu), "sub_get_vectors");
    return ev;
  }
  /* Subdomain::statistics analogue (include/HPDDM_subdomain.hpp:405-454) */
  void statistics() const
  {
    hpddm_b200_stats st;
    b200::check<K>(A_::sub_stats(sub_, &st), "sub_stats");
    std::printf(" --- subdomain %d: %lld dofs, %lld nnz, factor %lld entries (%.2f GB, %s), %lld fronts / %lld levels, halo %lld, nu %lld\n", rank_, (long long)st.n,
                (long long)st.nnz_a, (long long)st.nnz_factor, st.factor_bytes * 1e-9, st.symmetric ? "LL^T" : "LU", (long long)st.fronts, (long long)st.levels,
                (long long)st.halo, (long long)st.nu);
  }
  /* accessors used by the Krylov drivers (include/HPDDM_GMRES.hpp:40-62, HPDDM_iterative.hpp:441-468) */
  const double *getScaling() const { return d_; }
  int           getDof() const { return dof_; }
  unsigned short getLocal() const { return static_cast<unsigned short>(nu_); }
  const MatrixCSR<K> *getMatrix() const { return a_; }
  const std::vector<std::pair<unsigned short, std::vector<int>>> &getMap() const { return map_; }
  /* Subdomain::boundaryConditions (include/HPDDM_subdomain.hpp:327-336): penalised / identity rows with their diagonal values --
   * what IterativeMethod::initializeNorm needs to rescale penalised entries of the right-hand side (iterative.hpp:461-468) */
  std::unordered_map<unsigned int, K> boundaryConditions() const
  {
    int cnt = 0;
    b200::check<K>(A_::sub_boundary_conditions(sub_, nullptr, nullptr, &cnt), "sub_boundary_conditions");
    std::vector<int> idx(cnt);
    std::vector<K>   val(cnt);
    if (cnt) b200::check<K>(A_::sub_boundary_conditions(sub_, idx.data(), val.data(), &cnt), "sub_boundary_conditions");
    std::unordered_map<unsigned int, K> map;
    for (int i = 0; i < cnt; ++i) map[static_cast<unsigned int>(idx[i])] = val[i];
    return map;
  }
  typename A_::ctx_t *context() const { return ctx_; }
  typename A_::sub_t *handle() const { return sub_; }
};
}  // namespace HPDDM

#ifdef B200SUB
  #ifdef SUBDOMAIN
    #undef SUBDOMAIN
  #endif
  #define SUBDOMAIN HPDDM::B200Sub
#endif
#endif  // HPDDM_B200_HPP_
