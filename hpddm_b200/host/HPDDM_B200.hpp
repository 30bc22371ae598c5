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
 * Only K = double is implemented (FP64 is the metric's precision; complex is config 5, out of scope).
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

#ifndef HPDDM_NUMBERING
  #define HPDDM_NUMBERING 'C'
#endif

namespace HPDDM {
template <class K>
class MatrixCSR;  // include/HPDDM_matrix.hpp:156-165 (n_, m_, nnz_, ia_, ja_, a_, sym_)

namespace b200 {
inline void check(int rc, const char *what) {
  if (rc < 0) {
    std::fprintf(stderr, "[hpddm_b200] %s failed (%d): %s\n", what, rc, hpddm_b200_last_error());
    throw std::runtime_error(std::string(what) + ": " + hpddm_b200_last_error());
  }
}
/* one context per process, created on first use (after MPI_Init / fork).  Device: HPDDM_B200_DEVICE,
 * else the local MPI rank exported by the launcher, else 0. */
inline hpddm_b200_ctx *context(bool fresh = false) {
  static hpddm_b200_ctx *ctx = nullptr;
  if (fresh || !ctx) {
    int dev = 0;
    for (const char *v : {"HPDDM_B200_DEVICE", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "SLURM_LOCALID", "LOCAL_RANK"})
      if (const char *e = std::getenv(v)) {
        dev = std::atoi(e);
        break;
      }
    hpddm_b200_ctx *c = nullptr;
    check(hpddm_b200_ctx_create(dev, &c), "hpddm_b200_ctx_create");
    if (fresh) return c;
    ctx = c;
  }
  return ctx;
}
}  // namespace b200

/* ------------------------------------------------------------------ 1. SUBDOMAIN plugin */
template <class K>
class B200Sub {
  static_assert(std::is_same<K, double>::value, "hpddm_b200: only K = double is implemented");

private:
  hpddm_b200_ctx *ctx_;
  hpddm_b200_sub *sub_;
  int             n_;

public:
  B200Sub() : ctx_(), sub_(), n_() { }
  B200Sub(const B200Sub &) = delete;
  ~B200Sub() { dtor(); }
  static constexpr char numbering_ = 'C';
  void                  dtor()
  {
    if (sub_) hpddm_b200_sub_destroy(sub_);
    sub_ = nullptr;
    if (ctx_) hpddm_b200_ctx_destroy(ctx_);
    ctx_ = nullptr;
  }
  /* SUBDOMAIN::numfact (e.g. include/HPDDM_SuiteSparse.hpp:264-371) */
  template <char N = HPDDM_NUMBERING>
  void numfact(MatrixCSR<K> *const &A, bool = false, K *const & = nullptr)
  {
    dtor();
    ctx_ = b200::context(true);  // private context: solver objects are independent of each other
    b200::check(hpddm_b200_sub_create(ctx_, 0, &sub_), "hpddm_b200_sub_create");
    n_ = A->n_;
    b200::check(hpddm_b200_sub_set_matrix(sub_, A->n_, A->nnz_, A->ia_, A->ja_, A->a_, A->sym_ ? 1 : 0, N), "hpddm_b200_sub_set_matrix");
    if (const char *g = std::getenv("HPDDM_B200_GRID")) {  // optional "nx,ny,nz[,dof]" ordering hint
      int nx = 0, ny = 0, nz = 1, dof = 1;
      if (std::sscanf(g, "%d,%d,%d,%d", &nx, &ny, &nz, &dof) >= 2 && (long long)nx * ny * nz * dof == A->n_) hpddm_b200_sub_set_grid_hint(sub_, nx, ny, nz, dof);
    }
    b200::check(hpddm_b200_sub_numfact(sub_, HPDDM_B200_PRCNDTNR_GE, 0, 0, nullptr, nullptr, nullptr, 0, 'C'), "hpddm_b200_sub_numfact");
  }
  template <char = HPDDM_NUMBERING>
  int inertia(MatrixCSR<K> *const &)
  {
    return 0;
  }
  unsigned short deficiency() const { return 0; }
  /* SUBDOMAIN::solve (include/HPDDM_SuiteSparse.hpp:388-423): host pointers in/out */
  void solve(K *const x, const unsigned short &n = 1) const { b200::check(hpddm_b200_sub_solve(sub_, x, x, n, HPDDM_B200_HOST), "hpddm_b200_sub_solve"); }
  void solve(const K *const b, K *const x, const unsigned short &n = 1) const { b200::check(hpddm_b200_sub_solve(sub_, b, x, n, HPDDM_B200_HOST), "hpddm_b200_sub_solve"); }
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
  static_assert(std::is_same<K, double>::value, "hpddm_b200: only K = double is implemented");

public:
  typedef K scalar_type;
  /* Prcndtnr (include/HPDDM_enum.hpp) */
  enum class Prcndtnr : char { NO = HPDDM_B200_PRCNDTNR_NO, SY = HPDDM_B200_PRCNDTNR_SY, GE = HPDDM_B200_PRCNDTNR_GE, OS = HPDDM_B200_PRCNDTNR_OS, OG = HPDDM_B200_PRCNDTNR_OG };

private:
  hpddm_b200_ctx *ctx_;
  hpddm_b200_sub *sub_;
  MatrixCSR<K>   *a_;
  const double   *d_;
  int             dof_, rank_, size_, nu_, correction_;
  std::vector<std::pair<unsigned short, std::vector<int>>> map_;

public:
  B200Schwarz() : ctx_(b200::context()), sub_(), a_(), d_(), dof_(), rank_(), size_(1), nu_(), correction_(HPDDM_B200_CORRECTION_NONE) { }
  B200Schwarz(const B200Schwarz &) = delete;
  ~B200Schwarz()
  {
    if (sub_) hpddm_b200_sub_destroy(sub_);
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
      if (rank == 0) b200::check(hpddm_b200_nccl_unique_id(id), "hpddm_b200_nccl_unique_id");
      bcast(static_cast<void *>(id));
      b200::check(hpddm_b200_ctx_comm_init(ctx_, id, rank, size), "hpddm_b200_ctx_comm_init");
    }
  }
  /* Subdomain::initialize(a, o, r) (include/HPDDM_subdomain.hpp:165-236) */
  template <class Neighbor, class Mapping>
  void initialize(MatrixCSR<K> *const &a, const Neighbor &o, const Mapping &r)
  {
    a_   = a;
    dof_ = a->n_;
    if (!sub_) b200::check(hpddm_b200_sub_create(ctx_, rank_, &sub_), "hpddm_b200_sub_create");
    b200::check(hpddm_b200_sub_set_matrix(sub_, a->n_, a->nnz_, a->ia_, a->ja_, a->a_, a->sym_ ? 1 : 0, HPDDM_NUMBERING), "hpddm_b200_sub_set_matrix");
    std::vector<int> ranks(o.begin(), o.end()), sizes, idx;
    unsigned short   i = 0;
    map_.clear();
    for (const auto &m : r) {
      sizes.push_back(static_cast<int>(m.size()));
      idx.insert(idx.end(), m.begin(), m.end());
      map_.emplace_back(static_cast<unsigned short>(ranks[i++]), std::vector<int>(m.begin(), m.end()));
    }
    b200::check(hpddm_b200_sub_set_neighbors(sub_, static_cast<int>(ranks.size()), ranks.data(), sizes.data(), idx.data()), "hpddm_b200_sub_set_neighbors");
  }
  void setGridHint(int nx, int ny, int nz = 1, int dof = 1) { b200::check(hpddm_b200_sub_set_grid_hint(sub_, nx, ny, nz, dof), "hpddm_b200_sub_set_grid_hint"); }
  /* Schwarz::multiplicityScaling (include/HPDDM_schwarz.hpp:381-404) */
  void multiplicityScaling(double *const d) const
  {
    double *arr[1] = {d};
    b200::check(hpddm_b200_multiplicity_scaling(ctx_, arr), "hpddm_b200_multiplicity_scaling");
  }
  /* Schwarz::initialize(d) (schwarz.hpp:178): d stays owned by the caller */
  void initialize(double *const &d)
  {
    d_ = d;
    b200::check(hpddm_b200_sub_set_scaling(sub_, d), "hpddm_b200_sub_set_scaling");
  }
  /* Schwarz::callNumfact (schwarz.hpp:337-368) */
  template <char N = HPDDM_NUMBERING>
  void callNumfact(MatrixCSR<K> *const &A = nullptr, int prcndtnr = HPDDM_B200_PRCNDTNR_GE)
  {
    if (A) b200::check(hpddm_b200_sub_numfact(sub_, prcndtnr, A->n_, A->nnz_, A->ia_, A->ja_, A->a_, A->sym_ ? 1 : 0, N), "hpddm_b200_sub_numfact");
    else b200::check(hpddm_b200_sub_numfact(sub_, prcndtnr, 0, 0, nullptr, nullptr, nullptr, 0, 'C'), "hpddm_b200_sub_numfact");
  }
  /* Preconditioner::setVectors (include/HPDDM_preconditioner.hpp:358-362): ev[0] contiguous n x nu; ownership stays with the caller here */
  void setVectors(K **const &ev, unsigned short nu)
  {
    nu_ = nu;
    b200::check(hpddm_b200_sub_set_vectors(sub_, *ev, nu), "hpddm_b200_sub_set_vectors");
  }
  /* Schwarz::solveGEVP<EIGENSOLVER>(A_Neumann) (schwarz.hpp:665-715): GenEO vectors computed on the GPU; the template
   * parameter of the reference (the eigensolver plugin) has no meaning here and is accepted for source compatibility */
  template <template <class> class Eps = B200Sub>
  void solveGEVP(MatrixCSR<K> *const &A, unsigned short nu = 20, double tol = 1.0e-6)
  {
    b200::check(hpddm_b200_sub_solve_gevp(sub_, A->n_, A->nnz_, A->ia_, A->ja_, A->a_, A->sym_ ? 1 : 0, HPDDM_NUMBERING, nu, tol, 0, nullptr), "hpddm_b200_sub_solve_gevp");
    nu_ = nu;
  }
  /* Schwarz::buildTwo (schwarz.hpp:440-495) */
  template <unsigned short excluded = 0, class Comm = int>
  int buildTwo(const Comm & = Comm(), int correction = HPDDM_B200_CORRECTION_DEFLATED)
  {
    correction_ = correction;
    b200::check(hpddm_b200_build_coarse(ctx_), "hpddm_b200_build_coarse");
    return 0;
  }
  void setCorrection(int correction) { correction_ = correction; }
  /* Schwarz::start / Subdomain::end (schwarz.hpp:496-514, subdomain.hpp:289) */
  template <bool excluded = false>
  bool start(const K *const b, K *const x, const unsigned short &mu = 1) const
  {
    const double *bb[1] = {b};
    double       *xx[1] = {x};
    b200::check(hpddm_b200_start(ctx_, bb, xx, mu, HPDDM_B200_HOST), "hpddm_b200_start");
    return false;
  }
  void end(const bool = true) const { hpddm_b200_end(ctx_); }
  /* Schwarz::exchange<allocate> (schwarz.hpp:180-188) and Subdomain::exchange */
  template <bool allocate = false>
  void exchange(K *const x, const unsigned short &mu = 1) const
  {
    double *xx[1] = {x};
    b200::check(hpddm_b200_exchange(ctx_, xx, mu, 1, HPDDM_B200_HOST), "hpddm_b200_exchange");
  }
  /* Schwarz::apply<excluded>(in, out, mu, work) (schwarz.hpp:527-612); `in` is never clobbered */
  template <bool excluded = false>
  int apply(const K *const in, K *const out, const unsigned short &mu = 1, K * = nullptr) const
  {
    const double *ii[1] = {in};
    double       *oo[1] = {out};
    return hpddm_b200_apply(ctx_, ii, oo, mu, correction_, HPDDM_B200_HOST);
  }
  /* Schwarz::deflation<excluded, transpose> (schwarz.hpp:1602-1622) */
  template <bool excluded, bool transpose = false>
  void deflation(const K *const in, K *const out, const unsigned short &mu) const
  {
    const double *ii[1] = {in};
    double       *oo[1] = {out};
    b200::check(hpddm_b200_deflation(ctx_, ii, oo, mu, HPDDM_B200_HOST), "hpddm_b200_deflation");
  }
  /* Schwarz::GMV (schwarz.hpp:726-747) */
  int GMV(const K *const in, K *const out, const int &mu = 1) const
  {
    const double *ii[1] = {in};
    double       *oo[1] = {out};
    return hpddm_b200_gmv(ctx_, ii, oo, mu, HPDDM_B200_HOST);
  }
  /* Schwarz::computeResidual (schwarz.hpp:761-803), l2 norm: storage[2*nu] = ||f||_D, [2*nu+1] = ||Ax-f||_D */
  void computeResidual(const K *const x, const K *const f, double *const storage, const unsigned short mu = 1) const
  {
    std::vector<K> tmp(static_cast<std::size_t>(mu) * dof_);
    GMV(x, tmp.data(), mu);
    for (std::size_t i = 0; i < tmp.size(); ++i) tmp[i] -= f[i];
    std::vector<double> r(mu), b(mu);
    const double       *t[1] = {tmp.data()}, *ff[1] = {f};
    b200::check(hpddm_b200_dot(ctx_, t, t, mu, r.data(), HPDDM_B200_HOST), "hpddm_b200_dot");
    b200::check(hpddm_b200_dot(ctx_, ff, ff, mu, b.data(), HPDDM_B200_HOST), "hpddm_b200_dot");
    for (unsigned short nu = 0; nu < mu; ++nu) {
      storage[2 * nu]     = std::sqrt(b[nu]);
      storage[2 * nu + 1] = std::sqrt(r[nu]);
    }
  }
  /* accessors used by the Krylov drivers (include/HPDDM_GMRES.hpp:40-62, HPDDM_iterative.hpp:441-468) */
  const double *getScaling() const { return d_; }
  int           getDof() const { return dof_; }
  unsigned short getLocal() const { return static_cast<unsigned short>(nu_); }
  const MatrixCSR<K> *getMatrix() const { return a_; }
  const std::vector<std::pair<unsigned short, std::vector<int>>> &getMap() const { return map_; }
  std::unordered_map<unsigned int, K> boundaryConditions() const { return std::unordered_map<unsigned int, K>(); }
  hpddm_b200_ctx *context() const { return ctx_; }
  hpddm_b200_sub *handle() const { return sub_; }
};
}  // namespace HPDDM

#ifdef B200SUB
  #ifdef SUBDOMAIN
    #undef SUBDOMAIN
  #endif
  #define SUBDOMAIN HPDDM::B200Sub
#endif
#endif  // HPDDM_B200_HPP_
