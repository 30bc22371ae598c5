/*
 * HPDDM_B200_schwarz.hpp -- full-path seam: HPDDM::Schwarz<HPDDM::B200Sub, CoarseSolver, S, K>.
 *
 * A partial specialisation of the reference's OWN class template (include/HPDDM_schwarz.hpp:86) for
 * Solver = HPDDM::B200Sub: with -DB200SUB (SUBDOMAIN = HPDDM::B200Sub) the line
 *
 *     HPDDM::Schwarz<SUBDOMAIN, COARSEOPERATOR, symCoarse, K> A;          (examples/schwarz.cpp:90)
 *
 * of the UNMODIFIED driver instantiates this class, and every call it makes -- A.Subdomain::initialize, A.multiplicityScaling,
 * A.initialize(d), A.exchange<true>, A.setVectors / A.solveGEVP<EIGENSOLVER>, A.super::initialize(nu), A.buildTwo(comm),
 * A.callNumfact, IterativeMethod::solve(A, ...) -> A.start / apply / GMV / end, A.computeResidual, A.getCommunicator -- keeps
 * its name, signature and meaning while the whole hot path (apply, deflation, exchange, GMV, start, the local factorisation,
 * the coarse operator and, with -DEIGENSOLVER=HPDDM::B200Eps, the GenEO eigensolve) runs on the GPU behind the C ABI of
 * include/hpddm_b200.h.  Host-side state that the reference keeps in Subdomain<K> (matrix, neighbour map, MPI communicator,
 * option prefix, recycling storage of GCRO-DR) is the reference's own base class: this class derives from HPDDM::Subdomain<K>.
 *
 * Integration: included at the end of include/HPDDM.hpp under #ifdef B200SCHWARZ; in this repository
 * "-include HPDDM_B200.hpp -include HPDDM_B200_schwarz.hpp" on the compiler command line stands in for that patch
 * (oracle/ref_build/Makefile -> oracle/_ref/schwarz_b200_full, built from the reference's examples/schwarz.cpp as it is).
 *
 * Communication: the control plane (CUDA-IPC handles, coarse sizes) uses Subdomain::communicator_ through
 * hpddm_b200_ctx_comm_init_host (MPI_Allgather); the data plane is the library's peer-memory fabric.  HPDDM_B200_NCCL=1
 * additionally bootstraps NCCL (needed when peers are not IPC-reachable).  One MPI rank = one subdomain = one context; several
 * ranks may share a GPU (device = HPDDM_B200_DEVICE, else rank modulo the number of visible devices).
 */
#ifndef HPDDM_B200_SCHWARZ_HPP_
#define HPDDM_B200_SCHWARZ_HPP_

#include "HPDDM_B200.hpp"
#include <HPDDM.hpp>

#if !HPDDM_SCHWARZ || !HPDDM_MPI
  #error "HPDDM_B200_schwarz.hpp needs the reference's Schwarz layer (HPDDM_SCHWARZ && HPDDM_MPI)"
#endif

namespace HPDDM {
/* EIGENSOLVER placeholder: A.solveGEVP<EIGENSOLVER>(MatNeumann) of examples/schwarz.cpp:115 compiles when EIGENSOLVER is defined;
 * the specialisation below ignores the template argument and runs hpddm_b200_sub_solve_gevp (GPU GenEO, hb_geneo.cu) */
template <class K>
class B200Eps { };

namespace b200 {
/* plays the role of HPDDM::Preconditioner (include/HPDDM_preconditioner.hpp:98-109): owner of the deflation vectors and of the
 * "is there a coarse operator" state; `A.super::initialize(nu)` (examples/schwarz.cpp:124) lands here */
template <class K>
class Prcndtnr : public Subdomain<K> {
protected:
  K            **ev_;
  unsigned short nu_;
  bool           co_;

public:
  typedef Subdomain<K> super;
  Prcndtnr() : ev_(), nu_(), co_() { }
  Prcndtnr(const Prcndtnr &) = delete;
  ~Prcndtnr()
  {
    if (ev_) delete[] *ev_; /* one contiguous block (preconditioner.hpp:393-394) */
    delete[] ev_;
  }
  /* Preconditioner::initialize(deflation) (preconditioner.hpp:328-334) */
  void initialize(const unsigned short &deflation)
  {
    if (!co_) {
      co_ = true;
      nu_ = deflation;
    }
  }
  /* Preconditioner::setVectors (preconditioner.hpp:358-362): takes ownership of ev and *ev */
  template <class T>
  void setVectors(T const &ev)
  {
    ev_ = ev;
  }
  const K *const *getVectors() const { return ev_; }
  unsigned short  getLocal() const { return co_ ? nu_ : 0; }
};
template <class K>
inline int mpi_allgather_cb(const void *send, void *recv, size_t bytes, void *user)
{
  return MPI_Allgather(const_cast<void *>(send), static_cast<int>(bytes), MPI_BYTE, recv, static_cast<int>(bytes), MPI_BYTE, *static_cast<MPI_Comm *>(user)) == MPI_SUCCESS ? 0 : -1;
}
} // namespace b200

template <template <class> class CoarseSolver, char S, class K>
class Schwarz<B200Sub, CoarseSolver, S, K> : public b200::Prcndtnr<K> {
  static_assert(std::is_same<K, double>::value || std::is_same<K, std::complex<double>>::value, "hpddm_b200: K must be double or std::complex<double>");
  typedef b200::Api<K> A_;

public:
  enum class Prcndtnr : char { NO, SY, GE, OS, OG }; /* include/HPDDM_schwarz.hpp:104-110; same values as HPDDM_B200_PRCNDTNR_* */
  typedef b200::Prcndtnr<K> super;
  typedef K                 scalar_type;

protected:
  const underlying_type<K>   *d_;
  Prcndtnr                    type_;
  mutable typename A_::ctx_t *ctx_;
  mutable typename A_::sub_t *sub_;
  mutable MPI_Comm            comm_; /* stable storage for the control-plane callback */
  mutable bool                vectors_sent_;

  /* first device-side need: context on this rank's GPU, control plane on Subdomain::communicator_, matrix + neighbour map upload
   * (Subdomain::initialize itself is the reference's code and knows nothing about the GPU) */
  void device() const
  {
    if (sub_) return;
    comm_ = Subdomain<K>::communicator_;
    int rank = 0, size = 1;
    MPI_Comm_rank(comm_, &rank);
    MPI_Comm_size(comm_, &size);
    if (!std::getenv("HPDDM_B200_DEVICE")) {
      int ndev = 1;
      hpddm_b200_device_count(&ndev);
      setenv("HPDDM_B200_DEVICE", std::to_string(rank % std::max(1, ndev)).c_str(), 0);
    }
    ctx_ = b200::context<K>(true);
    if (size > 1) {
      if (const char *e = std::getenv("HPDDM_B200_NCCL")) {
        if (std::atoi(e)) {
          char id[128];
          if (rank == 0) b200::check<K>(A_::nccl_unique_id(id), "nccl_unique_id");
          MPI_Bcast(id, 128, MPI_BYTE, 0, comm_);
          b200::check<K>(A_::ctx_comm_init(ctx_, id, rank, size), "ctx_comm_init");
        }
      }
      b200::check<K>(A_::ctx_comm_init_host(ctx_, rank, size, &b200::mpi_allgather_cb<K>, &comm_), "ctx_comm_init_host");
    }
    b200::check<K>(A_::sub_create(ctx_, rank, &sub_), "sub_create");
    const MatrixCSR<K> *a = Subdomain<K>::a_;
    b200::check<K>(A_::sub_set_matrix(sub_, a->n_, a->nnz_, a->ia_, a->ja_, a->a_, a->sym_ ? 1 : 0, a->ia_[0] ? 'F' : 'C'), "sub_set_matrix");
    std::vector<int> ranks, sizes, idx;
    for (const auto &m : Subdomain<K>::map_) {
      ranks.push_back(m.first);
      sizes.push_back(static_cast<int>(m.second.size()));
      idx.insert(idx.end(), m.second.begin(), m.second.end());
    }
    b200::check<K>(A_::sub_set_neighbors(sub_, static_cast<int>(ranks.size()), ranks.data(), sizes.data(), idx.data()), "sub_set_neighbors");
    if (const char *g = std::getenv("HPDDM_B200_GRID")) { /* optional "nx,ny,nz[,dof]" ordering hint */
      int nx = 0, ny = 0, nz = 1, dof = 1;
      if (std::sscanf(g, "%d,%d,%d,%d", &nx, &ny, &nz, &dof) >= 2 && static_cast<long long>(nx) * ny * nz * dof == a->n_) A_::sub_set_grid_hint(sub_, nx, ny, nz, dof);
    }
    if (d_) b200::check<K>(A_::sub_set_scaling(sub_, d_), "sub_set_scaling");
  }
  int correction() const
  {
    if (!super::co_) return HPDDM_B200_CORRECTION_NONE;
    /* read at every apply like the reference (schwarz.hpp:530): the driver may change it between solves */
    return static_cast<int>(Option::get()->val<char>(super::prefix("schwarz_coarse_correction"), -1));
  }

public:
  Schwarz() : d_(), type_(Prcndtnr::NO), ctx_(), sub_(), comm_(), vectors_sent_() { }
  Schwarz(const Schwarz &) = delete;
  ~Schwarz()
  {
    if (std::getenv("HPDDM_B200_DEBUG") && ctx_)
      std::fprintf(stderr, "[hpddm_b200] HPDDM::Schwarz<B200Sub>: %lld kernel launches, transport %d\n", static_cast<long long>(hpddm_launches()), A_::ctx_transport(ctx_));
    if (sub_) A_::sub_destroy(sub_);
    if (ctx_) A_::ctx_destroy(ctx_);
  }
  long long hpddm_launches() const { return ctx_ ? static_cast<long long>(b200::launch_count<K>(ctx_)) : 0; }
  /* Schwarz::initialize(d) (schwarz.hpp:178): d stays owned by the caller (examples/schwarz.cpp:185 deletes it) */
  void initialize(underlying_type<K> *const &d)
  {
    d_ = d;
    if (sub_) b200::check<K>(A_::sub_set_scaling(sub_, d_), "sub_set_scaling");
  }
  /* Schwarz::multiplicityScaling (schwarz.hpp:381-404) */
  void multiplicityScaling(underlying_type<K> *const d) const
  {
    device();
    underlying_type<K> *arr[1] = {d};
    b200::check<K>(A_::multiplicity_scaling(ctx_, arr), "multiplicity_scaling");
  }
  /* Schwarz::exchange<allocate> (schwarz.hpp:180-188): x <- sum_j R_j^T D_j x_j */
  template <bool allocate = false>
  void exchange(K *const x, const unsigned short &mu = 1) const
  {
    device();
    K *xx[1] = {x};
    b200::check<K>(A_::exchange(ctx_, xx, mu, 1, HPDDM_B200_HOST), "exchange");
  }
  const underlying_type<K> *getScaling() const { return d_; }
  /* Schwarz::callNumfact (schwarz.hpp:337-368): the method comes from -hpddm_schwarz_method exactly as in the reference */
  template <char N = HPDDM_NUMBERING>
  void callNumfact(MatrixCSR<K> *const &A = nullptr)
  {
    device();
    const unsigned short m = Option::get()->val<unsigned short>(super::prefix("schwarz_method"));
    switch (m) {
    case HPDDM_SCHWARZ_METHOD_SORAS:
      type_ = (A ? Prcndtnr::OS : Prcndtnr::SY);
      break;
    case HPDDM_SCHWARZ_METHOD_ASM:
      type_ = Prcndtnr::SY;
      break;
    case HPDDM_SCHWARZ_METHOD_NONE:
      type_ = Prcndtnr::NO;
      break;
    default:
      type_ = (A && (m == HPDDM_SCHWARZ_METHOD_ORAS || m == HPDDM_SCHWARZ_METHOD_OSM) ? Prcndtnr::OG : Prcndtnr::GE);
    }
    const int t = static_cast<int>(type_);
    if ((type_ == Prcndtnr::OS || type_ == Prcndtnr::OG) && A) b200::check<K>(A_::sub_numfact(sub_, t, A->n_, A->nnz_, A->ia_, A->ja_, A->a_, A->sym_ ? 1 : 0, N), "sub_numfact");
    else b200::check<K>(A_::sub_numfact(sub_, t, 0, 0, nullptr, nullptr, nullptr, 0, 'C'), "sub_numfact");
  }
  /* Schwarz::solveGEVP<Eps>(A_Neumann, B, pattern) (schwarz.hpp:665-715): nu from -hpddm_geneo_nu, optional -hpddm_geneo_threshold
   * (selectNu, include/HPDDM_eigensolver.hpp:69-159: keep the pairs below the threshold), vectors owned like the reference's ev_,
   * -hpddm_geneo_nu updated to the number kept.  The pencil is the reference's: A_Neu x = lambda (D A_Neu D restricted to the overlap) x;
   * a user matrix B is not supported (-> error).  Real scalars only (hpddm_b200z_sub_solve_gevp reports the error). */
  template <template <class> class Eps = B200Eps>
  void solveGEVP(MatrixCSR<K> *const &A, MatrixCSR<K> *const &B = nullptr, const MatrixCSR<K> *const & = nullptr)
  {
    device();
    if (B) throw std::runtime_error("hpddm_b200: solveGEVP with a user right-hand-side matrix is not supported");
    Option                  &opt       = *Option::get();
    const std::string        prefix    = super::prefix();
    const underlying_type<K> threshold = opt.val(prefix + "geneo_threshold", 0.0);
    unsigned short           nu        = opt.val<unsigned short>(prefix + "geneo_nu", 20);
    if (super::ev_) {
      delete[] *super::ev_;
      delete[] super::ev_;
      super::ev_ = nullptr;
    }
    vectors_sent_ = false;
    if (nu > 0) {
      std::vector<double> lambda(nu);
      b200::check<K>(A_::sub_solve_gevp(sub_, A->n_, A->nnz_, A->ia_, A->ja_, A->a_, A->sym_ ? 1 : 0, A->ia_[0] ? 'F' : 'C', nu, opt.val(prefix + "eigensolver_tol", 1.0e-6), 0, lambda.data()), "sub_solve_gevp");
      if (threshold > 0.0) {
        unsigned short keep = 0;
        while (keep < nu && lambda[keep] < threshold) ++keep;
        nu = std::max<unsigned short>(keep, 1);
      }
      const int n = Subdomain<K>::dof_;
      int       have = 0;
      b200::check<K>(A_::sub_get_vectors(sub_, nullptr, &have), "sub_get_vectors");
      std::vector<K> all(static_cast<std::size_t>(have) * n);
      b200::check<K>(A_::sub_get_vectors(sub_, all.data(), &have), "sub_get_vectors");
      super::ev_  = new K *[nu];
      *super::ev_ = new K[static_cast<std::size_t>(nu) * n];
      for (unsigned short i = 0; i < nu; ++i) super::ev_[i] = *super::ev_ + static_cast<std::size_t>(i) * n;
      std::copy_n(all.begin(), static_cast<std::size_t>(nu) * n, *super::ev_);
      if (nu != have) b200::check<K>(A_::sub_set_vectors(sub_, *super::ev_, nu), "sub_set_vectors");
      vectors_sent_ = true;
    } else b200::check<K>(A_::sub_set_vectors(sub_, nullptr, 0), "sub_set_vectors");
    opt[prefix + "geneo_nu"] = nu;
    if (super::co_) super::nu_ = nu;
  }
  /* Schwarz::buildTwo<excluded>(comm, A) (schwarz.hpp:440-495 -> Preconditioner::buildTwo -> CoarseOperator::construction with the
   * Galerkin blocks of include/HPDDM_operator.hpp:395-528): E = Z^H A Z assembled on the GPUs, replicated, factored.  The reference's
   * coarse-solver distribution options (-hpddm_level_2_p, _distribution, _aggregate_size, ...) have no meaning for a replicated solve. */
  template <unsigned short excluded = 0>
  int buildTwo(const MPI_Comm &, MatrixCSR<K> *const & = nullptr)
  {
    static_assert(excluded == 0, "hpddm_b200: every rank owns a subdomain (excluded = 0)");
    device();
    if (!vectors_sent_) {
      b200::check<K>(A_::sub_set_vectors(sub_, super::ev_ ? *super::ev_ : nullptr, super::ev_ ? super::nu_ : 0), "sub_set_vectors");
      vectors_sent_ = true;
    }
    b200::check<K>(A_::build_coarse(ctx_), "build_coarse");
    return 0;
  }
  /* Schwarz::start<excluded>(b, x, mu) (schwarz.hpp:496-514); returns what Subdomain::setBuffer would: nothing to free on the host */
  template <bool excluded = false>
  bool start(const K *const b, K *const x, const unsigned short &mu = 1) const
  {
    device();
    const K *bb[1] = {b};
    K       *xx[1] = {x};
    b200::check<K>(A_::start(ctx_, bb, xx, mu, HPDDM_B200_HOST), "start");
    return false;
  }
  /* Subdomain::end (subdomain.hpp:289) */
  void end(const bool = true) const
  {
    if (ctx_) A_::end(ctx_);
  }
  /* Schwarz::apply<excluded>(in, out, mu, work) (schwarz.hpp:527-612).  `in` is never clobbered. */
  template <bool excluded = false>
  int apply(const K *const in, K *const out, const unsigned short &mu = 1, K * = nullptr) const
  {
    device();
    const K *ii[1] = {in};
    K       *oo[1] = {out};
    return A_::apply(ctx_, ii, oo, mu, correction(), HPDDM_B200_HOST);
  }
  /* Schwarz::deflation<excluded, transpose> (schwarz.hpp:1602-1622); E is applied as is for transpose = false only */
  template <bool excluded, bool transpose = false>
  void deflation(const K *const in, K *const out, const unsigned short &mu) const
  {
    static_assert(!transpose, "hpddm_b200: transposed deflation is not implemented");
    device();
    const K *ii[1] = {in};
    K       *oo[1] = {out};
    b200::check<K>(A_::deflation(ctx_, ii, oo, mu, HPDDM_B200_HOST), "deflation");
  }
  /* Schwarz::GMV (schwarz.hpp:726-747) */
  int GMV(const K *const in, K *const out, const int &mu = 1, MatrixCSR<K> *const & = nullptr) const
  {
    device();
    const K *ii[1] = {in};
    K       *oo[1] = {out};
    return A_::gmv(ctx_, ii, oo, mu, HPDDM_B200_HOST);
  }
  /* Schwarz::computeResidual (schwarz.hpp:761-803): boundary rows masked, penalised entries of f rescaled, l2 / l1 / l-infinity */
  void computeResidual(const K *const x, const K *const f, underlying_type<K> *const storage, const unsigned short mu = 1, const unsigned short norm = HPDDM_COMPUTE_RESIDUAL_L2) const
  {
    device();
    const K *xx[1] = {x}, *ff[1] = {f};
    b200::check<K>(A_::compute_residual(ctx_, xx, ff, storage, mu, norm, HPDDM_B200_HOST), "compute_residual");
  }
  /* Subdomain::boundaryConditions (subdomain.hpp:327-336), from the device-side scan of the same matrix */
  std::unordered_map<unsigned int, K> boundaryConditions() const
  {
    device();
    int cnt = 0;
    b200::check<K>(A_::sub_boundary_conditions(sub_, nullptr, nullptr, &cnt), "sub_boundary_conditions");
    std::vector<int> idx(cnt);
    std::vector<K>   val(cnt);
    if (cnt) b200::check<K>(A_::sub_boundary_conditions(sub_, idx.data(), val.data(), &cnt), "sub_boundary_conditions");
    std::unordered_map<unsigned int, K> map;
    map.reserve(cnt);
    for (int i = 0; i < cnt; ++i) map[static_cast<unsigned int>(idx[i])] = val[i];
    return map;
  }
  /* OptionsPrefix::destroy (include/HPDDM_option.hpp:431-443): drops the recycled Krylov subspace -- the host one of the reference's
   * GCRO-DR drivers (A.storage()) and the device one of solveOnDevice (kept in the context) */
  template <bool reset = true>
  void destroy()
  {
    Subdomain<K>::template destroy<reset>();
    if (ctx_) A_::recycle_destroy(ctx_);
  }
  /* device-resident counterpart of IterativeMethod::solve: see B200Schwarz::solve (HPDDM_B200.hpp) */
  int solveOnDevice(const K *const f, K *const x, const unsigned short mu = 1) const
  {
    device();
    const Option     &opt    = *Option::get();
    const std::string prefix = super::prefix();
    const int         method = opt.val<char>(prefix + "krylov_method", HPDDM_KRYLOV_METHOD_GMRES), restart = opt.val<unsigned short>(prefix + "gmres_restart", 40), max_it = opt.val<unsigned short>(prefix + "max_it", 100);
    const double      tol = opt.val(prefix + "tol", 1.0e-6);
    const K          *bb[1] = {f};
    K                *xx[1] = {x};
    int               it = 0, rc;
    if (method == HPDDM_KRYLOV_METHOD_GCRODR || method == HPDDM_KRYLOV_METHOD_BGCRODR) {  // -hpddm_recycle / _recycle_target / _recycle_strategy / _recycle_same_system as IterativeMethod::options reads them (iterative.hpp:215-217)
      const int same = std::min(opt.val<unsigned short>(prefix + "recycle_same_system"), static_cast<unsigned short>(2));
      const int k = opt.val<int>(prefix + "recycle", 0), target = opt.val<char>(prefix + "recycle_target", HPDDM_RECYCLE_TARGET_SM), strategy = opt.val<char>(prefix + "recycle_strategy", HPDDM_RECYCLE_STRATEGY_A);
      rc = method == HPDDM_KRYLOV_METHOD_GCRODR ? A_::solve_gcrodr(ctx_, bb, xx, mu, correction(), restart, k, target, strategy, same, max_it, tol, HPDDM_B200_HOST, &it, nullptr)
                                                 : A_::solve_bgcrodr(ctx_, bb, xx, mu, correction(), restart, k, target, strategy, same, max_it, tol, HPDDM_B200_HOST, &it, nullptr);
      if (rc == 0 && it != 0 && it != max_it && same) (*Option::get())[prefix + "recycle_same_system"] += 1;  // GCRODR.hpp:435
    }
    else if (method == HPDDM_KRYLOV_METHOD_BGMRES) rc = A_::solve_bgmres(ctx_, bb, xx, mu, correction(), restart, max_it, tol, HPDDM_B200_HOST, &it, nullptr);
    else if (method == HPDDM_KRYLOV_METHOD_CG) rc = A_::solve_cg(ctx_, bb, xx, mu, correction(), max_it, tol, HPDDM_B200_HOST, &it, nullptr);
    else rc = A_::solve(ctx_, bb, xx, mu, correction(), restart, max_it, tol, HPDDM_B200_HOST, &it, nullptr);
    return rc < 0 ? rc : it;
  }
  typename A_::ctx_t *context() const { return ctx_; }
  typename A_::sub_t *handle() const { return sub_; }
};
} // namespace HPDDM
#endif // HPDDM_B200_SCHWARZ_HPP_
