/*
 * hpddm_b200z.h -- complex double (K = std::complex<double>) instantiation of the C ABI of
 * libhpddm_b200.so.  Same entry points, same semantics and same reference citations as
 * include/hpddm_b200.h (read that header for the conventions); only the scalar type changes:
 * the reference is templated on K (HPDDM::Schwarz<Solver, CoarseSolver, S, K>,
 * include/HPDDM_schwarz.hpp:52; K = std::complex<double> under FORCE_COMPLEX,
 * examples/schwarz.hpp:48-62), a C ABI needs one symbol set per K.  This is BASELINE.json
 * config 5 (Helmholtz, complex FP64, ORAS).
 *
 *  - hpddm_b200_z is layout-compatible with std::complex<double> / C99 double _Complex
 *    (interleaved re, im); device pointers must be 16-byte aligned.
 *  - quantities of type underlying_type<K> in the reference stay real: the partition of unity d
 *    (Schwarz::initialize(underlying_type<K>*), include/HPDDM_schwarz.hpp:178), tolerances,
 *    residual norms.
 *  - transposes are conjugate where the reference uses Wrapper<K>::transc: the coarse
 *    restriction Z^H D x (schwarz.hpp:1616), E = Z^H A Z (include/HPDDM_operator.hpp:395-528),
 *    and the inner products sum_i d_i conj(x_i) y_i (include/HPDDM_iterative.hpp:503).
 *  - local factorisation: LU without pivoting (no LL^T for complex symmetric matrices);
 *    hpddm_b200z_sub_solve_gevp is not provided for complex scalars (config 5 supplies its own
 *    coarse vectors through hpddm_b200z_sub_set_vectors) and returns HPDDM_B200_ERR_STATE.
 *  - contexts / subdomains of the two scalar types are distinct opaque types.
 */
#ifndef HPDDM_B200Z_H
#define HPDDM_B200Z_H

#include "hpddm_b200.h" /* constants, error codes, hpddm_b200_stats */

#ifndef HPDDM_B200_Z_T
  #ifdef __cplusplus
    #include <complex>
    #define HPDDM_B200_Z_T std::complex<double>
  #else
    #include <complex.h>
    #define HPDDM_B200_Z_T double _Complex
  #endif
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef HPDDM_B200_Z_T hpddm_b200_z;
typedef struct hpddm_b200z_ctx hpddm_b200z_ctx;
typedef struct hpddm_b200z_sub hpddm_b200z_sub;

const char *hpddm_b200z_last_error(void);
const char *hpddm_b200z_version(void);

/* ---- context ------------------------------------------------------------- */
/* One per process (or per GPU).  Replaces MPI_Init + Subdomain::communicator_
 * ownership (include/HPDDM_subdomain.hpp:49-63). */
/* number of CUDA devices visible to this process (0 when there is none: nothing in this library runs without one) */
int hpddm_b200z_device_count(int *count);
int hpddm_b200z_ctx_create(int device, hpddm_b200z_ctx **ctx);
int hpddm_b200z_ctx_destroy(hpddm_b200z_ctx *ctx);
/* NCCL bootstrap (replaces the MPI communicator on the hot path only, SURVEY.md
 * section 5): rank 0 of the job calls unique_id, broadcasts the 128 bytes with
 * whatever it has (MPI_Bcast, torch.distributed, a file), every process then
 * calls comm_init.  Not calling comm_init = single-process decomposition. */
int hpddm_b200z_nccl_unique_id(void *id128);
int hpddm_b200z_ctx_comm_init(hpddm_b200z_ctx *ctx, const void *id128, int proc_rank, int nproc);
/* Alternative / complement to the NCCL bootstrap: the host program lends its own communicator for the CONTROL plane (exchange of
 * CUDA-IPC handles, coarse-space sizes, the rows of E at setup) as a blocking all-gather between host buffers -- in an MPI
 * program `MPI_Allgather(send, bytes, MPI_BYTE, recv, bytes, MPI_BYTE, comm)` on Subdomain::communicator_
 * (include/HPDDM_subdomain.hpp:49-63).  The DATA plane of the hot path (halo sums, coarse gather, Krylov reductions) then runs
 * over the library's peer-memory fabric (CUDA IPC + NVLink stores, hb_p2p.cu); NCCL is only needed when a peer is not IPC-reachable
 * or a process hosts several subdomains.  Several processes may share one GPU.  Returns 0 / a negative error code. */
typedef int (*hpddm_b200z_allgather_fn)(const void *send, void *recv, size_t bytes_per_rank, void *user);
int hpddm_b200z_ctx_comm_init_host(hpddm_b200z_ctx *ctx, int proc_rank, int nproc, hpddm_b200z_allgather_fn allgather, void *user);
/* which transport carries the hot-path collectives: 0 = single process, 1 = NCCL, 2 = peer-memory fabric (decided collectively the
 * first time work space is sized, i.e. after the first hot-path or setup call) */
int hpddm_b200z_ctx_transport(hpddm_b200z_ctx *ctx);
int hpddm_b200z_ctx_synchronize(hpddm_b200z_ctx *ctx);
/* raw cudaStream_t of the context (for callers that time with CUDA events) */
void *hpddm_b200z_ctx_stream(hpddm_b200z_ctx *ctx);
/* number of kernels this library launched on the context since creation */
int64_t hpddm_b200z_ctx_launch_count(hpddm_b200z_ctx *ctx);
/* number of caller host ranges pinned in place (cudaHostRegister) since creation: HOST pointers are read / written where they are;
 * a range passed a fourth time between start() and end() is registered lazily (the Krylov arena of an unchanged driver,
 * include/HPDDM_GMRES.hpp:45-50), all registrations are dropped by end().  HPDDM_B200_HOSTREG=0 disables this. */
int64_t hpddm_b200z_ctx_hostreg_count(hpddm_b200z_ctx *ctx);
/* device allocation helpers for callers without a CUDA runtime binding */
int hpddm_b200z_malloc(hpddm_b200z_ctx *ctx, size_t bytes, void **dptr);
int hpddm_b200z_free(hpddm_b200z_ctx *ctx, void *dptr);
int hpddm_b200z_memcpy(hpddm_b200z_ctx *ctx, void *dst, const void *src, size_t bytes, int dst_where, int src_where);

/* ---- subdomain setup ------------------------------------------------------ */
/* HPDDM::Schwarz<...> A;  (examples/schwarz.cpp:90) */
int hpddm_b200z_sub_create(hpddm_b200z_ctx *ctx, int global_rank, hpddm_b200z_sub **sub);
int hpddm_b200z_sub_destroy(hpddm_b200z_sub *sub);
/* Subdomain::initialize(MatrixCSR*, ...) matrix part (include/HPDDM_subdomain.hpp:165-183).
 * MatrixCSR layout (include/HPDDM_matrix.hpp:33-57,156-165): ia[n+1], ja[nnz], a[nnz];
 * sym != 0 => lower triangle only; numbering 'C' or 'F'.  The arrays are copied. */
int hpddm_b200z_sub_set_matrix(hpddm_b200z_sub *sub, int n, int nnz, const int *ia, const int *ja, const hpddm_b200_z *a, int sym, char numbering);
/* Subdomain::initialize neighbour part, C-array overload (subdomain.hpp:238-259):
 * `count` neighbours, global ranks `ranks[i]`, `sizes[i]` shared dofs each,
 * indices concatenated in `idx` (C numbering).  Sorted by rank inside, empty
 * lists dropped, as the reference does. */
int hpddm_b200z_sub_set_neighbors(hpddm_b200z_sub *sub, int count, const int *ranks, const int *sizes, const int *idx);
/* Schwarz::initialize(d) (include/HPDDM_schwarz.hpp:178): the partition of unity
 * (after multiplicityScaling).  Copied to the device. */
int hpddm_b200z_sub_set_scaling(hpddm_b200z_sub *sub, const double *d);
/* Optional ordering hint: the dofs are a lexicographic nx x ny x nz grid with
 * `dof` unknowns per node (x fastest) -> geometric nested dissection.  Without
 * it an algebraic (BFS level-set) nested dissection is used.  New; nothing in
 * the reference corresponds (orderings are chosen inside MUMPS/CHOLMOD). */
int hpddm_b200z_sub_set_grid_hint(hpddm_b200z_sub *sub, int nx, int ny, int nz, int dof);
/* Schwarz::multiplicityScaling (include/HPDDM_schwarz.hpp:381-404), collective:
 * d[s] in/out (host), one array per local subdomain. */
int hpddm_b200z_multiplicity_scaling(hpddm_b200z_ctx *ctx, double *const *d);
/* Schwarz::callNumfact -> SUBDOMAIN::numfact (schwarz.hpp:337-368; e.g.
 * include/HPDDM_SuiteSparse.hpp:264-371): GPU multifrontal LL^T / LU of the
 * local matrix (or of the ORAS/SORAS matrix passed as ia/ja/a when
 * prcndtnr = OG / OS; pass NULLs to factor the matrix given to set_matrix).
 * Leaves the supernodal panels in HBM in the layout the SpTRSV consumes. */
int hpddm_b200z_sub_numfact(hpddm_b200z_sub *sub, int prcndtnr, int n, int nnz, const int *ia, const int *ja, const hpddm_b200_z *a, int sym, char numbering);
/* Preconditioner::setVectors (include/HPDDM_preconditioner.hpp:358-362): Z =
 * ev_[0], one contiguous column-major n x nu block.  Copied to the device. */
int hpddm_b200z_sub_set_vectors(hpddm_b200z_sub *sub, const hpddm_b200_z *Z, int nu);
/* Schwarz::solveGEVP<EIGENSOLVER>(A_Neumann) (include/HPDDM_schwarz.hpp:665-715 + scaleIntoOverlap 622-657; the reference's
 * EIGENSOLVER is ARPACK in shift-invert mode, include/HPDDM_ARPACK.hpp:84-178): the nu smallest eigenpairs of
 * A_Neu x = lambda (D A_Neu D restricted to the overlap) x, computed on the GPU (block subspace iteration on A_Neu^-1 B with the
 * block SpTRSV) and installed as the deflation vectors.  tol <= 0 -> 1e-6 (eigensolver_tol), max_it <= 0 -> 100.
 * eigenvalues[nu] optional.  Returns the number of iterations (> 0) or a negative error code.  Needs set_neighbors + set_scaling. */
int hpddm_b200z_sub_solve_gevp(hpddm_b200z_sub *sub, int n, int nnz, const int *ia, const int *ja, const hpddm_b200_z *a, int sym, char numbering, int nu, double tol,
                               int max_it, double *eigenvalues);
/* Preconditioner::getVectors (include/HPDDM_preconditioner.hpp): copy Z (n x nu, column-major) back to the host; Z may be NULL to query nu */
int hpddm_b200z_sub_get_vectors(hpddm_b200z_sub *sub, hpddm_b200_z *Z, int *nu);
/* Schwarz::buildTwo (include/HPDDM_schwarz.hpp:440-495 -> preconditioner.hpp:124-257
 * -> CoarseOperator::construction, coarse_operator_impl.hpp:220-272; Galerkin
 * blocks of include/HPDDM_operator.hpp:395-528), collective: assembles
 * E = Z^T A Z on the GPUs, replicates and factors it. */
int hpddm_b200z_build_coarse(hpddm_b200z_ctx *ctx);
/* Alternative to build_coarse: install a user-assembled dense coarse operator
 * (N_c x N_c column-major, N_c = sum of all nu).  UserCoarseOperator analogue,
 * include/HPDDM_operator.hpp:351-375. */
int hpddm_b200z_set_coarse(hpddm_b200z_ctx *ctx, const hpddm_b200_z *E, int Nc);
/* copy out the assembled coarse operator (host, N_c x N_c column-major);
 * -hpddm_level_2_dump_matrix analogue (coarse_operator_impl.hpp:1031-1061) */
int hpddm_b200z_get_coarse(hpddm_b200z_ctx *ctx, hpddm_b200_z *E, int *Nc);

/* ---- the hot path (collective over the context's subdomains) --------------- */
/* Schwarz::start<false>(b, x, mu) (include/HPDDM_schwarz.hpp:496-514): impose
 * penalised rows on x, exchange(x), size the coarse work space for mu columns. */
int hpddm_b200z_start(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *b, hpddm_b200_z *const *x, int mu, int where);
/* Subdomain::end (include/HPDDM_subdomain.hpp:289) */
int hpddm_b200z_end(hpddm_b200z_ctx *ctx);
/* Schwarz::apply<false>(in, out, mu, work) (include/HPDDM_schwarz.hpp:527-612).
 * `in` is never modified (a private work space is used). */
int hpddm_b200z_apply(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *in, hpddm_b200_z *const *out, int mu, int correction, int where);
/* Schwarz::deflation<false>(in, out, mu) (include/HPDDM_schwarz.hpp:1602-1622) */
int hpddm_b200z_deflation(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *in, hpddm_b200_z *const *out, int mu, int where);
/* Schwarz::exchange(x, mu) when scaled != 0 (schwarz.hpp:180-188), else
 * Subdomain::exchange(x, mu) (include/HPDDM_subdomain.hpp:115-130); in place. */
int hpddm_b200z_exchange(hpddm_b200z_ctx *ctx, hpddm_b200_z *const *x, int mu, int scaled, int where);
/* Schwarz::GMV(in, out, mu) (include/HPDDM_schwarz.hpp:726-747) */
int hpddm_b200z_gmv(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *in, hpddm_b200_z *const *out, int mu, int where);
/* SUBDOMAIN::solve(b, x, n) (e.g. include/HPDDM_SuiteSparse.hpp:388-423):
 * local triangular solves only, no communication. */
int hpddm_b200z_sub_solve(hpddm_b200z_sub *sub, const hpddm_b200_z *b, hpddm_b200_z *x, int mu, int where);
/* CoarseOperator::callSolver(rhs, mu) (include/HPDDM_coarse_operator_impl.hpp:1630-1732):
 * rhs[s] = nu_s x mu (column-major, ld = nu_s) in, solution out. */
int hpddm_b200z_coarse_solve(hpddm_b200z_ctx *ctx, hpddm_b200_z *const *rhs, int mu, int where);
/* D-weighted dot products per column, summed over the context's subdomains and
 * (when a communicator exists) all processes: sum_i d_i x_i y_i -- the
 * reduction the Krylov layer performs (include/HPDDM_GMRES.hpp:59-68,
 * include/HPDDM_iterative.hpp:455-468).  result[mu] on the host. */
int hpddm_b200z_dot(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *x, const hpddm_b200_z *const *y, int mu, hpddm_b200_z *result, int where);
/* Subdomain::boundaryConditions (include/HPDDM_subdomain.hpp:310-336): the rows of the matrix given to set_matrix that impose a
 * (penalised or identity-row) Dirichlet condition, with their diagonal values.  idx / val may be NULL to query *count. */
int hpddm_b200z_sub_boundary_conditions(hpddm_b200z_sub *sub, int *idx, hpddm_b200_z *val, int *count);
/* ||b|| as IterativeMethod::initializeNorm computes it for right preconditioning (include/HPDDM_iterative.hpp:455-468): D-weighted l2
 * norm per column, entries larger than HPDDM_PEN * HPDDM_EPS on boundary-condition rows divided by HPDDM_PEN first.  norm[mu], host.
 * The device-resident Krylov drivers use the same reduction. */
int hpddm_b200z_rhs_norm(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *b, int mu, double *norm, int where);
/* Schwarz::computeResidual(x, f, storage, mu, norm) (include/HPDDM_schwarz.hpp:761-803): storage[2 nu] = ||f||, storage[2 nu + 1] =
 * ||A x - f|| off the boundary-condition rows, per column; norm = 0 (l2, HPDDM_COMPUTE_RESIDUAL_L2), 1 (l1) or 2 (l-infinity);
 * D-weighted, reduced over all subdomains and processes; storage on the host. */
int hpddm_b200z_compute_residual(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *x, const hpddm_b200_z *const *f, double *storage, int mu, int norm, int where);


/* ---- device-resident Krylov driver ("next" row of SURVEY.md section 8f) ------------------------------ */
/* IterativeMethod::solve -> GMRES (include/HPDDM_iterative.hpp:1013-1111, include/HPDDM_GMRES.hpp:31-158) with the
 * reference defaults: right preconditioning, classical Gram-Schmidt, D-weighted inner products, convergence when
 * |s_i| / ||b||_D <= tol.  The Krylov basis stays in HBM; x holds the initial guess on entry, the solution on exit;
 * rel_residual[mu] (optional) receives the last preconditioned residual estimates.  Returns the iteration count in
 * *iterations (the value IterativeMethod::solve returns). */
int hpddm_b200z_solve(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *b, hpddm_b200_z *const *x, int mu, int correction, int restart, int max_it, double tol, int where,
                      int *iterations, double *rel_residual);

/* IterativeMethod::CG (include/HPDDM_CG.hpp:31-168); see hpddm_b200_solve_cg */
int hpddm_b200z_solve_cg(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *b, hpddm_b200_z *const *x, int mu, int correction, int max_it, double tol, int where,
                         int *iterations, double *rel_residual);

/* IterativeMethod::BGMRES (include/HPDDM_GMRES.hpp:160-313); see hpddm_b200_solve_bgmres */
int hpddm_b200z_solve_bgmres(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *b, hpddm_b200_z *const *x, int mu, int correction, int restart, int max_it, double tol,
                             int where, int *iterations, double *rel_residual);

/* IterativeMethod::GCRODR (include/HPDDM_GCRODR.hpp:35-444); see hpddm_b200_solve_gcrodr (targets / strategies: HPDDM_B200_RECYCLE_*) */
int hpddm_b200z_solve_gcrodr(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *b, hpddm_b200_z *const *x, int mu, int correction, int restart, int recycle,
                             int recycle_target, int recycle_strategy, int recycle_same_system, int max_it, double tol, int where, int *iterations, double *rel_residual);
/* IterativeMethod::BGCRODR (include/HPDDM_GCRODR.hpp:445-907); see hpddm_b200_solve_bgcrodr */
int hpddm_b200z_solve_bgcrodr(hpddm_b200z_ctx *ctx, const hpddm_b200_z *const *b, hpddm_b200_z *const *x, int mu, int correction, int restart, int recycle,
                              int recycle_target, int recycle_strategy, int recycle_same_system, int max_it, double tol, int where, int *iterations, double *rel_residual);
int hpddm_b200z_recycle_dim(hpddm_b200z_ctx *ctx);
int hpddm_b200z_recycle_destroy(hpddm_b200z_ctx *ctx);

/* ---- host-only planning entry points (no GPU needed): the N > 1 logic the hot path uses, exposed so that CPU-only multi-process
 * tests exercise the product's own code.  coarse_layout: offsets[nproc + 1] of the process blocks of the coarse vector and the padded
 * block length of the communication layout.  halo_schedule: the ordered NCCL sends / receives of one halo round for the `nlocal`
 * subdomains of a process (global ranks granks[], nb_count[q] neighbours each, their global ranks concatenated in nb_ranks[]):
 * quadruples (destination subdomain, source subdomain, local subdomain index, neighbour slot); sends / recvs may be NULL to query *nmsg. */
int hpddm_b200z_debug_coarse_layout(int nproc, const int *rows_per_proc, int *offsets, int *lmax);
int hpddm_b200z_debug_halo_schedule(int nlocal, const int *granks, const int *nb_count, const int *nb_ranks, int *sends, int *recvs, int *nmsg);

/* ---- introspection (Subdomain::statistics analogue, subdomain.hpp:405-454); factor_bytes counts 16-byte scalars */
int hpddm_b200z_sub_stats(hpddm_b200z_sub *sub, hpddm_b200_stats *st);

#ifdef __cplusplus
}
#endif
#endif /* HPDDM_B200Z_H */
