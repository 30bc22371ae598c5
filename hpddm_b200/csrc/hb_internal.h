// Internal data structures of libhpddm_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "hb_scalar.h"
#include "../../include/hpddm_b200.h"
#ifdef HB_COMPLEX
#define HPDDM_B200_Z_T hbz::cplx
#include "../../include/hpddm_b200z.h"
#endif

#ifdef HB_COMPLEX
typedef hpddm_b200z_ctx hb_ctx_t;
typedef hpddm_b200z_sub hb_sub_t;
#define HB_SCALAR_NAME "complex double"
#else
typedef hpddm_b200_ctx hb_ctx_t;
typedef hpddm_b200_sub hb_sub_t;
#define HB_SCALAR_NAME "double"
#endif

namespace hb {

// ---------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
#define HB_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      hb::set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), __FILE__, __LINE__, #call); \
      return HPDDM_B200_ERR_CUDA;                                                             \
    }                                                                                         \
  } while (0)
#define HB_CHECK(call)       \
  do {                       \
    int r__ = (call);        \
    if (r__ < 0) return r__; \
  } while (0)

// ---------------------------------------------------------------- panel layout
// One *panel* per front and per triangular factor: a row-major (s1+s2) x s1
// matrix  [ W ; M ]  with W = (diagonal block)^{-1} (lower triangular) and
// M = (sub-diagonal block) * W.  Pivot rows are stored in blocks of RB rows as
// a trapezoid: block k (rows [RB*k, RB*k+RB)) has row stride
// w_k = min(ldp, RB*(k+1)); update rows have stride ldp = roundup(s1, 4).
constexpr int RB = 32;
__host__ __device__ inline int hb_ldp(int s1) { return (s1 + 3) & ~3; }
__host__ __device__ inline int hb_wblk(int s1, int k) {
  int w = RB * (k + 1), l = hb_ldp(s1);
  return w < l ? w : l;
}
// element offset of pivot row block k inside a panel
__host__ __device__ inline int64_t hb_blk_off(int k) { return (int64_t)(RB * RB / 2) * k * (k + 1); }
// element offset of the first update row
__host__ __device__ inline int64_t hb_upd_off(int s1) {
  int nb = (s1 + RB - 1) / RB;
  if (nb == 0) return 0;
  int rows_last = s1 - RB * (nb - 1);
  return hb_blk_off(nb - 1) + (int64_t)rows_last * hb_wblk(s1, nb - 1);
}
__host__ __device__ inline int64_t hb_panel_size(int s1, int s2) { return hb_upd_off(s1) + (int64_t)s2 * hb_ldp(s1); }

// ---------------------------------------------------------------- symbolic
struct Front {
  int p0;        // first pivot (permuted index)
  int s1;        // pivots
  int s2;        // update rows (|struct|)
  int parent;    // parent front or -1
  int level;     // 0 = deepest; children are exactly one level below their parent
  int64_t rptr;  // offset of struct rows in Symbolic::rowidx / rel
  int64_t poff;  // element offset of this front's panel in the panel store
};

// forward-sweep work item (one warp): RB rows x up to FCH columns of one panel
struct FwdItem {
  int front;
  int rblk;  // row block index: < nb1 -> pivot block, else update block (rblk - nb1)
  int c0;    // first column
  int cw;    // columns of this item (<= FCH; smaller on levels that would otherwise under-fill the GPU)
};
// backward-sweep work item (one warp): BCH columns x up to BROWS rows
struct BwdItem {
  int front;
  int c0;  // first column of the chunk
  int r0;  // first row (panel row index in [0, s1+s2))
  int nr;  // rows
};
constexpr int FCH = 256 * VE;  // forward column chunk (4 KB of right-hand side per warp): 512 real / 256 complex
constexpr int BCH = 128 * VE;  // backward column chunk (4 x 128-bit accumulators per lane): 256 real / 128 complex
constexpr int BROWS = 128;  // backward rows per item

struct Symbolic {
  int n = 0;
  std::vector<int> perm, iperm;  // perm[new] = old
  std::vector<Front> fronts;     // in elimination (post)order
  std::vector<int> rowidx;       // struct rows, permuted indices, sorted per front
  std::vector<int> rel;          // position of each struct row in its parent's [pivots; struct]
  std::vector<int> front_of;     // permuted index -> front
  int nlevels = 0;
  std::vector<int> level_ptr;    // fronts sorted by level: level_order[level_ptr[l]..level_ptr[l+1])
  std::vector<int> level_order;
  std::vector<std::vector<int>> children;
  int64_t panel_elems = 0;       // total panel store elements (one triangular factor)
  int64_t nnz_factor = 0;        // structurally non-zero factor entries
  std::vector<FwdItem> fwd;      // sorted by level
  std::vector<int64_t> fwd_ptr;  // per level
  std::vector<BwdItem> bwd;
  std::vector<int64_t> bwd_ptr;
};

struct HostCSR {
  int n = 0;
  std::vector<int> ia, ja;  // full pattern, C numbering
  std::vector<K> a;
  bool symmetric = false;   // numerically symmetric (real scalars only: LL^T candidate)
};

// builds ordering + supernodal structure.  grid hint: nx*ny*nz*dof == n or nx==0
int symbolic_analyze(const HostCSR &A, int nx, int ny, int nz, int dof, int leaf, Symbolic &S);

// ---------------------------------------------------------------- device side
struct DeviceFactor {
  bool valid = false;
  bool symmetric = true;
  K *panL = nullptr;  // forward panels
  K *panU = nullptr;  // backward panels (== panL when symmetric)
  Front *fronts = nullptr;
  int *rowidx = nullptr;
  FwdItem *fwd = nullptr;
  BwdItem *bwd = nullptr;
  int *perm = nullptr;   // perm[new] = old
  K *b = nullptr, *y = nullptr, *x = nullptr;  // permuted work vectors (n * 8 each: up to 8 RHS per pass)
  // dependency-driven persistent sweeps (hb_solve.cu): per-front child counts / item counts, the backward items in launch order
  // (root level first), and the counters a solve works on
  int *fwd_children = nullptr, *fwd_total = nullptr, *bwd_total = nullptr;
  BwdItem *bwd_ordered = nullptr;
  int *sync_pending = nullptr, *sync_done = nullptr;
  unsigned long long *sync_next = nullptr;
  int *sync_err_host = nullptr, *sync_err = nullptr;  // mapped pinned word set by a warp that gave up waiting
  int nfronts = 0;
  cudaGraphExec_t graph[9] = {};  // captured sweep launches, indexed by the number of right-hand sides of the pass (1 .. 8)
  int sweep_launches = 0;                                  // kernels inside one captured graph
};

struct Ctx;
struct P2P;

struct Sub {
  Ctx *ctx = nullptr;
  int grank = -1;
  int n = 0;
  HostCSR A;                // matrix used by GMV / the coarse correction
  int gx = 0, gy = 0, gz = 0, gdof = 1;
  // neighbours
  std::vector<int> nb_rank;
  std::vector<int> nb_ptr;  // size nb+1 into nb_idx
  std::vector<int> nb_idx;
  std::vector<double> d_host;
  // device
  int *d_ia = nullptr, *d_ja = nullptr;
  K *d_a = nullptr;
  double *d_d = nullptr;    // partition of unity: real (underlying_type<K>)
  int *d_map = nullptr;     // concatenated neighbour indices (h)
  int *d_ebase = nullptr, *d_esize = nullptr;  // per entry: start / size of its neighbour segment
  std::vector<int> peer_seg;  // for a local neighbour: its segment index that points back to us
  int h = 0;
  K *d_send = nullptr, *d_recv = nullptr;  // h * mu_cap each
  int *d_uidx = nullptr, *d_useg = nullptr, *d_upos = nullptr;  // deterministic unpack (CSR by unique target)
  int nuniq = 0;
  std::vector<std::pair<int, K>> bc;  // penalised rows
  int *d_bc_idx = nullptr;
  K *d_bc_val = nullptr;
  unsigned char *d_bcflag = nullptr;  // n flags: 1 on boundary-condition rows (nullptr when there is none)
  // factor
  Symbolic sym;
  DeviceFactor fac;
  int prcndtnr = HPDDM_B200_PRCNDTNR_GE;
  double t_symbolic = 0, t_numfact = 0;
  // deflation
  int nu = 0;
  K *d_Z = nullptr;
  int coff = 0;             // offset of this subdomain in the coarse vector
  // work vectors (n * mu_cap)
  K *d_in = nullptr, *d_out = nullptr, *d_work = nullptr, *d_tmp = nullptr, *d_tmp2 = nullptr;
  int mu_cap = 0;
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;                 // second stream: the coarse correction of the additive variant runs beside the local solves
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<Sub *> subs;
  int64_t launches = 0;
  // communicator
  void *nccl = nullptr;  // ncclComm_t
  int proc_rank = 0, nproc = 1;
  // control-plane all-gather supplied by the host program (hpddm_b200_ctx_comm_init_host; e.g. MPI_Allgather on the
  // communicator of Subdomain::communicator_): host buffers, blocking.  nullptr -> NCCL carries the control plane too.
  int (*host_allgather)(const void *, void *, size_t, void *) = nullptr;
  void *host_allgather_user = nullptr;
  // coarse
  int Nc = 0;
  std::vector<int> coarse_off;  // per global rank, size P+1
  std::vector<int> nu_all;
  K *d_E = nullptr, *d_Einv = nullptr;  // Nc x Nc column-major
  K *d_T = nullptr, *d_Y = nullptr;     // Nc x mu_cap, layout [proc][col][row-in-proc]
  K *d_R = nullptr;                     // Nc x mu_cap residual of the coarse refinement step
  bool coarse_multipass = false;        // large N_c: the replicated solve runs as three grid-wide passes instead of one CTA
  int Lnu = 0;                               // Lmax: coarse rows per process block in the (padded) communication layout
  std::vector<int> Lnu_p;                    // actual coarse rows of every process
  int *d_rowproc = nullptr, *d_rowloc = nullptr;  // coarse row -> (process, row inside its block)
  int loc_off = 0;                           // first coarse row of this process
  K *d_res = nullptr;                   // small device scratch (dots)
  std::vector<K> E_host;
  int mu_cap = 0;
  bool started = false;
  // whole-apply CUDA graphs (single-process contexts): one per (mu, correction), rebuilt when any setter bumps `epoch`
  struct ApplyGraph {
    cudaGraphExec_t exec = nullptr;
    int64_t epoch = -1, launches = 0;
    int warm = 0;  // eager calls seen for this key (the first one also builds the nested sweep graphs)
  };
  std::map<std::pair<int, int>, ApplyGraph> apply_graphs;
  int64_t epoch = 0;
  P2P *p2p = nullptr;  // NVLink peer-memory halo state (hb_p2p.cu)
  // caller host memory pinned lazily (cudaHostRegister) between start() and end(): page-aligned, disjoint, sorted ranges;
  // host_seen counts how often a host pointer was passed in this bracket (a range is registered on its second sighting)
  struct HostRange {
    uintptr_t a, b;
  };
  std::vector<HostRange> hostreg;
  std::unordered_map<uintptr_t, int> host_seen;
  int64_t hostreg_calls = 0;  // successful cudaHostRegister calls (statistics / tests)
  void *recycled = nullptr;   // GCRO-DR: recycled pair (U, C) kept between solves (hb_krylov.cu, released by gcrodr_release)
};

// ---------------------------------------------------------------- kernels (launchers)
// K = scalar type of this build (hb_scalar.h); `d` / `scale` (partition of unity) are always real
int numfact_device(Sub *s, const HostCSR &A);
void free_factor(DeviceFactor &f);
int dense_inverse_device(Ctx *c, int N, const K *dE, K *dEinv);  // cuSOLVER LU inverse of a large coarse operator
// x = A^{-1} b for mu columns in ONE pass over the panels (column stride n), natural ordering in/out, device pointers: mu in {1, 2, 4}
// on the register-tiled kernels, any mu <= sptrsv_max_block() at or above the tensor-pipe threshold (hb_solve.cu).
// scale: optional d (natural order) applied on output (out = d .* x); accumulate: out += instead of =
int sptrsv_solve(Sub *s, const K *b, K *x, int mu, const double *scale, bool accumulate);
int sptrsv_group(int left);   // how many of `left` remaining columns the next pass takes
int sptrsv_prepare(Sub *s);   // per-device kernel attributes (opt-in shared memory) + persistent-sweep tables, once per factorisation
int sptrsv_check(Sub *s);     // after a stream synchronisation: did a persistent sweep give up waiting on a dependency?
int sptrsv_max_block();

int k_scale(Ctx *c, int n, int mu, const double *d, const K *in, K *out);      // out = d.*in
int k_axpy(Ctx *c, int64_t n, double a, const K *x, K *y);                     // y += a x
int k_copy(Ctx *c, int64_t n, const K *x, K *y);
int k_fill(Ctx *c, int64_t n, K v, K *y);
// y = beta*yin + alpha * A x, optionally scaled by d: out = d .* (...)
int k_spmv(Ctx *c, const Sub *s, int mu, double alpha, const K *x, double beta, const K *yin, K *out, const double *d);
// T[k + nu*col] = sum_i conj(Z[i,k]) d[i] x[i,col]
int k_zt_project(Ctx *c, const Sub *s, int mu, const K *x, K *T, int ldT);
// out[i,col] = d[i] * sum_k Z[i,k] Y[k,col]
int k_z_expand(Ctx *c, const Sub *s, int mu, const K *Y, int ldY, K *out);
int k_pack(Ctx *c, const Sub *s, int mu, const K *x, K *send);
int k_unpack(Ctx *c, const Sub *s, int mu, K *x);  // x[map] += d_recv, deterministic order
int k_dot(Ctx *c, const Sub *s, int mu, const K *x, const K *y, K *res);  // res[col] += sum_i d_i conj(x_i) y_i
int k_coarse_solve(Ctx *c, int mu);
// res[col] += sum_i d_i |b_i|^2 with penalised boundary rows divided by HPDDM_PEN (initializeNorm, iterative.hpp:455-468)
int k_rhs_norm(Ctx *c, const Sub *s, int mu, const K *b, double *res);
// Schwarz::computeResidual reductions (schwarz.hpp:761-803): res[2 col] from f, res[2 col + 1] from t = A x - f off the boundary rows;
// norm: 0 = l2 (sums of squares), 1 = l1, 2 = l-infinity (HPDDM_COMPUTE_RESIDUAL_*)
int k_residual_norms(Ctx *c, const Sub *s, int mu, int norm, const K *f, const K *t, double *res);
// ||b||_D per column over all subdomains and processes, penalised rows rescaled (host result)
int rhs_norms(Ctx *c, const std::vector<const K *> &b, int mu, std::vector<double> &out);  // d_Y = E^{-1} d_T with one refinement step
int k_bc(Ctx *c, const Sub *s, int mu, const K *b, K *x);

int k_zt_raw(Ctx *c, int n, int nu, const K *Z, const double *d, int mu, const K *x, K *T, int ldT);
int k_zexp_raw(Ctx *c, int n, int nu, const K *Z, const double *d, int mu, const K *Y, int ldY, K *out);
int k_spmv_raw(Ctx *c, int n, int64_t nnz, const int *ia, const int *ja, const K *a, int mu, double alpha, const K *x, double beta, const K *yin, K *out,
               const double *d);
int k_flush_tiny(Ctx *c, int64_t n, double tiny, K *v);
int to_host_csr(int n, int nnz, const int *ia, const int *ja, const K *a, int sym, char numbering, HostCSR &H);
int solve_cols(Sub *s, const K *b, K *x, int mu, const double *scale, bool acc);
// orchestration helpers shared by hb_api.cu and hb_krylov.cu (device pointers, one per local subdomain)
void plan_halo_messages(const std::vector<int> &granks, const std::vector<int> &nbcnt, const std::vector<int> &nbr, std::vector<int> &sends, std::vector<int> &recvs);
void plan_coarse_layout(const std::vector<int> &rows_per_proc, std::vector<int> &off, int &lmax);
int check_ready(Ctx *c, int mu);
int halo(Ctx *c, K *const *x, int mu);
int apply_core(Ctx *c, const std::vector<const K *> &in, const std::vector<K *> &out, int mu, int correction);
int gmv_core(Ctx *c, const std::vector<const K *> &in, const std::vector<K *> &out, int mu);
int stage_in(Ctx *c, const K *const *in, int mu, int where, std::vector<const K *> &dev);
void out_ptrs(Ctx *c, K *const *out, int where, std::vector<K *> &dev);
int stage_out(Ctx *c, K *const *out, int mu, int where);
// copy between a caller HOST pointer and device memory on the context's stream; pins the host range lazily (see Ctx::hostreg)
int host_copy(Ctx *c, void *dst, const void *src, size_t bytes, bool to_device);
void hostreg_release(Ctx *c);
int nccl_allreduce_sum(Ctx *c, double *buf, int count);
int nccl_allreduce_max(Ctx *c, double *buf, int count);  // count doubles (a K is KD doubles)
// control plane: all-gather of `bytes` per rank between HOST buffers (host callback, else NCCL through a device bounce buffer)
int ctrl_allgather(Ctx *c, const void *send, void *recv, size_t bytes);
// peer-memory fabric (hb_p2p.cu): each returns 1 = done over peer memory, 0 = caller uses NCCL, < 0 = error
int fabric_setup(Ctx *c, int mu);  // collective; called by ensure_capacity
bool fabric_on(Ctx *c);
int p2p_halo(Ctx *c, K *const *x, int mu);
int fabric_allgather(Ctx *c, K *buf, int count);               // buf: nproc blocks of `count` elements, own block in place
int fabric_allreduce(Ctx *c, double *buf, int count, int op);  // op 0: sum in rank order, 1: max
int p2p_check(Ctx *c);
void p2p_free(Ctx *c);
const K *p2p_last_halo_window(Ctx *c);  // receive slot of the last peer-memory halo round (nullptr: that round went over NCCL)
// Krylov helper kernels (hb_kernels.cu)
// V: k vectors of length n, stride ldv between them (n, or mu * n for one column of a block basis)
int k_vdots(Ctx *c, const Sub *s, int k, const K *V, int64_t ldv, const K *w, K *T);               // T[j] += sum_i d_i conj(V[i,j]) w[i]
int k_vupdate(Ctx *c, const Sub *s, int k, const K *V, int64_t ldv, const K *h, double sign, K *w);  // w += sign * V h
int k_scal_copy(Ctx *c, int64_t n, double a, const K *x, K *y);                       // y = a x
// block Krylov helpers: W (n x mu) += sign * V (n x k, ld n) * H (k x mu, ld ldh, device) ; W <- W * R (R mu x mu upper, device)
int k_vupdate_blk(Ctx *c, int n, int k, int mu, const K *V, const K *H, int ldh, double sign, K *W);
int k_rmul_upper(Ctx *c, int n, int mu, const K *R, K *W);
void gcrodr_release(Ctx *c);  // frees the recycled pair of the GCRO-DR driver, if any

}  // namespace hb
