// Internal data structures of libhpddm_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/hpddm_b200.h"

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
constexpr int FCH = 512;    // forward column chunk
constexpr int BCH = 256;    // backward column chunk (8 accumulators per lane)
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
  std::vector<double> a;
  bool symmetric = false;   // numerically symmetric
};

// builds ordering + supernodal structure.  grid hint: nx*ny*nz*dof == n or nx==0
int symbolic_analyze(const HostCSR &A, int nx, int ny, int nz, int dof, int leaf, Symbolic &S);

// ---------------------------------------------------------------- device side
struct DeviceFactor {
  bool valid = false;
  bool symmetric = true;
  double *panL = nullptr;  // forward panels
  double *panU = nullptr;  // backward panels (== panL when symmetric)
  Front *fronts = nullptr;
  int *rowidx = nullptr;
  FwdItem *fwd = nullptr;
  BwdItem *bwd = nullptr;
  int *perm = nullptr;   // perm[new] = old
  double *b = nullptr, *y = nullptr, *x = nullptr;  // permuted work vectors (n * 4 each: up to 4 RHS per pass)
  cudaGraphExec_t graph[3] = {nullptr, nullptr, nullptr};  // captured sweep launches for mu = 1, 2, 4
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
  double *d_a = nullptr;
  double *d_d = nullptr;
  int *d_map = nullptr;     // concatenated neighbour indices (h)
  int *d_ebase = nullptr, *d_esize = nullptr;  // per entry: start / size of its neighbour segment
  std::vector<int> peer_seg;  // for a local neighbour: its segment index that points back to us
  int h = 0;
  double *d_send = nullptr, *d_recv = nullptr;  // h * mu_cap each
  int *d_uidx = nullptr, *d_useg = nullptr, *d_upos = nullptr;  // deterministic unpack (CSR by unique target)
  int nuniq = 0;
  std::vector<std::pair<int, double>> bc;  // penalised rows
  int *d_bc_idx = nullptr;
  double *d_bc_val = nullptr;
  // factor
  Symbolic sym;
  DeviceFactor fac;
  int prcndtnr = HPDDM_B200_PRCNDTNR_GE;
  double t_symbolic = 0, t_numfact = 0;
  // deflation
  int nu = 0;
  double *d_Z = nullptr;
  int coff = 0;             // offset of this subdomain in the coarse vector
  // work vectors (n * mu_cap)
  double *d_in = nullptr, *d_out = nullptr, *d_work = nullptr, *d_tmp = nullptr;
  int mu_cap = 0;
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::vector<Sub *> subs;
  int64_t launches = 0;
  // communicator
  void *nccl = nullptr;  // ncclComm_t
  int proc_rank = 0, nproc = 1;
  // coarse
  int Nc = 0;
  std::vector<int> coarse_off;  // per global rank, size P+1
  std::vector<int> nu_all;
  double *d_E = nullptr, *d_Einv = nullptr;  // Nc x Nc column-major
  double *d_T = nullptr, *d_Y = nullptr;     // Nc x mu_cap, layout [proc][col][row-in-proc]
  double *d_R = nullptr;                     // Nc residual of the coarse refinement step
  int Lnu = 0;                               // Lmax: coarse rows per process block in the (padded) communication layout
  std::vector<int> Lnu_p;                    // actual coarse rows of every process
  int *d_rowproc = nullptr, *d_rowloc = nullptr;  // coarse row -> (process, row inside its block)
  int loc_off = 0;                           // first coarse row of this process
  double *d_res = nullptr;                   // small device scratch (dots)
  std::vector<double> E_host;
  // pinned staging
  double *pin = nullptr;
  size_t pin_bytes = 0;
  int mu_cap = 0;
  bool started = false;
  P2P *p2p = nullptr;  // NVLink peer-memory halo state (hb_p2p.cu)
};

// ---------------------------------------------------------------- kernels (launchers)
int numfact_device(Sub *s, const HostCSR &A);
void free_factor(DeviceFactor &f);
// x = A^{-1} b for mu in {1,2,4} columns (column stride n), natural ordering in/out, device pointers.
// scale: optional d (natural order) applied on output (out = d .* x); accumulate: out += instead of =
int sptrsv_solve(Sub *s, const double *b, double *x, int mu, const double *scale, bool accumulate);

int k_scale(Ctx *c, int n, int mu, const double *d, const double *in, double *out);      // out = d.*in
int k_axpy(Ctx *c, int64_t n, double a, const double *x, double *y);                     // y += a x
int k_copy(Ctx *c, int64_t n, const double *x, double *y);
int k_fill(Ctx *c, int64_t n, double v, double *y);
// y = beta*yin + alpha * A x, optionally scaled by d: out = d .* (...)
int k_spmv(Ctx *c, const Sub *s, int mu, double alpha, const double *x, double beta, const double *yin, double *out, const double *d);
// T[k + nu*col] = sum_i Z[i,k] d[i] x[i,col]
int k_zt_project(Ctx *c, const Sub *s, int mu, const double *x, double *T, int ldT);
// out[i,col] = d[i] * sum_k Z[i,k] Y[k,col]
int k_z_expand(Ctx *c, const Sub *s, int mu, const double *Y, int ldY, double *out);
int k_pack(Ctx *c, const Sub *s, int mu, const double *x, double *send);
int k_unpack(Ctx *c, const Sub *s, int mu, double *x);  // x[map] += d_recv, deterministic order
int k_dot(Ctx *c, const Sub *s, int mu, const double *x, const double *y, double *res);
int k_coarse_solve(Ctx *c, int mu);  // d_Y = E^{-1} d_T with one refinement step
int k_bc(Ctx *c, const Sub *s, int mu, const double *b, double *x);

int k_zt_raw(Ctx *c, int n, int nu, const double *Z, const double *d, int mu, const double *x, double *T, int ldT);
int k_zexp_raw(Ctx *c, int n, int nu, const double *Z, const double *d, int mu, const double *Y, int ldY, double *out);
int k_spmv_raw(Ctx *c, int n, int64_t nnz, const int *ia, const int *ja, const double *a, int mu, double alpha, const double *x, double beta, const double *yin,
               double *out, const double *d);
int k_flush_tiny(Ctx *c, int64_t n, double tiny, double *v);
int to_host_csr(int n, int nnz, const int *ia, const int *ja, const double *a, int sym, char numbering, HostCSR &H);
int solve_cols(Sub *s, const double *b, double *x, int mu, const double *scale, bool acc);
// orchestration helpers shared by hb_api.cu and hb_krylov.cu (device pointers, one per local subdomain)
int check_ready(Ctx *c, int mu);
int halo(Ctx *c, double *const *x, int mu, bool allow_p2p = true);
int apply_core(Ctx *c, const std::vector<const double *> &in, const std::vector<double *> &out, int mu, int correction);
int gmv_core(Ctx *c, const std::vector<const double *> &in, const std::vector<double *> &out, int mu);
int stage_in(Ctx *c, const double *const *in, int mu, int where, std::vector<const double *> &dev);
void out_ptrs(Ctx *c, double *const *out, int where, std::vector<double *> &dev);
int stage_out(Ctx *c, double *const *out, int mu, int where);
int nccl_allreduce_sum(Ctx *c, double *buf, int count);
int nccl_allgather_bytes(Ctx *c, const void *send, void *recv, size_t bytes_per_rank);
int p2p_halo(Ctx *c, double *const *x, int mu);  // 1 = done over peer memory, 0 = use NCCL
int p2p_check(Ctx *c);
void p2p_free(Ctx *c);
// Krylov helper kernels (hb_kernels.cu)
int k_vdots(Ctx *c, const Sub *s, int k, const double *V, const double *w, double *T);      // T[j] += sum_i d_i V[i,j] w[i]
int k_vupdate(Ctx *c, const Sub *s, int k, const double *V, const double *h, double sign, double *w);  // w += sign * V h
int k_scal_copy(Ctx *c, int64_t n, double a, const double *x, double *y);                  // y = a x

}  // namespace hb
