// CPU stand-in for the lowest layer under hpddm_b200/csrc/hb_krylov.cu -- the CUDA runtime calls and the kernel launchers it uses --
// so that the REAL translation unit (gmres_device, cg_device, bgmres_device, gcrodr_device + DeviceBackend, krylov_entry and the
// exported hpddm_b200[z]_solve* entry points) can be compiled with g++ and run on a machine without a GPU.  "Device" memory is host
// memory; every launcher below does, with a plain loop, what the comment of its declaration in hb_internal.h says the kernel does.
// The "decomposition" is a global CSR matrix split into consecutive row blocks without overlap (halo = nothing to add, d = 1), the
// preconditioner is the identity or the diagonal (block-diagonal, so it splits over the blocks), the operator product couples the
// blocks.  TEST INFRASTRUCTURE ONLY (tests/test_cpu_krylov_mock.py); nothing here is linked into the product.
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "../../hpddm_b200/csrc/hb_internal.h"

// ---- CUDA runtime: host memory, synchronous "stream"
extern "C" {
cudaError_t cudaMalloc(void **p, size_t bytes) {
  *p = malloc(bytes ? bytes : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) {
  free(p);
  return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t bytes, cudaStream_t) {
  memset(p, v, bytes);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
  memmove(dst, src, bytes);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "mock"; }
}

namespace hb {

namespace {
struct World {  // the global operator behind gmv_core / apply_core
  int n = 0;
  std::vector<int> ia, ja;
  std::vector<K> a, diag;
  bool jacobi = false;
  std::vector<int> off;  // row offset of every block
  char err[512] = "";
} g;
}  // namespace

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g.err, sizeof(g.err), fmt, ap);
  va_end(ap);
}
int check_ready(Ctx *, int) { return 0; }
int sptrsv_check(Sub *) { return 0; }
int p2p_check(Ctx *) { return 0; }
int nccl_allreduce_sum(Ctx *c, double *, int) {
  c->launches++;
  return 0;  // one process
}
int halo(Ctx *, K *const *, int) { return 0; }  // no overlap: nothing to add

int k_scale(Ctx *c, int n, int mu, const double *d, const K *in, K *out) {  // out = d .* in
  for (int col = 0; col < mu; ++col)
    for (int i = 0; i < n; ++i) out[i + (size_t)col * n] = d[i] * in[i + (size_t)col * n];
  c->launches++;
  return 0;
}
int k_axpy(Ctx *c, int64_t n, double a, const K *x, K *y) {  // y += a x
  for (int64_t i = 0; i < n; ++i) y[i] = y[i] + a * x[i];
  c->launches++;
  return 0;
}
int k_copy(Ctx *c, int64_t n, const K *x, K *y) {
  if (x != y) memmove(y, x, n * sizeof(K));
  c->launches++;
  return 0;
}
int k_scal_copy(Ctx *c, int64_t n, double a, const K *x, K *y) {  // y = a x
  for (int64_t i = 0; i < n; ++i) y[i] = a * x[i];
  c->launches++;
  return 0;
}
int k_dot(Ctx *c, const Sub *s, int mu, const K *x, const K *y, K *res) {  // res[col] += sum_i d_i conj(x_i) y_i
  for (int col = 0; col < mu; ++col)
    for (int i = 0; i < s->n; ++i) res[col] = res[col] + s->d_d[i] * (hb_conj(x[i + (size_t)col * s->n]) * y[i + (size_t)col * s->n]);
  c->launches++;
  return 0;
}
int k_vdots(Ctx *c, const Sub *s, int k, const K *V, int64_t ldv, const K *w, K *T) {  // T[j] += sum_i d_i conj(V[i,j]) w[i]
  for (int j = 0; j < k; ++j)
    for (int i = 0; i < s->n; ++i) T[j] = T[j] + s->d_d[i] * (hb_conj(V[i + j * ldv]) * w[i]);
  c->launches++;
  return 0;
}
int k_vupdate(Ctx *c, const Sub *s, int k, const K *V, int64_t ldv, const K *h, double sign, K *w) {  // w += sign * V h
  for (int i = 0; i < s->n; ++i) {
    K acc = mk(0.0);
    for (int j = 0; j < k; ++j) acc = acc + V[i + j * ldv] * h[j];
    w[i] = w[i] + sign * acc;
  }
  c->launches++;
  return 0;
}
int k_zt_raw(Ctx *c, int n, int nu, const K *Z, const double *d, int mu, const K *x, K *T, int ldT) {  // T[k + ldT col] += sum_i conj(Z[i,k]) d_i x[i,col]
  for (int col = 0; col < mu; ++col)
    for (int k = 0; k < nu; ++k)
      for (int i = 0; i < n; ++i) T[k + (size_t)ldT * col] = T[k + (size_t)ldT * col] + d[i] * (hb_conj(Z[i + (size_t)k * n]) * x[i + (size_t)col * n]);
  c->launches++;
  return 0;
}
int k_vupdate_blk(Ctx *c, int n, int k, int mu, const K *V, const K *H, int ldh, double sign, K *W) {  // W += sign * V H
  for (int col = 0; col < mu; ++col)
    for (int i = 0; i < n; ++i) {
      K acc = mk(0.0);
      for (int j = 0; j < k; ++j) acc = acc + V[i + (size_t)j * n] * H[j + (size_t)ldh * col];
      W[i + (size_t)col * n] = W[i + (size_t)col * n] + sign * acc;
    }
  c->launches++;
  return 0;
}
int k_rmul_upper(Ctx *c, int n, int mu, const K *R, K *W) {  // W <- W R, R upper triangular mu x mu (ld mu)
  std::vector<K> row(mu);
  for (int i = 0; i < n; ++i) {
    for (int col = 0; col < mu; ++col) {
      K acc = mk(0.0);
      for (int l = 0; l <= col; ++l) acc = acc + W[i + (size_t)l * n] * R[l + (size_t)col * mu];
      row[col] = acc;
    }
    for (int col = 0; col < mu; ++col) W[i + (size_t)col * n] = row[col];
  }
  c->launches++;
  return 0;
}
int k_bc(Ctx *c, const Sub *s, int mu, const K *b, K *x) {  // penalised rows: x = b / value
  for (const auto &e : s->bc)
    for (int col = 0; col < mu; ++col) x[e.first + (size_t)col * s->n] = b[e.first + (size_t)col * s->n] / e.second;
  c->launches++;
  return 0;
}
int rhs_norms(Ctx *c, const std::vector<const K *> &b, int mu, std::vector<double> &out) {
  out.assign(mu, 0.0);
  for (size_t q = 0; q < c->subs.size(); ++q)
    for (int col = 0; col < mu; ++col)
      for (int i = 0; i < c->subs[q]->n; ++i) out[col] += c->subs[q]->d_d[i] * hb_norm(b[q][i + (size_t)col * c->subs[q]->n]);
  for (double &v : out) v = std::sqrt(v);
  return 0;
}
int apply_core(Ctx *c, const std::vector<const K *> &in, const std::vector<K *> &out, int mu, int) {  // identity or Jacobi
  for (size_t q = 0; q < c->subs.size(); ++q) {
    const int n = c->subs[q]->n;
    for (int col = 0; col < mu; ++col)
      for (int i = 0; i < n; ++i) out[q][i + (size_t)col * n] = g.jacobi ? in[q][i + (size_t)col * n] / g.diag[g.off[q] + i] : in[q][i + (size_t)col * n];
  }
  c->launches++;
  return 0;
}
int gmv_core(Ctx *c, const std::vector<const K *> &in, const std::vector<K *> &out, int mu) {  // global CSR product across the blocks
  std::vector<K> x(g.n);
  for (int col = 0; col < mu; ++col) {
    for (size_t q = 0; q < c->subs.size(); ++q)
      for (int i = 0; i < c->subs[q]->n; ++i) x[g.off[q] + i] = in[q][i + (size_t)col * c->subs[q]->n];
    for (size_t q = 0; q < c->subs.size(); ++q)
      for (int i = 0; i < c->subs[q]->n; ++i) {
        K acc = mk(0.0);
        const int row = g.off[q] + i;
        for (int p = g.ia[row]; p < g.ia[row + 1]; ++p) acc = acc + g.a[p] * x[g.ja[p]];
        out[q][i + (size_t)col * c->subs[q]->n] = acc;
      }
  }
  c->launches++;
  return 0;
}

}  // namespace hb

using namespace hb;

extern "C" {

// C-numbered CSR; `sizes`: rows of every block (sum = n).  Returns the context handle the hpddm_b200[z]_solve* entry points take.
void *krylov_mock_create(int n, const int *ia, const int *ja, const K *a, int nsub, const int *sizes, int jacobi) {
  g.n = n;
  g.ia.assign(ia, ia + n + 1);
  g.ja.assign(ja, ja + ia[n]);
  g.jacobi = jacobi != 0;
  g.off.assign(1, 0);
  Ctx *c = new Ctx();
  for (int q = 0; q < nsub; ++q) {
    Sub *s = new Sub();
    s->ctx = c;
    s->grank = q;
    s->n = sizes[q];
    s->d_d = new double[sizes[q] > 0 ? sizes[q] : 1];
    for (int i = 0; i < sizes[q]; ++i) s->d_d[i] = 1.0;
    c->subs.push_back(s);
    g.off.push_back(g.off.back() + sizes[q]);
  }
  g.a.assign(a, a + ia[n]);
  g.diag.assign(n, mk(1.0));
  for (int i = 0; i < n; ++i)
    for (int p = ia[i]; p < ia[i + 1]; ++p)
      if (ja[p] == i) g.diag[i] = a[p];
  return c;
}
// new values on the same pattern (a sequence of slowly changing systems)
void krylov_mock_set_values(const K *a) {
  g.a.assign(a, a + g.ia[g.n]);
  for (int i = 0; i < g.n; ++i)
    for (int p = g.ia[i]; p < g.ia[i + 1]; ++p)
      if (g.ja[p] == i) g.diag[i] = a[p];
}
// Prcndtnr of every block (HPDDM_B200_PRCNDTNR_*): solve_cg runs CG only for the symmetric ones, like the reference (CG.hpp:41-44)
void krylov_mock_set_prcndtnr(void *ctx, int value) {
  for (Sub *s : static_cast<Ctx *>(ctx)->subs) s->prcndtnr = value;
}
long krylov_mock_launches(void *ctx) { return (long)static_cast<Ctx *>(ctx)->launches; }
const char *krylov_mock_error() { return g.err; }
void krylov_mock_destroy(void *ctx) {
  Ctx *c = static_cast<Ctx *>(ctx);
  gcrodr_release(c);
  for (Sub *s : c->subs) {
    delete[] s->d_d;
    delete s;
  }
  delete c;
}
}
