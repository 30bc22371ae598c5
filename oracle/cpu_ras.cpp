// CPU arm of the benchmark ("port" of the reference path for the host cores).
// TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never used by the product path.
//
// One subdomain of the BASELINE workload (3-D 7-point Poisson m^3, two-level deflated RAS, nu
// vectors) solved the way the reference's CPU path does it: a supernodal sparse Cholesky of the
// local matrix (what SUBDOMAIN = CHOLMOD / MUMPS computes in numfact, include/HPDDM_SuiteSparse.hpp:
// 264-371) whose solve phase is a forward and a backward substitution over the supernodes with
// BLAS-2 kernels (dtrsv + dgemv per supernode, as CHOLMOD's supernodal solve does for one RHS),
// plus the apply chain of Schwarz::apply (include/HPDDM_schwarz.hpp:572-608): two dense projections
// with Z (dgemv, schwarz.hpp:1616,1618), one CSR SpMV (wrapper.hpp:697-733, OpenMP like the
// reference), the coarse solve and the vector updates.  With a single subdomain the halo exchanges
// are no-ops and d = 1.  Threads: OpenMP over independent supernodes of a level + threaded OpenBLAS
// on the large ones (the reference = MPI ranks x OpenMP/BLAS threads on the same cores).
//
// Ordering: its own geometric nested dissection (independent of the product's analysis code).
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" {
void scipy_dpotrf_(const char *, const int *, double *, const int *, int *);
void scipy_dtrsm_(const char *, const char *, const char *, const char *, const int *, const int *, const double *, const double *, const int *, double *, const int *);
void scipy_dsyrk_(const char *, const char *, const int *, const int *, const double *, const double *, const int *, const double *, double *, const int *);
void scipy_dtrsv_(const char *, const char *, const char *, const int *, const double *, const int *, double *, const int *);
void scipy_dgemv_(const char *, const int *, const int *, const double *, const double *, const int *, const double *, const int *, const double *, double *, const int *);
void scipy_dpotrs_(const char *, const int *, const int *, const double *, const int *, double *, const int *, int *);
void scipy_openblas_set_num_threads(int);
}

namespace {

struct Front {
  int p0, s1, s2, parent, level;
  long rptr;
  std::vector<double> L;  // (s1+s2) x s1 column-major: [L11; L21]
};

struct Problem {
  int m, n, nu, nthreads;
  std::vector<int> ia, ja;
  std::vector<double> a;
  std::vector<int> perm, iperm, front_of;
  std::vector<Front> fronts;
  std::vector<int> rowidx, rel;
  std::vector<std::vector<int>> children, levels;
  std::vector<double> Z, E;  // Z n x nu col-major, E nu x nu Cholesky factor
  std::vector<double> work, q, t;
  long nnzL = 0;
  double t_fact = 0;
};

struct Box { int lo[3], hi[3]; };

void nd(const Box &b, int m, int leaf, std::vector<int> &order, std::vector<int> &fend) {
  int ext[3] = {b.hi[0] - b.lo[0], b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]};
  if (ext[0] <= 0 || ext[1] <= 0 || ext[2] <= 0) return;
  auto emit = [&](const Box &x) {
    for (int k = x.lo[2]; k < x.hi[2]; ++k)
      for (int j = x.lo[1]; j < x.hi[1]; ++j)
        for (int i = x.lo[0]; i < x.hi[0]; ++i) order.push_back((k * m + j) * m + i);
    fend.push_back((int)order.size());
  };
  int ax = 0;
  if (ext[1] > ext[ax]) ax = 1;
  if (ext[2] > ext[ax]) ax = 2;
  if ((long)ext[0] * ext[1] * ext[2] <= leaf || ext[ax] < 3) { emit(b); return; }
  int mid = b.lo[ax] + ext[ax] / 2;
  Box l = b, r = b, s = b;
  l.hi[ax] = mid; r.lo[ax] = mid + 1; s.lo[ax] = mid; s.hi[ax] = mid + 1;
  nd(l, m, leaf, order, fend);
  nd(r, m, leaf, order, fend);
  emit(s);
}

void analyze(Problem &P) {
  const int n = P.n;
  std::vector<int> fend;
  Box b{{0, 0, 0}, {P.m, P.m, P.m}};
  nd(b, P.m, 64, P.perm, fend);
  P.iperm.assign(n, 0);
  for (int i = 0; i < n; ++i) P.iperm[P.perm[i]] = i;
  const int F = (int)fend.size();
  P.fronts.resize(F);
  P.front_of.resize(n);
  int p = 0;
  for (int f = 0; f < F; ++f) {
    P.fronts[f].p0 = p; P.fronts[f].s1 = fend[f] - p; P.fronts[f].parent = -1;
    for (; p < fend[f]; ++p) P.front_of[p] = f;
  }
  P.children.assign(F, {});
  std::vector<int> mark(n, -1), cur;
  std::vector<std::vector<int>> st(F);
  for (int f = 0; f < F; ++f) {
    Front &fr = P.fronts[f];
    const int last = fr.p0 + fr.s1 - 1;
    cur.clear();
    auto add = [&](int q) { if (q > last && mark[q] != f) { mark[q] = f; cur.push_back(q); } };
    for (int q = fr.p0; q <= last; ++q) { int v = P.perm[q]; for (int k = P.ia[v]; k < P.ia[v + 1]; ++k) add(P.iperm[P.ja[k]]); }
    for (int c : P.children[f]) for (int q : st[c]) add(q);
    std::sort(cur.begin(), cur.end());
    st[f] = cur; fr.s2 = (int)cur.size();
    if (!cur.empty()) { fr.parent = P.front_of[cur[0]]; P.children[fr.parent].push_back(f); }
  }
  long tot = 0;
  for (int f = 0; f < F; ++f) { P.fronts[f].rptr = tot; tot += P.fronts[f].s2; }
  P.rowidx.resize(tot); P.rel.resize(tot);
  for (int f = 0; f < F; ++f) {
    Front &fr = P.fronts[f];
    std::copy(st[f].begin(), st[f].end(), P.rowidx.begin() + fr.rptr);
    if (fr.parent >= 0) {
      const Front &pf = P.fronts[fr.parent]; const std::vector<int> &ps = st[fr.parent];
      for (int i = 0; i < fr.s2; ++i) { int q = st[f][i]; P.rel[fr.rptr + i] = q < pf.p0 + pf.s1 ? q - pf.p0 : pf.s1 + (int)(std::lower_bound(ps.begin(), ps.end(), q) - ps.begin()); }
    }
  }
  std::vector<int> depth(F, 0); int maxd = 0;
  for (int f = F - 1; f >= 0; --f) { depth[f] = P.fronts[f].parent < 0 ? 0 : depth[P.fronts[f].parent] + 1; maxd = std::max(maxd, depth[f]); }
  P.levels.assign(maxd + 1, {});
  for (int f = 0; f < F; ++f) { P.fronts[f].level = maxd - depth[f]; P.levels[P.fronts[f].level].push_back(f); }
}

void factor(Problem &P) {
  const int F = (int)P.fronts.size();
  std::vector<std::vector<double>> U(F);  // update matrices: packed lower triangles (column j starts at j*s2 - j(j-1)/2)
  const double one = 1.0, mone = -1.0;
  for (size_t l = 0; l < P.levels.size(); ++l) {
    const std::vector<int> &lv = P.levels[l];
    const bool par = (int)lv.size() >= 2 * P.nthreads;
    scipy_openblas_set_num_threads(par ? 1 : P.nthreads);
    auto body = [&](int f) {
      Front &fr = P.fronts[f];
      const int s1 = fr.s1, s2 = fr.s2, s = s1 + s2;
      std::vector<double> M((size_t)s * s, 0.0);
      auto loc = [&](int q) { if (q < fr.p0 + s1) return q - fr.p0; const int *b = P.rowidx.data() + fr.rptr; return s1 + (int)(std::lower_bound(b, b + s2, q) - b); };
      for (int q = fr.p0; q < fr.p0 + s1; ++q) {
        const int v = P.perm[q];
        for (int k = P.ia[v]; k < P.ia[v + 1]; ++k) { const int pi = P.iperm[P.ja[k]]; if (pi >= q) M[loc(pi) + (size_t)(q - fr.p0) * s] += P.a[k]; }
      }
      for (int c : P.children[f]) {
        const Front &cf = P.fronts[c]; const int *r = P.rel.data() + cf.rptr; const std::vector<double> &Uc = U[c];
        for (int j = 0; j < cf.s2; ++j) {
          const double *col = Uc.data() + ((size_t)j * cf.s2 - (size_t)j * (j - 1) / 2) - j;  // col[i] = U(i, j), i >= j
          for (int i = j; i < cf.s2; ++i) M[r[i] + (size_t)r[j] * s] += col[i];
        }
        std::vector<double>().swap(U[c]);
      }
      int info = 0;
      scipy_dpotrf_("L", &s1, M.data(), &s, &info);
      if (info != 0) { fprintf(stderr, "cpu_ras: potrf failed front %d info %d\n", f, info); abort(); }
      if (s2 > 0) {
        scipy_dtrsm_("R", "L", "T", "N", &s2, &s1, &one, M.data(), &s, M.data() + s1, &s);
        scipy_dsyrk_("L", "N", &s2, &s1, &mone, M.data() + s1, &s, &one, M.data() + s1 + (size_t)s1 * s, &s);
        U[f].resize((size_t)s2 * (s2 + 1) / 2);
        for (int j = 0; j < s2; ++j) memcpy(&U[f][(size_t)j * s2 - (size_t)j * (j - 1) / 2], &M[s1 + j + (size_t)(s1 + j) * s], (size_t)(s2 - j) * sizeof(double));
      }
      fr.L.resize((size_t)s * s1);
      for (int j = 0; j < s1; ++j) memcpy(&fr.L[(size_t)j * s], &M[(size_t)j * s], (size_t)s * sizeof(double));
    };
    if (par) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(P.nthreads)
      for (int q = 0; q < (int)lv.size(); ++q) body(lv[q]);
    } else
      for (int f : lv) body(f);
  }
  P.nnzL = 0;
  for (const Front &fr : P.fronts) P.nnzL += (long)fr.s1 * (fr.s1 + 1) / 2 + (long)fr.s1 * fr.s2;
}

// ---- threaded BLAS-2 building blocks (OpenBLAS pinned to 1 thread; OpenMP does the splitting)
static void par_gemv_n(int nt, int mrows, int ncols, double alpha, const double *A, int lda, const double *x, double beta, double *y) {
  const int one = 1;
  if ((long)mrows * ncols < 200000 || nt == 1) { scipy_dgemv_("N", &mrows, &ncols, &alpha, A, &lda, x, &one, &beta, y, &one); return; }
  const int chunk = std::max(64, (mrows + nt - 1) / nt);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
  for (int r0 = 0; r0 < mrows; r0 += chunk) {
    const int nr = std::min(chunk, mrows - r0);
    scipy_dgemv_("N", &nr, &ncols, &alpha, A + r0, &lda, x, &one, &beta, y + r0, &one);
  }
}
static void par_gemv_t(int nt, int mrows, int ncols, double alpha, const double *A, int lda, const double *x, double beta, double *y) {
  const int one = 1;
  if ((long)mrows * ncols < 200000 || nt == 1) { scipy_dgemv_("T", &mrows, &ncols, &alpha, A, &lda, x, &one, &beta, y, &one); return; }
  const int chunk = std::max(16, (ncols + nt - 1) / nt);
#pragma omp parallel for schedule(static, 1) num_threads(nt)
  for (int c0 = 0; c0 < ncols; c0 += chunk) {
    const int nc = std::min(chunk, ncols - c0);
    scipy_dgemv_("T", &mrows, &nc, &alpha, A + (size_t)c0 * lda, &lda, x, &one, &beta, y + c0, &one);
  }
}
// blocked triangular solves with the lower factor L (n x n, ld): L x = b and L^T x = b
static void par_trsv_n(int nt, int n, const double *L, int ld, double *x) {
  const int one = 1, NB = 256;
  for (int k = 0; k < n; k += NB) {
    const int nb = std::min(NB, n - k);
    scipy_dtrsv_("L", "N", "N", &nb, L + k + (size_t)k * ld, &ld, x + k, &one);
    if (k + nb < n) par_gemv_n(nt, n - k - nb, nb, -1.0, L + (k + nb) + (size_t)k * ld, ld, x + k, 1.0, x + k + nb);
  }
}
static void par_trsv_t(int nt, int n, const double *L, int ld, double *x) {
  const int one = 1, NB = 256;
  for (int k = ((n - 1) / NB) * NB; k >= 0; k -= NB) {
    const int nb = std::min(NB, n - k);
    scipy_dtrsv_("L", "T", "N", &nb, L + k + (size_t)k * ld, &ld, x + k, &one);
    if (k > 0) par_gemv_t(nt, nb, k, -1.0, L + k, ld, x + k, 1.0, x);
  }
}

// x <- A^-1 x in permuted space; supernodal forward/backward substitution (dtrsv + dgemv per front)
void solve(Problem &P, double *x) {
  std::vector<std::vector<double>> tmp(P.nthreads);  // per-thread scratch
  const bool prof = getenv("CPU_RAS_PROFILE") != nullptr;
  for (size_t l = 0; l < P.levels.size(); ++l) {
    auto tl = std::chrono::steady_clock::now();
    const std::vector<int> &lv = P.levels[l];
    const bool par = (int)lv.size() >= 2 * P.nthreads;
    auto body = [&](int f, std::vector<double> &w, int nt) {
      Front &fr = P.fronts[f];
      const int s1 = fr.s1, s2 = fr.s2, s = s1 + s2;
      double *x1 = x + fr.p0;
      par_trsv_n(nt, s1, fr.L.data(), s, x1);
      if (s2 > 0) {
        w.resize(s2);
        par_gemv_n(nt, s2, s1, 1.0, fr.L.data() + s1, s, x1, 0.0, w.data());
        const int *r = P.rowidx.data() + fr.rptr;
        if (nt == 1) {
          for (int i = 0; i < s2; ++i) {
#pragma omp atomic
            x[r[i]] -= w[i];
          }
        } else
          for (int i = 0; i < s2; ++i) x[r[i]] -= w[i];
      }
    };
    if (par) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(P.nthreads)
      for (int q = 0; q < (int)lv.size(); ++q) body(lv[q], tmp[omp_get_thread_num()], 1);
    } else
      for (int f : lv) body(f, tmp[0], P.nthreads);
    if (prof) fprintf(stderr, "fwd level %zu: %zu fronts par=%d %.3f ms\n", l, lv.size(), (int)par, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tl).count());
  }
  for (int l = (int)P.levels.size() - 1; l >= 0; --l) {
    const std::vector<int> &lv = P.levels[l];
    const bool par = (int)lv.size() >= 2 * P.nthreads;
    auto body = [&](int f, std::vector<double> &w, int nt) {
      Front &fr = P.fronts[f];
      const int s1 = fr.s1, s2 = fr.s2, s = s1 + s2;
      double *x1 = x + fr.p0;
      if (s2 > 0) {
        w.resize(s2);
        const int *r = P.rowidx.data() + fr.rptr;
        for (int i = 0; i < s2; ++i) w[i] = x[r[i]];
        par_gemv_t(nt, s2, s1, -1.0, fr.L.data() + s1, s, w.data(), 1.0, x1);
      }
      par_trsv_t(nt, s1, fr.L.data(), s, x1);
    };
    if (par) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(P.nthreads)
      for (int q = 0; q < (int)lv.size(); ++q) body(lv[q], tmp[omp_get_thread_num()], 1);
    } else
      for (int f : lv) body(f, tmp[0], P.nthreads);
  }
}

void spmv(const Problem &P, const double *x, double *y, int nt) {
#pragma omp parallel for schedule(static, 50000) num_threads(nt)
  for (int i = 0; i < P.n; ++i) {
    double acc = 0.0;
    for (int k = P.ia[i]; k < P.ia[i + 1]; ++k) acc += P.a[k] * x[P.ja[k]];
    y[i] = acc;
  }
}

}  // namespace

extern "C" {

// builds the m^3 Poisson subdomain, factors it, installs Z (n x nu col-major) and E = Z^T A Z
void *cpu_ras_create(int m, int nu, const double *Z, int nthreads) {
  Problem *P = new Problem;
  P->m = m; P->n = m * m * m; P->nu = nu; P->nthreads = nthreads > 0 ? nthreads : omp_get_max_threads();
  const int n = P->n;
  const double h = 10.0 / m, c = -1.0 / (h * h), dg = 6.0 / (h * h);
  P->ia.assign(1, 0);
  auto id = [&](int i, int j, int k) { return (k * m + j) * m + i; };
  for (int k = 0; k < m; ++k) for (int j = 0; j < m; ++j) for (int i = 0; i < m; ++i) {
    if (k > 0) { P->ja.push_back(id(i, j, k - 1)); P->a.push_back(c); }
    if (j > 0) { P->ja.push_back(id(i, j - 1, k)); P->a.push_back(c); }
    if (i > 0) { P->ja.push_back(id(i - 1, j, k)); P->a.push_back(c); }
    P->ja.push_back(id(i, j, k)); P->a.push_back(dg);
    if (i < m - 1) { P->ja.push_back(id(i + 1, j, k)); P->a.push_back(c); }
    if (j < m - 1) { P->ja.push_back(id(i, j + 1, k)); P->a.push_back(c); }
    if (k < m - 1) { P->ja.push_back(id(i, j, k + 1)); P->a.push_back(c); }
    P->ia.push_back((int)P->ja.size());
  }
  auto t0 = std::chrono::steady_clock::now();
  analyze(*P);
  factor(*P);
  P->t_fact = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  P->Z.assign(Z, Z + (size_t)n * nu);
  // E = Z^T A Z (dense nu x nu), Cholesky-factored
  P->E.assign((size_t)nu * nu, 0.0);
  std::vector<double> AZ(n);
  for (int k = 0; k < nu; ++k) {
    spmv(*P, P->Z.data() + (size_t)k * n, AZ.data(), P->nthreads);
    for (int l = 0; l < nu; ++l) { double acc = 0; for (int i = 0; i < n; ++i) acc += P->Z[(size_t)l * n + i] * AZ[i]; P->E[l + (size_t)k * nu] = acc; }
  }
  int info = 0;
  scipy_dpotrf_("L", &nu, P->E.data(), &nu, &info);
  P->work.resize(n); P->q.resize(n); P->t.resize(nu);
  return P;
}

// out = M^-1 in, deflated two-level apply (schwarz.hpp:572-608) for a single subdomain (d = 1, no halo)
void cpu_ras_apply(void *h, const double *in, double *out) {
  Problem &P = *(Problem *)h;
  const int n = P.n, nu = P.nu, one = 1;
  const double d1 = 1.0, d0 = 0.0;
  scipy_openblas_set_num_threads(1);
  par_gemv_t(P.nthreads, n, nu, 1.0, P.Z.data(), n, in, 0.0, P.t.data());              // uc = Z^T D in   (schwarz.hpp:1616)
  int info = 0;
  scipy_dpotrs_("L", &nu, &one, P.E.data(), &nu, P.t.data(), &nu, &info);              // callSolver       (1617)
  par_gemv_n(P.nthreads, n, nu, 1.0, P.Z.data(), n, P.t.data(), 0.0, out);             // out = Z uc       (1618)
  spmv(P, out, P.q.data(), P.nthreads);                                                  // work = in - A out (581-586)
#pragma omp parallel for schedule(static, 50000) num_threads(P.nthreads)
  for (int i = 0; i < n; ++i) P.work[P.iperm[i]] = in[i] - P.q[i];
  solve(P, P.work.data());                                                               // s_.solve(work)   (590)
#pragma omp parallel for schedule(static, 50000) num_threads(P.nthreads)
  for (int i = 0; i < n; ++i) out[i] += P.work[P.iperm[i]];                              // out += work      (607)
}

// plain local solve x = A^-1 b (natural ordering), for validation
void cpu_ras_solve(void *h, const double *b, double *x) {
  Problem &P = *(Problem *)h;
  scipy_openblas_set_num_threads(1);
  for (int i = 0; i < P.n; ++i) P.work[P.iperm[i]] = b[i];
  solve(P, P.work.data());
  for (int i = 0; i < P.n; ++i) x[i] = P.work[P.iperm[i]];
}
// host bytes cpu_ras_create(m, ...) needs at its peak (analysis only, no numerics): factor storage + the update matrices of two
// consecutive levels + the frontal matrices being factored concurrently.  bench.py uses it to refuse a size the host cannot hold.
double cpu_ras_estimate_bytes(int m, int nthreads) {
  Problem P;
  P.m = m; P.n = m * m * m; P.nthreads = nthreads > 0 ? nthreads : omp_get_max_threads();
  P.ia.assign(1, 0);
  auto id = [&](int i, int j, int k) { return (k * m + j) * m + i; };
  for (int k = 0; k < m; ++k) for (int j = 0; j < m; ++j) for (int i = 0; i < m; ++i) {
    if (k > 0) P.ja.push_back(id(i, j, k - 1));
    if (j > 0) P.ja.push_back(id(i, j - 1, k));
    if (i > 0) P.ja.push_back(id(i - 1, j, k));
    P.ja.push_back(id(i, j, k));
    if (i < m - 1) P.ja.push_back(id(i + 1, j, k));
    if (j < m - 1) P.ja.push_back(id(i, j + 1, k));
    if (k < m - 1) P.ja.push_back(id(i, j, k + 1));
    P.ia.push_back((int)P.ja.size());
  }
  analyze(P);
  double Lb = 0, peak = 0, prevU = 0;
  for (const Front &fr : P.fronts) Lb += 8.0 * (double)(fr.s1 + fr.s2) * fr.s1;
  for (size_t l = 0; l < P.levels.size(); ++l) {
    double U = 0, Mmax = 0;
    for (int f : P.levels[l]) {
      const Front &fr = P.fronts[f];
      U += 4.0 * (double)fr.s2 * (fr.s2 + 1);
      Mmax = std::max(Mmax, 8.0 * (double)(fr.s1 + fr.s2) * (fr.s1 + fr.s2));
    }
    const bool par = (int)P.levels[l].size() >= 2 * P.nthreads;
    peak = std::max(peak, prevU + U + Mmax * (par ? P.nthreads : 1));
    prevU = U;
  }
  return Lb + peak + 8.0 * P.n * 40 + 12.0 * P.ja.size();
}
long cpu_ras_nnz_factor(void *h) { return ((Problem *)h)->nnzL; }
double cpu_ras_factor_seconds(void *h) { return ((Problem *)h)->t_fact; }
int cpu_ras_threads(void *h) { return ((Problem *)h)->nthreads; }
void cpu_ras_destroy(void *h) { delete (Problem *)h; }
}
