// Host-only check of hb_symbolic.cpp: run the analysis on a 3-D Poisson matrix, then a plain
// dense multifrontal Cholesky driven by the same structures (struct rows, rel indices, levels)
// and verify the pivots; then check that the forward / backward work items of the SpTRSV tile every structurally
// non-zero panel entry exactly once (compiled for both scalar builds: FCH / BCH depend on sizeof(K)).
// Usage: test_symbolic nx ny nz [hint(1/0)] [leaf]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../hpddm_b200/csrc/hb_internal.h"
using namespace hb;
int main(int argc, char **argv) {
  int nx = atoi(argv[1]), ny = atoi(argv[2]), nz = atoi(argv[3]);
  int hint = argc > 4 ? atoi(argv[4]) : 1, leaf = argc > 5 ? atoi(argv[5]) : 64;
  int n = nx * ny * nz;
  HostCSR A; A.n = n; A.ia.assign(1, 0); A.symmetric = true;
  auto id = [&](int i, int j, int k) { return (k * ny + j) * nx + i; };
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
    if (k > 0) { A.ja.push_back(id(i, j, k - 1)); A.a.push_back(mk(-1.0)); }
    if (j > 0) { A.ja.push_back(id(i, j - 1, k)); A.a.push_back(mk(-1.0)); }
    if (i > 0) { A.ja.push_back(id(i - 1, j, k)); A.a.push_back(mk(-1.0)); }
    A.ja.push_back(id(i, j, k)); A.a.push_back(mk(6.0));
    if (i < nx - 1) { A.ja.push_back(id(i + 1, j, k)); A.a.push_back(mk(-1.0)); }
    if (j < ny - 1) { A.ja.push_back(id(i, j + 1, k)); A.a.push_back(mk(-1.0)); }
    if (k < nz - 1) { A.ja.push_back(id(i, j, k + 1)); A.a.push_back(mk(-1.0)); }
    A.ia.push_back((int)A.ja.size());
  }
  Symbolic S;
  if (symbolic_analyze(A, hint ? nx : 0, ny, nz, 1, leaf, S) < 0) { printf("symbolic failed\n"); return 2; }
  const int F = (int)S.fronts.size();
  // permutation validity
  std::vector<int> seen(n, 0);
  for (int p : S.perm) seen[p]++;
  for (int v : seen) if (v != 1) { printf("perm invalid\n"); return 3; }
  // level property
  for (int f = 0; f < F; ++f) if (S.fronts[f].parent >= 0 && S.fronts[S.fronts[f].parent].level != S.fronts[f].level + 1) { printf("level property violated at front %d\n", f); return 4; }
  // dense multifrontal
  // ---- work-item tiling: every entry of every panel (pivot trapezoid + update rows) is covered by exactly one forward item,
  // every entry on or below the diagonal by exactly one backward item
  {
    std::vector<std::vector<unsigned char>> cf(F), cb(F);
    for (int f = 0; f < F; ++f) { const Front &fr = S.fronts[f]; cf[f].assign((size_t)(fr.s1 + fr.s2) * std::max(fr.s1, 1), 0); cb[f] = cf[f]; }
    for (const FwdItem &w : S.fwd) {
      const Front &fr = S.fronts[w.front];
      const int nb1 = (fr.s1 + RB - 1) / RB;
      int r0, nrows, cmax;
      if (w.rblk < nb1) { r0 = RB * w.rblk; nrows = std::min(RB, fr.s1 - r0); cmax = std::min(fr.s1, RB * (w.rblk + 1)); }
      else { r0 = fr.s1 + RB * (w.rblk - nb1); nrows = std::min(RB, fr.s1 + fr.s2 - r0); cmax = fr.s1; }
      if (VE == 2 && (w.c0 & 1)) { printf("odd forward chunk start (128-bit loads need even columns)\n"); return 8; }
      if (w.cw > FCH) { printf("forward item wider than FCH\n"); return 8; }
      for (int r = r0; r < r0 + nrows; ++r) for (int c = w.c0; c < std::min(cmax, w.c0 + w.cw); ++c) cf[w.front][(size_t)r * fr.s1 + c]++;
    }
    for (const BwdItem &w : S.bwd) {
      const Front &fr = S.fronts[w.front];
      if (w.nr > BROWS || w.c0 % BCH != 0) { printf("bad backward item\n"); return 9; }
      for (int r = w.r0; r < w.r0 + w.nr; ++r) for (int c = w.c0; c < std::min(fr.s1, w.c0 + BCH); ++c) cb[w.front][(size_t)r * fr.s1 + c]++;
    }
    for (int f = 0; f < F; ++f) {
      const Front &fr = S.fronts[f];
      for (int r = 0; r < fr.s1 + fr.s2; ++r) for (int c = 0; c < fr.s1; ++c) {
        const bool in_block = r >= fr.s1 || c < std::min(fr.s1, RB * (r / RB + 1));   // stored part of the trapezoid
        const bool nonzero = r >= fr.s1 || c <= r;
        const int nf = cf[f][(size_t)r * fr.s1 + c], nb = cb[f][(size_t)r * fr.s1 + c];
        if (nf != (in_block ? 1 : 0)) { printf("forward tiling: front %d entry (%d,%d) covered %d times\n", f, r, c, nf); return 10; }
        if (nonzero && nb != 1) { printf("backward tiling: front %d entry (%d,%d) covered %d times\n", f, r, c, nb); return 11; }
        if (nb > 1) { printf("backward tiling: front %d entry (%d,%d) covered twice\n", f, r, c); return 11; }
      }
    }
    // levels: items sorted by level, item ranges consistent
    if ((int64_t)S.fwd.size() != S.fwd_ptr[S.nlevels] || (int64_t)S.bwd.size() != S.bwd_ptr[S.nlevels]) { printf("item pointers inconsistent\n"); return 12; }
    for (int l = 0; l < S.nlevels; ++l) for (int64_t q = S.fwd_ptr[l]; q < S.fwd_ptr[l + 1]; ++q) if (S.fronts[S.fwd[q].front].level != l) { printf("forward item in the wrong level\n"); return 12; }
    for (int l = 0; l < S.nlevels; ++l) for (int64_t q = S.bwd_ptr[l]; q < S.bwd_ptr[l + 1]; ++q) if (S.fronts[S.bwd[q].front].level != l) { printf("backward item in the wrong level\n"); return 12; }
  }
#ifndef HB_COMPLEX
  std::vector<std::vector<double>> Fm(F);
  auto loc = [&](int f, int p) { const Front &fr = S.fronts[f]; if (p < fr.p0 + fr.s1) return p - fr.p0; const int *b = S.rowidx.data() + fr.rptr; const int *q = std::lower_bound(b, b + fr.s2, p); if (q == b + fr.s2 || *q != p) { printf("entry not in struct: front %d p %d\n", f, p); exit(5); } return fr.s1 + (int)(q - b); };
  for (int f = 0; f < F; ++f) { int s = S.fronts[f].s1 + S.fronts[f].s2; Fm[f].assign((size_t)s * s, 0.0); }
  for (int i = 0; i < n; ++i) { int pi = S.iperm[i]; for (int k = A.ia[i]; k < A.ia[i + 1]; ++k) { int pj = S.iperm[A.ja[k]]; if (pi < pj) continue; int f = S.front_of[pj]; int s = S.fronts[f].s1 + S.fronts[f].s2; Fm[f][loc(f, pi) + (size_t)(pj - S.fronts[f].p0) * s] += A.a[k]; } }
  std::vector<double> b(n), x(n);
  for (int i = 0; i < n; ++i) b[i] = 1.0 + 0.01 * (i % 17);
  std::vector<double> bp(n);
  for (int i = 0; i < n; ++i) bp[i] = b[S.perm[i]];
  for (int f = 0; f < F; ++f) {
    const Front &fr = S.fronts[f]; int s1 = fr.s1, s2 = fr.s2, s = s1 + s2; double *M = Fm[f].data();
    for (int c : S.children[f]) { const Front &cf = S.fronts[c]; int sc = cf.s1 + cf.s2; const int *r = S.rel.data() + cf.rptr; for (int j = 0; j < cf.s2; ++j) for (int i = j; i < cf.s2; ++i) { if (r[i] < r[j] || r[i] >= s) { printf("bad rel front %d child %d\n", f, c); return 6; } M[r[i] + (size_t)r[j] * s] += Fm[c][cf.s1 + i + (size_t)(cf.s1 + j) * sc]; } }
    for (int c : S.children[f]) { std::vector<double>().swap(Fm[c]); }
    for (int k = 0; k < s1; ++k) { double piv = M[k + (size_t)k * s]; if (!(piv > 0)) { printf("pivot breakdown front %d (s1 %d s2 %d level %d) k %d piv %g\n", f, s1, s2, fr.level, k, piv); return 7; } double sq = sqrt(piv); for (int i = k; i < s; ++i) M[i + (size_t)k * s] /= sq; for (int j = k + 1; j < s; ++j) { double l = M[j + (size_t)k * s]; if (l != 0) for (int i = j; i < s; ++i) M[i + (size_t)j * s] -= M[i + (size_t)k * s] * l; } }
  }
  // children freed their fronts: keep L parts by recomputation-free approach -> redo storing L separately
#endif
  printf("factor ok: fronts %d levels %d nnz_factor %lld panel_elems %lld fwd items %zu bwd items %zu (FCH %d BCH %d)\n", F, S.nlevels, (long long)S.nnz_factor,
         (long long)S.panel_elems, S.fwd.size(), S.bwd.size(), FCH, BCH);
  return 0;
}
