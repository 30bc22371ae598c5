// Host-side analysis: fill-reducing ordering (geometric or algebraic nested
// dissection), supernodal symbolic factorisation, level schedule and the
// work-item lists consumed by the SpTRSV kernels.
//
// The reference delegates all of this to its third-party SUBDOMAIN solver
// (MUMPS / CHOLMOD / UMFPACK analysis phase, include/HPDDM_MUMPS.hpp:228-291,
// include/HPDDM_SuiteSparse.hpp:264-371); nothing here is derived from HPDDM
// sources.
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <numeric>
#include <queue>

#include "hb_internal.h"

namespace hb {

static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char *get_error() { return g_err; }

namespace {

struct Builder {
  std::vector<int> order;      // order[new] = old
  std::vector<int> front_end;  // exclusive end (in new numbering) of each front
  void emit(const std::vector<int> &v) {
    if (v.empty()) return;
    order.insert(order.end(), v.begin(), v.end());
    front_end.push_back((int)order.size());
  }
};

// ---- geometric nested dissection on an nx x ny x nz grid with dof unknowns/node
struct Box {
  int lo[3], hi[3];
};
static void geo_nd(const Box &b, const int dims[3], int dof, int leaf, Builder &B) {
  // explicit stack, post-order: children first, separator last
  struct Task {
    Box b;
    int stage;  // 0 = expand, 1 = emit separator
    int axis, mid;
  };
  std::vector<Task> st;
  st.push_back({b, 0, 0, 0});
  std::vector<int> tmp;
  auto emit_box = [&](const Box &x) {
    tmp.clear();
    for (int k = x.lo[2]; k < x.hi[2]; ++k)
      for (int j = x.lo[1]; j < x.hi[1]; ++j)
        for (int i = x.lo[0]; i < x.hi[0]; ++i) {
          int node = (k * dims[1] + j) * dims[0] + i;
          for (int c = 0; c < dof; ++c) tmp.push_back(node * dof + c);
        }
    B.emit(tmp);
  };
  while (!st.empty()) {
    Task t = st.back();
    st.pop_back();
    int ext[3] = {t.b.hi[0] - t.b.lo[0], t.b.hi[1] - t.b.lo[1], t.b.hi[2] - t.b.lo[2]};
    if (ext[0] <= 0 || ext[1] <= 0 || ext[2] <= 0) continue;
    if (t.stage == 1) {
      Box s = t.b;
      s.lo[t.axis] = t.mid;
      s.hi[t.axis] = t.mid + 1;
      emit_box(s);
      continue;
    }
    int64_t vol = (int64_t)ext[0] * ext[1] * ext[2] * dof;
    int ax = 0;
    if (ext[1] > ext[ax]) ax = 1;
    if (ext[2] > ext[ax]) ax = 2;
    if (vol <= leaf || ext[ax] < 3) {
      emit_box(t.b);
      continue;
    }
    int mid = t.b.lo[ax] + ext[ax] / 2;
    Box l = t.b, r = t.b;
    l.hi[ax] = mid;
    r.lo[ax] = mid + 1;
    st.push_back({t.b, 1, ax, mid});  // separator emitted after both halves
    st.push_back({r, 0, 0, 0});
    st.push_back({l, 0, 0, 0});
  }
}

// ---- algebraic nested dissection: BFS level structures from a pseudo-peripheral vertex
static void alg_nd(const HostCSR &A, const std::vector<std::vector<int>> &adjT, int leaf, Builder &B) {
  const int n = A.n;
  std::vector<int> part(n, 0);  // current subset id of each vertex
  std::vector<int> dist(n, -1);
  struct Task {
    std::vector<int> verts;
    int stage;
    std::vector<int> sep;
  };
  std::vector<Task> st;
  {
    Task t;
    t.verts.resize(n);
    std::iota(t.verts.begin(), t.verts.end(), 0);
    t.stage = 0;
    st.push_back(std::move(t));
  }
  int next_id = 1;
  auto neighbours = [&](int v, auto &&fn) {
    for (int k = A.ia[v]; k < A.ia[v + 1]; ++k) fn(A.ja[k]);
    for (int u : adjT[v]) fn(u);
  };
  std::vector<int> q;
  while (!st.empty()) {
    Task t = std::move(st.back());
    st.pop_back();
    if (t.stage == 1) {
      B.emit(t.sep);
      continue;
    }
    if ((int)t.verts.size() <= leaf) {
      B.emit(t.verts);
      continue;
    }
    const int id = next_id++;
    for (int v : t.verts) part[v] = id;
    // BFS helper restricted to `id`
    auto bfs = [&](int root) {
      for (int v : t.verts) dist[v] = -1;
      q.clear();
      q.push_back(root);
      dist[root] = 0;
      for (size_t h = 0; h < q.size(); ++h) {
        int v = q[h];
        neighbours(v, [&](int u) {
          if (part[u] == id && dist[u] < 0) {
            dist[u] = dist[v] + 1;
            q.push_back(u);
          }
        });
      }
      return q.back();
    };
    int root = t.verts[0];
    int far = bfs(root);
    for (int it = 0; it < 2; ++it) {  // pseudo-peripheral refinement
      int d0 = dist[far];
      int far2 = bfs(far);
      if (dist[far2] <= d0) break;
      far = far2;
    }
    bfs(far);
    int reached = (int)q.size();
    int maxd = dist[q.back()];
    std::vector<int> A1, A2, sep;
    if (reached < (int)t.verts.size()) {
      // disconnected: component vs rest, no separator
      for (int v : t.verts) (dist[v] >= 0 ? A1 : A2).push_back(v);
    } else if (maxd < 2) {
      B.emit(t.verts);  // clique-like
      continue;
    } else {
      std::vector<int> cnt(maxd + 1, 0);
      for (int v : t.verts) cnt[dist[v]]++;
      int half = (int)t.verts.size() / 2, acc = 0, cut = 1;
      for (int l = 0; l <= maxd; ++l) {
        if (acc + cnt[l] >= half) {
          cut = l;
          break;
        }
        acc += cnt[l];
      }
      cut = std::max(1, std::min(cut, maxd - 1));
      for (int v : t.verts) {
        if (dist[v] < cut) A1.push_back(v);
        else if (dist[v] == cut) sep.push_back(v);
        else A2.push_back(v);
      }
    }
    Task ts;
    ts.stage = 1;
    ts.sep = std::move(sep);
    st.push_back(std::move(ts));
    Task t2;
    t2.stage = 0;
    t2.verts = std::move(A2);
    st.push_back(std::move(t2));
    Task t1;
    t1.stage = 0;
    t1.verts = std::move(A1);
    st.push_back(std::move(t1));
  }
}

}  // namespace

int symbolic_analyze(const HostCSR &A, int nx, int ny, int nz, int dof, int leaf, Symbolic &S) {
  const int n = A.n;
  S = Symbolic();
  S.n = n;
  Builder B;
  B.order.reserve(n);
  // transposed adjacency (pattern of A^T) so that the graph of A + A^T is used
  std::vector<std::vector<int>> adjT;
  bool need_T = !A.symmetric;
  if (need_T) {
    adjT.assign(n, {});
    for (int i = 0; i < n; ++i)
      for (int k = A.ia[i]; k < A.ia[i + 1]; ++k)
        if (A.ja[k] != i) adjT[A.ja[k]].push_back(i);
  } else
    adjT.assign(n, {});
  if (nx > 0 && (int64_t)nx * ny * nz * dof == n) {
    Box b{{0, 0, 0}, {nx, ny, nz}};
    int dims[3] = {nx, ny, nz};
    geo_nd(b, dims, dof, leaf, B);
  } else {
    alg_nd(A, adjT, leaf, B);
  }
  if ((int)B.order.size() != n) {
    set_error("symbolic: ordering covers %d of %d dofs", (int)B.order.size(), n);
    return HPDDM_B200_ERR_STATE;
  }
  S.perm = B.order;
  S.iperm.assign(n, -1);
  for (int i = 0; i < n; ++i) S.iperm[S.perm[i]] = i;
  const int F = (int)B.front_end.size();
  S.fronts.resize(F);
  S.front_of.resize(n);
  {
    int p = 0;
    for (int f = 0; f < F; ++f) {
      S.fronts[f].p0 = p;
      S.fronts[f].s1 = B.front_end[f] - p;
      S.fronts[f].parent = -1;
      for (; p < B.front_end[f]; ++p) S.front_of[p] = f;
    }
  }
  // ---- supernodal structure: struct(f) = (adj(cols f) U struct(children)) \ {<= last col of f}
  S.children.assign(F, {});
  std::vector<int> mark(n, -1);
  std::vector<int> cur;
  std::vector<std::vector<int>> structs(F);
  for (int f = 0; f < F; ++f) {
    Front &fr = S.fronts[f];
    const int last = fr.p0 + fr.s1 - 1;
    cur.clear();
    auto add = [&](int p) {
      if (p > last && mark[p] != f) {
        mark[p] = f;
        cur.push_back(p);
      }
    };
    for (int p = fr.p0; p <= last; ++p) {
      int v = S.perm[p];
      for (int k = A.ia[v]; k < A.ia[v + 1]; ++k) add(S.iperm[A.ja[k]]);
      if (need_T)
        for (int u : adjT[v]) add(S.iperm[u]);
    }
    for (int c : S.children[f])
      for (int p : structs[c]) add(p);
    std::sort(cur.begin(), cur.end());
    structs[f] = cur;
    fr.s2 = (int)cur.size();
    if (!cur.empty()) {
      fr.parent = S.front_of[cur[0]];
      S.children[fr.parent].push_back(f);
    }
    // children's structs are no longer needed once merged into the parent: keep (needed for rel)
  }
  // ---- flatten struct, relative indices
  int64_t tot = 0;
  for (int f = 0; f < F; ++f) {
    S.fronts[f].rptr = tot;
    tot += S.fronts[f].s2;
  }
  if (tot > (int64_t)2000000000) {
    set_error("symbolic: structure too large (%lld)", (long long)tot);
    return HPDDM_B200_ERR_NOMEM;
  }
  S.rowidx.resize(tot);
  S.rel.resize(tot);
  for (int f = 0; f < F; ++f) {
    const Front &fr = S.fronts[f];
    std::copy(structs[f].begin(), structs[f].end(), S.rowidx.begin() + fr.rptr);
    if (fr.parent >= 0) {
      const Front &pf = S.fronts[fr.parent];
      const std::vector<int> &ps = structs[fr.parent];
      for (int i = 0; i < fr.s2; ++i) {
        int p = structs[f][i];
        int r;
        if (p < pf.p0 + pf.s1) r = p - pf.p0;
        else r = pf.s1 + (int)(std::lower_bound(ps.begin(), ps.end(), p) - ps.begin());
        S.rel[fr.rptr + i] = r;
      }
    }
  }
  // ---- levels: depth from the roots, level = maxdepth - depth
  std::vector<int> depth(F, 0);
  int maxd = 0;
  for (int f = F - 1; f >= 0; --f) {
    depth[f] = S.fronts[f].parent < 0 ? 0 : depth[S.fronts[f].parent] + 1;
    maxd = std::max(maxd, depth[f]);
  }
  S.nlevels = maxd + 1;
  S.level_ptr.assign(S.nlevels + 1, 0);
  for (int f = 0; f < F; ++f) {
    S.fronts[f].level = maxd - depth[f];
    S.level_ptr[S.fronts[f].level + 1]++;
  }
  for (int l = 0; l < S.nlevels; ++l) S.level_ptr[l + 1] += S.level_ptr[l];
  S.level_order.resize(F);
  {
    std::vector<int> pos(S.level_ptr.begin(), S.level_ptr.end() - 1);
    for (int f = 0; f < F; ++f) S.level_order[pos[S.fronts[f].level]++] = f;
  }
  // ---- panel offsets (level-major so that a level's panels are contiguous in HBM)
  int64_t off = 0, nnzf = 0;
  for (int q = 0; q < F; ++q) {
    Front &fr = S.fronts[S.level_order[q]];
    fr.poff = off;
    off += hb_panel_size(fr.s1, fr.s2);
    off = (off + 15) & ~(int64_t)15;  // 128-byte aligned panels
    nnzf += (int64_t)fr.s1 * (fr.s1 + 1) / 2 + (int64_t)fr.s1 * fr.s2;
  }
  S.panel_elems = off;
  S.nnz_factor = nnzf;
  // ---- work items.  Default granularity: forward 32 rows x FCH columns, backward BROWS rows x BCH columns per warp.
  // Levels whose panels would yield fewer items than the GPU has warp slots (the few huge fronts near the
  // root of small problems) get proportionally finer items so that all SMs stream.
  const int64_t target_items = 148 * 4 * 8;  // SMs x resident CTAs x warps
  S.fwd_ptr.assign(S.nlevels + 1, 0);
  S.bwd_ptr.assign(S.nlevels + 1, 0);
  for (int l = 0; l < S.nlevels; ++l) {
    auto count = [&](int fch, int brows, int64_t &nf, int64_t &nbw) {
      nf = nbw = 0;
      for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
        const Front &fr = S.fronts[S.level_order[q]];
        const int nb1 = (fr.s1 + RB - 1) / RB, nb2 = (fr.s2 + RB - 1) / RB;
        for (int k = 0; k < nb1; ++k) nf += (std::min(fr.s1, RB * (k + 1)) + fch - 1) / fch;
        nf += (int64_t)nb2 * ((fr.s1 + fch - 1) / fch);
        for (int c0 = 0; c0 < fr.s1; c0 += BCH) nbw += (fr.s1 + fr.s2 - (c0 / RB) * RB + brows - 1) / brows;
      }
    };
    int fch = FCH, brows = BROWS;
    int64_t nf, nbw;
    count(fch, brows, nf, nbw);
    while (nf < target_items && fch > 128) {
      fch /= 2;
      count(fch, brows, nf, nbw);
    }
    while (nbw < target_items && brows > 32) {
      brows /= 2;
      count(fch, brows, nf, nbw);
    }
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
      int f = S.level_order[q];
      const Front &fr = S.fronts[f];
      const int nb1 = (fr.s1 + RB - 1) / RB, nb2 = (fr.s2 + RB - 1) / RB;
      for (int k = 0; k < nb1; ++k) {
        int w = std::min(fr.s1, RB * (k + 1));
        for (int c0 = 0; c0 < w; c0 += fch) S.fwd.push_back({f, k, c0, fch});
      }
      for (int k = 0; k < nb2; ++k)
        for (int c0 = 0; c0 < fr.s1; c0 += fch) S.fwd.push_back({f, nb1 + k, c0, fch});
      for (int c0 = 0; c0 < fr.s1; c0 += BCH) {
        int rstart = (c0 / RB) * RB;  // rows above hold structural zeros in this chunk
        for (int r0 = rstart; r0 < fr.s1 + fr.s2; r0 += brows) S.bwd.push_back({f, c0, r0, std::min(brows, fr.s1 + fr.s2 - r0)});
      }
    }
    S.fwd_ptr[l + 1] = (int64_t)S.fwd.size();
    S.bwd_ptr[l + 1] = (int64_t)S.bwd.size();
  }
  return 0;
}

}  // namespace hb
