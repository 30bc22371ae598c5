// In-box MPI shim: ranks = forked processes, messages through a shared anonymous mapping.
// See mpi.h.  TEST INFRASTRUCTURE ONLY.
#include "mpi.h"

#include <pthread.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Msg {
  int src, tag;
  long ctx;
  size_t bytes;
  Msg *next;
  // payload follows
};
struct Box {
  pthread_mutex_t mu;
  pthread_cond_t cv;
  Msg *head, *tail;
};
struct Shared {
  pthread_mutex_t alloc_mu;
  size_t used, cap;
  int np;
  Box box[256];
};
Shared *g_sh = nullptr;
char *g_arena = nullptr;
int g_rank = 0, g_np = 1;
bool g_init = false, g_final = false;
std::vector<pid_t> g_children;

struct Comm {
  long ctx;
  std::vector<int> ranks;  // world ranks
  int me;                  // my index, -1 if not a member
  long counter = 0;        // children created from this communicator
  bool alive = true;
};
std::vector<Comm> g_comms;  // index = handle
std::vector<std::vector<int>> g_groups;
struct Req {
  int kind = 0;  // 0 free, 1 send (complete), 2 recv pending, 3 recv done
  void *buf = nullptr;
  size_t bytes = 0;
  int src = 0, tag = 0;
  long ctx = 0;
  MPI_Status st{};
};
std::vector<Req> g_reqs(1);
std::vector<MPI_User_function *> g_ops;

size_t tsize(MPI_Datatype t) {
  switch (t) {
    case MPI_CHAR: case MPI_UNSIGNED_CHAR: case MPI_BYTE: case MPI_C_BOOL: return 1;
    case MPI_SHORT: case MPI_UNSIGNED_SHORT: return 2;
    case MPI_INT: case MPI_UNSIGNED: case MPI_FLOAT: return 4;
    case MPI_LONG: case MPI_UNSIGNED_LONG: case MPI_LONG_LONG: case MPI_UNSIGNED_LONG_LONG: case MPI_DOUBLE: case MPI_C_COMPLEX: return 8;
    case MPI_C_DOUBLE_COMPLEX: case MPI_LONG_DOUBLE: return 16;
  }
  fprintf(stderr, "mpi shim: unknown datatype %d\n", t);
  abort();
}

void *arena_alloc(size_t n) {
  n = (n + 63) & ~(size_t)63;
  pthread_mutex_lock(&g_sh->alloc_mu);
  if (g_sh->used + n > g_sh->cap) {
    fprintf(stderr, "mpi shim: arena exhausted\n");
    abort();
  }
  void *p = g_arena + g_sh->used;
  g_sh->used += n;
  pthread_mutex_unlock(&g_sh->alloc_mu);
  return p;
}

void post(int dst_world, long ctx, int src_rank_in_comm, int tag, const void *data, size_t bytes) {
  Msg *m = (Msg *)arena_alloc(sizeof(Msg) + bytes);
  m->src = src_rank_in_comm;
  m->tag = tag;
  m->ctx = ctx;
  m->bytes = bytes;
  m->next = nullptr;
  if (bytes) memcpy(m + 1, data, bytes);
  Box &b = g_sh->box[dst_world];
  pthread_mutex_lock(&b.mu);
  if (b.tail) b.tail->next = m;
  else b.head = m;
  b.tail = m;
  pthread_cond_broadcast(&b.cv);
  pthread_mutex_unlock(&b.mu);
}

// caller holds my box lock; tries to complete a pending recv request
bool try_match(Req &r) {
  Box &b = g_sh->box[g_rank];
  Msg *prev = nullptr;
  for (Msg *m = b.head; m; prev = m, m = m->next) {
    if (m->ctx != r.ctx) continue;
    if (r.src != MPI_ANY_SOURCE && m->src != r.src) continue;
    if (r.tag != MPI_ANY_TAG && m->tag != r.tag) continue;
    if (m->bytes > r.bytes) {
      fprintf(stderr, "mpi shim: message truncated (%zu > %zu)\n", m->bytes, r.bytes);
      abort();
    }
    if (m->bytes) memcpy(r.buf, m + 1, m->bytes);
    r.st.MPI_SOURCE = m->src;
    r.st.MPI_TAG = m->tag;
    r.st.bytes_ = m->bytes;
    if (prev) prev->next = m->next;
    else b.head = m->next;
    if (b.tail == m) b.tail = prev;
    r.kind = 3;
    return true;
  }
  return false;
}

void wait_one(Req &r) {
  if (r.kind != 2) return;
  Box &b = g_sh->box[g_rank];
  pthread_mutex_lock(&b.mu);
  while (!try_match(r)) pthread_cond_wait(&b.cv, &b.mu);
  pthread_mutex_unlock(&b.mu);
}

int new_req() {
  for (size_t i = 1; i < g_reqs.size(); ++i)
    if (g_reqs[i].kind == 0) return (int)i;
  g_reqs.push_back(Req());
  return (int)g_reqs.size() - 1;
}

Comm &C(MPI_Comm c) {
  if (c <= 0 || c >= (int)g_comms.size() || !g_comms[c].alive) {
    fprintf(stderr, "mpi shim: invalid communicator %d\n", c);
    abort();
  }
  return g_comms[c];
}

long child_ctx(Comm &parent) {
  ++parent.counter;
  return parent.ctx * 1000003L + parent.counter * 7919L + 17;
}

const int TAG_COLL = -1000;  // internal collective tags are negative

void send_raw(Comm &c, int dst, int tag, const void *buf, size_t bytes) { post(c.ranks[dst], c.ctx, c.me, tag, buf, bytes); }
void recv_raw(Comm &c, int src, int tag, void *buf, size_t bytes, MPI_Status *st = nullptr) {
  Req r;
  r.kind = 2;
  r.buf = buf;
  r.bytes = bytes;
  r.src = src;
  r.tag = tag;
  r.ctx = c.ctx;
  wait_one(r);
  if (st) *st = r.st;
}

template <class T>
void reduce_t(const T *in, T *io, int n, MPI_Op op) {
  for (int i = 0; i < n; ++i) {
    switch (op) {
      case MPI_SUM: io[i] = io[i] + in[i]; break;
      case MPI_PROD: io[i] = io[i] * in[i]; break;
      case MPI_MAX: io[i] = std::max(io[i], in[i]); break;
      case MPI_MIN: io[i] = std::min(io[i], in[i]); break;
      case MPI_LOR: io[i] = (T)(io[i] || in[i]); break;
      case MPI_LAND: io[i] = (T)(io[i] && in[i]); break;
      default: fprintf(stderr, "mpi shim: unsupported op %d\n", op); abort();
    }
  }
}
template <class T>
void reduce_c(const std::complex<T> *in, std::complex<T> *io, int n, MPI_Op op) {
  if (op != MPI_SUM) { fprintf(stderr, "mpi shim: complex op %d\n", op); abort(); }
  for (int i = 0; i < n; ++i) io[i] += in[i];
}
void apply_op(const void *in, void *io, int n, MPI_Datatype t, MPI_Op op) {
  if (op >= MPI_OP_USER_BASE) {
    g_ops[op - MPI_OP_USER_BASE](const_cast<void *>(in), io, &n, &t);
    return;
  }
  switch (t) {
    case MPI_CHAR: reduce_t((const char *)in, (char *)io, n, op); break;
    case MPI_UNSIGNED_CHAR: case MPI_BYTE: case MPI_C_BOOL: reduce_t((const unsigned char *)in, (unsigned char *)io, n, op); break;
    case MPI_SHORT: reduce_t((const short *)in, (short *)io, n, op); break;
    case MPI_UNSIGNED_SHORT: reduce_t((const unsigned short *)in, (unsigned short *)io, n, op); break;
    case MPI_INT: reduce_t((const int *)in, (int *)io, n, op); break;
    case MPI_UNSIGNED: reduce_t((const unsigned *)in, (unsigned *)io, n, op); break;
    case MPI_LONG: case MPI_LONG_LONG: reduce_t((const long long *)in, (long long *)io, n, op); break;
    case MPI_UNSIGNED_LONG: case MPI_UNSIGNED_LONG_LONG: reduce_t((const unsigned long long *)in, (unsigned long long *)io, n, op); break;
    case MPI_FLOAT: reduce_t((const float *)in, (float *)io, n, op); break;
    case MPI_DOUBLE: reduce_t((const double *)in, (double *)io, n, op); break;
    case MPI_C_COMPLEX: reduce_c((const std::complex<float> *)in, (std::complex<float> *)io, n, op); break;
    case MPI_C_DOUBLE_COMPLEX: reduce_c((const std::complex<double> *)in, (std::complex<double> *)io, n, op); break;
    default: fprintf(stderr, "mpi shim: reduce on datatype %d\n", t); abort();
  }
}

}  // namespace

extern "C" {

int MPI_Init(int *, char ***) {
  if (g_init) return 0;
  const char *e = getenv("HPDDM_SHIM_NP");
  g_np = e ? atoi(e) : 1;
  if (g_np < 1 || g_np > 256) g_np = 1;
  size_t cap = (size_t)(getenv("HPDDM_SHIM_ARENA_GB") ? atof(getenv("HPDDM_SHIM_ARENA_GB")) : 8.0) * (1ull << 30);
  void *p = mmap(nullptr, sizeof(Shared) + cap, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) { perror("mpi shim mmap"); abort(); }
  g_sh = (Shared *)p;
  g_arena = (char *)p + ((sizeof(Shared) + 63) & ~(size_t)63);
  pthread_mutexattr_t ma;
  pthread_mutexattr_init(&ma);
  pthread_mutexattr_setpshared(&ma, PTHREAD_PROCESS_SHARED);
  pthread_condattr_t ca;
  pthread_condattr_init(&ca);
  pthread_condattr_setpshared(&ca, PTHREAD_PROCESS_SHARED);
  pthread_mutex_init(&g_sh->alloc_mu, &ma);
  g_sh->used = 0;
  g_sh->cap = cap - 4096;
  g_sh->np = g_np;
  for (int i = 0; i < g_np; ++i) {
    pthread_mutex_init(&g_sh->box[i].mu, &ma);
    pthread_cond_init(&g_sh->box[i].cv, &ca);
    g_sh->box[i].head = g_sh->box[i].tail = nullptr;
  }
  fflush(stdout);
  fflush(stderr);
  g_rank = 0;
  for (int r = 1; r < g_np; ++r) {
    pid_t pid = fork();
    if (pid == 0) {
      g_rank = r;
      g_children.clear();
      break;
    }
    g_children.push_back(pid);
  }
  g_comms.resize(3);
  g_comms[0].alive = false;
  g_comms[1].ctx = 1;
  g_comms[1].ranks.resize(g_np);
  for (int i = 0; i < g_np; ++i) g_comms[1].ranks[i] = i;
  g_comms[1].me = g_rank;
  g_comms[2].ctx = 2 + 1000L * (g_rank + 1);
  g_comms[2].ranks = {g_rank};
  g_comms[2].me = 0;
  g_groups.resize(1);
  g_init = true;
  return 0;
}
int MPI_Initialized(int *f) { *f = g_init; return 0; }
int MPI_Finalized(int *f) { *f = g_final; return 0; }
int MPI_Finalize(void) {
  MPI_Barrier(MPI_COMM_WORLD);
  g_final = true;
  fflush(stdout);
  fflush(stderr);
  if (g_rank != 0) _exit(0);
  int bad = 0;
  for (pid_t p : g_children) {
    int st = 0;
    waitpid(p, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad = 1;
  }
  if (bad) fprintf(stderr, "mpi shim: a rank failed\n");
  return 0;
}
int MPI_Abort(MPI_Comm, int code) { _exit(code ? code : 1); }
double MPI_Wtime(void) {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
int MPI_Comm_rank(MPI_Comm c, int *r) { *r = C(c).me; return 0; }
int MPI_Comm_size(MPI_Comm c, int *s) { *s = (int)C(c).ranks.size(); return 0; }
int MPI_Comm_dup(MPI_Comm c, MPI_Comm *out) {
  Comm n = C(c);
  n.ctx = child_ctx(C(c));
  n.counter = 0;
  g_comms.push_back(n);
  *out = (int)g_comms.size() - 1;
  return 0;
}
int MPI_Comm_free(MPI_Comm *c) {
  if (*c > 2 && *c < (int)g_comms.size()) g_comms[*c].alive = false;
  *c = MPI_COMM_NULL;
  return 0;
}
int MPI_Comm_group(MPI_Comm c, MPI_Group *g) {
  g_groups.push_back(C(c).ranks);
  *g = (int)g_groups.size() - 1;
  return 0;
}
int MPI_Group_size(MPI_Group g, int *s) { *s = (int)g_groups[g].size(); return 0; }
int MPI_Group_incl(MPI_Group g, int n, const int *r, MPI_Group *out) {
  std::vector<int> v(n);
  for (int i = 0; i < n; ++i) v[i] = g_groups[g][r[i]];
  g_groups.push_back(v);
  *out = (int)g_groups.size() - 1;
  return 0;
}
int MPI_Group_excl(MPI_Group g, int n, const int *r, MPI_Group *out) {
  std::vector<int> v;
  for (int i = 0; i < (int)g_groups[g].size(); ++i)
    if (std::find(r, r + n, i) == r + n) v.push_back(g_groups[g][i]);
  g_groups.push_back(v);
  *out = (int)g_groups.size() - 1;
  return 0;
}
int MPI_Group_free(MPI_Group *g) { *g = MPI_GROUP_NULL; return 0; }
int MPI_Comm_create(MPI_Comm c, MPI_Group g, MPI_Comm *out) {
  Comm &p = C(c);
  long ctx = child_ctx(p);
  const std::vector<int> &rk = g_groups[g];
  auto it = std::find(rk.begin(), rk.end(), g_rank);
  if (g == MPI_GROUP_NULL || it == rk.end()) {
    *out = MPI_COMM_NULL;
    return 0;
  }
  Comm n;
  n.ctx = ctx;
  n.ranks = rk;
  n.me = (int)(it - rk.begin());
  g_comms.push_back(n);
  *out = (int)g_comms.size() - 1;
  return 0;
}
int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *out) {
  Comm &p = C(c);
  int n = (int)p.ranks.size();
  std::vector<int> all(2 * n);
  int mine[2] = {color, key};
  MPI_Allgather(mine, 2, MPI_INT, all.data(), 2, MPI_INT, c);
  long ctx = child_ctx(p);
  if (color == MPI_UNDEFINED) { *out = MPI_COMM_NULL; return 0; }
  std::vector<std::pair<int, int>> mem;
  for (int i = 0; i < n; ++i)
    if (all[2 * i] == color) mem.push_back({all[2 * i + 1], i});
  std::stable_sort(mem.begin(), mem.end());
  Comm nc;
  nc.ctx = ctx * 31 + color;
  for (auto &m : mem) nc.ranks.push_back(p.ranks[m.second]);
  nc.me = (int)(std::find(nc.ranks.begin(), nc.ranks.end(), g_rank) - nc.ranks.begin());
  g_comms.push_back(nc);
  *out = (int)g_comms.size() - 1;
  return 0;
}
int MPI_Comm_compare(MPI_Comm a, MPI_Comm b, int *res) {
  if (a == b) { *res = MPI_IDENT; return 0; }
  if (a == MPI_COMM_NULL || b == MPI_COMM_NULL) { *res = MPI_UNEQUAL; return 0; }
  Comm &x = C(a), &y = C(b);
  if (x.ctx == y.ctx) *res = MPI_IDENT;
  else if (x.ranks == y.ranks) *res = MPI_CONGRUENT;
  else {
    std::vector<int> p = x.ranks, q = y.ranks;
    std::sort(p.begin(), p.end());
    std::sort(q.begin(), q.end());
    *res = p == q ? MPI_SIMILAR : MPI_UNEQUAL;
  }
  return 0;
}
int MPI_Send(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c) {
  send_raw(C(c), dst, tag, buf, (size_t)n * tsize(t));
  return 0;
}
int MPI_Recv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st) {
  recv_raw(C(c), src, tag, buf, (size_t)n * tsize(t), st);
  return 0;
}
int MPI_Isend(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm c, MPI_Request *rq) {
  send_raw(C(c), dst, tag, buf, (size_t)n * tsize(t));
  int i = new_req();
  g_reqs[i] = Req();
  g_reqs[i].kind = 1;
  *rq = i;
  return 0;
}
int MPI_Irecv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *rq) {
  int i = new_req();
  Req &r = g_reqs[i];
  r = Req();
  r.kind = 2;
  r.buf = buf;
  r.bytes = (size_t)n * tsize(t);
  r.src = src;
  r.tag = tag;
  r.ctx = C(c).ctx;
  *rq = i;
  return 0;
}
int MPI_Wait(MPI_Request *rq, MPI_Status *st) {
  if (*rq == MPI_REQUEST_NULL) return 0;
  Req &r = g_reqs[*rq];
  wait_one(r);
  if (st) *st = r.st;
  r.kind = 0;
  *rq = MPI_REQUEST_NULL;
  return 0;
}
int MPI_Waitall(int n, MPI_Request *rq, MPI_Status *st) {
  for (int i = 0; i < n; ++i) MPI_Wait(rq + i, st ? st + i : nullptr);
  return 0;
}
int MPI_Waitany(int n, MPI_Request *rq, int *index, MPI_Status *st) {
  bool any = false;
  for (int i = 0; i < n; ++i) any = any || rq[i] != MPI_REQUEST_NULL;
  if (!any) { *index = MPI_UNDEFINED; return 0; }
  Box &b = g_sh->box[g_rank];
  pthread_mutex_lock(&b.mu);
  for (;;) {
    for (int i = 0; i < n; ++i) {
      if (rq[i] == MPI_REQUEST_NULL) continue;
      Req &r = g_reqs[rq[i]];
      if (r.kind == 1 || r.kind == 3 || (r.kind == 2 && try_match(r))) {
        pthread_mutex_unlock(&b.mu);
        if (st) *st = r.st;
        r.kind = 0;
        rq[i] = MPI_REQUEST_NULL;
        *index = i;
        return 0;
      }
    }
    pthread_cond_wait(&b.cv, &b.mu);
  }
}
int MPI_Test(MPI_Request *rq, int *flag, MPI_Status *st) {
  if (*rq == MPI_REQUEST_NULL) { *flag = 1; return 0; }
  Req &r = g_reqs[*rq];
  if (r.kind == 2) {
    Box &b = g_sh->box[g_rank];
    pthread_mutex_lock(&b.mu);
    try_match(r);
    pthread_mutex_unlock(&b.mu);
  }
  *flag = r.kind != 2;
  if (*flag) {
    if (st) *st = r.st;
    r.kind = 0;
    *rq = MPI_REQUEST_NULL;
  }
  return 0;
}
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *n) { *n = (int)(st->bytes_ / tsize(t)); return 0; }

int MPI_Barrier(MPI_Comm c) {
  Comm &cm = C(c);
  int n = (int)cm.ranks.size();
  char z = 0;
  if (cm.me == 0) {
    for (int i = 1; i < n; ++i) recv_raw(cm, i, TAG_COLL - 1, &z, 1);
    for (int i = 1; i < n; ++i) send_raw(cm, i, TAG_COLL - 2, &z, 1);
  } else {
    send_raw(cm, 0, TAG_COLL - 1, &z, 1);
    recv_raw(cm, 0, TAG_COLL - 2, &z, 1);
  }
  return 0;
}
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm c) {
  Comm &cm = C(c);
  size_t b = (size_t)n * tsize(t);
  if (cm.me == root) {
    for (int i = 0; i < (int)cm.ranks.size(); ++i)
      if (i != root) send_raw(cm, i, TAG_COLL - 3, buf, b);
  } else
    recv_raw(cm, root, TAG_COLL - 3, buf, b);
  return 0;
}
int MPI_Gatherv(const void *sb, int sn, MPI_Datatype st, void *rb, const int *rc, const int *displs, MPI_Datatype rt, int root, MPI_Comm c) {
  Comm &cm = C(c);
  if (cm.me == root) {
    size_t ts = tsize(rt);
    for (int i = 0; i < (int)cm.ranks.size(); ++i) {
      char *dst = (char *)rb + (size_t)displs[i] * ts;
      if (i == root) {
        if (sb != MPI_IN_PLACE) memcpy(dst, sb, (size_t)sn * tsize(st));
      } else
        recv_raw(cm, i, TAG_COLL - 4, dst, (size_t)rc[i] * ts);
    }
  } else
    send_raw(cm, root, TAG_COLL - 4, sb, (size_t)sn * tsize(st));
  return 0;
}
int MPI_Gather(const void *sb, int sn, MPI_Datatype st, void *rb, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  Comm &cm = C(c);
  int n = (int)cm.ranks.size();
  std::vector<int> rc(n, rn), dp(n);
  for (int i = 0; i < n; ++i) dp[i] = i * rn;
  return MPI_Gatherv(sb, sn, st, rb, rc.data(), dp.data(), rt, root, c);
}
int MPI_Scatterv(const void *sb, const int *sc, const int *displs, MPI_Datatype st, void *rb, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  Comm &cm = C(c);
  if (cm.me == root) {
    size_t ts = tsize(st);
    for (int i = 0; i < (int)cm.ranks.size(); ++i) {
      const char *src = (const char *)sb + (size_t)displs[i] * ts;
      if (i == root) {
        if (rb != MPI_IN_PLACE) memcpy(rb, src, (size_t)sc[i] * ts);
      } else
        send_raw(cm, i, TAG_COLL - 5, src, (size_t)sc[i] * ts);
    }
  } else
    recv_raw(cm, root, TAG_COLL - 5, rb, (size_t)rn * tsize(rt));
  return 0;
}
int MPI_Scatter(const void *sb, int sn, MPI_Datatype st, void *rb, int rn, MPI_Datatype rt, int root, MPI_Comm c) {
  Comm &cm = C(c);
  int n = (int)cm.ranks.size();
  std::vector<int> sc(n, sn), dp(n);
  for (int i = 0; i < n; ++i) dp[i] = i * sn;
  return MPI_Scatterv(sb, sc.data(), dp.data(), st, rb, rn, rt, root, c);
}
int MPI_Allgatherv(const void *sb, int sn, MPI_Datatype st, void *rb, const int *rc, const int *displs, MPI_Datatype rt, MPI_Comm c) {
  Comm &cm = C(c);
  int n = (int)cm.ranks.size();
  size_t ts = tsize(rt);
  if (sb == MPI_IN_PLACE) MPI_Gatherv(cm.me == 0 ? MPI_IN_PLACE : (char *)rb + (size_t)displs[cm.me] * ts, rc[cm.me], rt, rb, rc, displs, rt, 0, c);
  else MPI_Gatherv(sb, sn, st, rb, rc, displs, rt, 0, c);
  int tot = 0;
  for (int i = 0; i < n; ++i) tot = std::max(tot, displs[i] + rc[i]);
  return MPI_Bcast(rb, tot, rt, 0, c);
}
int MPI_Allgather(const void *sb, int sn, MPI_Datatype st, void *rb, int rn, MPI_Datatype rt, MPI_Comm c) {
  Comm &cm = C(c);
  int n = (int)cm.ranks.size();
  std::vector<int> rc(n, rn), dp(n);
  for (int i = 0; i < n; ++i) dp[i] = i * rn;
  return MPI_Allgatherv(sb, sn, st, rb, rc.data(), dp.data(), rt, c);
}
int MPI_Reduce(const void *sb, void *rb, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c) {
  Comm &cm = C(c);
  size_t b = (size_t)n * tsize(t);
  if (cm.me == root) {
    if (sb != MPI_IN_PLACE) memcpy(rb, sb, b);
    std::vector<char> tmp(b);
    for (int i = 0; i < (int)cm.ranks.size(); ++i) {
      if (i == root) continue;
      recv_raw(cm, i, TAG_COLL - 6, tmp.data(), b);
      apply_op(tmp.data(), rb, n, t, op);
    }
  } else
    send_raw(cm, root, TAG_COLL - 6, sb == MPI_IN_PLACE ? rb : sb, b);
  return 0;
}
int MPI_Allreduce(const void *sb, void *rb, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c) {
  MPI_Reduce(sb, rb, n, t, op, 0, c);
  return MPI_Bcast(rb, n, t, 0, c);
}
int MPI_Op_create(MPI_User_function *f, int, MPI_Op *op) {
  g_ops.push_back(f);
  *op = MPI_OP_USER_BASE + (int)g_ops.size() - 1;
  return 0;
}
int MPI_Op_free(MPI_Op *op) { *op = 0; return 0; }
int MPI_Igather(const void *a, int b, MPI_Datatype c, void *d, int e, MPI_Datatype f, int g, MPI_Comm h, MPI_Request *r) { *r = MPI_REQUEST_NULL; return MPI_Gather(a, b, c, d, e, f, g, h); }
int MPI_Igatherv(const void *a, int b, MPI_Datatype c, void *d, const int *e, const int *f, MPI_Datatype g, int h, MPI_Comm i, MPI_Request *r) { *r = MPI_REQUEST_NULL; return MPI_Gatherv(a, b, c, d, e, f, g, h, i); }
int MPI_Iscatter(const void *a, int b, MPI_Datatype c, void *d, int e, MPI_Datatype f, int g, MPI_Comm h, MPI_Request *r) { *r = MPI_REQUEST_NULL; return MPI_Scatter(a, b, c, d, e, f, g, h); }
int MPI_Iscatterv(const void *a, const int *b, const int *c, MPI_Datatype d, void *e, int f, MPI_Datatype g, int h, MPI_Comm i, MPI_Request *r) { *r = MPI_REQUEST_NULL; return MPI_Scatterv(a, b, c, d, e, f, g, h, i); }
int MPI_Iallreduce(const void *a, void *b, int c, MPI_Datatype d, MPI_Op e, MPI_Comm f, MPI_Request *r) { *r = MPI_REQUEST_NULL; return MPI_Allreduce(a, b, c, d, e, f); }

}  // extern "C"
