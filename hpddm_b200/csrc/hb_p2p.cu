// Peer-memory fabric: every collective of the hot path carried by this library's own kernels over NVLink /
// NVSwitch peer memory (CUDA IPC, one process per GPU -- or several processes sharing one GPU), no NCCL call
// and no staging copy on the hot path:
//
//   * halo exchange (Subdomain::exchange, include/HPDDM_subdomain.hpp:115-130: MPI_Irecv / gthr / MPI_Isend /
//     MPI_Waitany): the pack kernel stores every shared value straight into the neighbour's receive window, a
//     one-warp kernel publishes a round number, the unpack-add kernel of the neighbour waits for it;
//   * coarse gather (CoarseOperator::callSolver, include/HPDDM_coarse_operator_impl.hpp:1706-1720: MPI_Gather(v) to
//     the main rank + scatter): every rank stores its nu x mu block into ALL peers' windows (all-gather by peer
//     stores), the coarse solve is replicated -- no scatter;
//   * the reductions of the device-resident Krylov drivers (MPI_Allreduce of include/HPDDM_GMRES.hpp:59-68,
//     include/HPDDM_CG.hpp:103,120): all-gather of the partial sums + a sum in rank order on every rank, i.e.
//     bit-identical results everywhere (all ranks must take the same convergence decisions).
//
// Every window region is double-buffered by round parity: a rank can be at most one collective of a kind ahead
// of a peer (it needs that peer's data of round r to finish round r), so two slots never collide.
// Bootstrap (exchange of the IPC handles and layout blobs) goes through ctrl_allgather: the host program's own
// communicator when it gave one (hpddm_b200_ctx_comm_init_host, e.g. MPI_Allgather) or NCCL.  If any rank
// cannot map a peer (not IPC-reachable), all ranks agree to fall back to NCCL for everything.
#include <cstring>
#include <map>

#include "hb_internal.h"

namespace hb {

struct P2P {
  bool tried = false, on = false, halo_on = false;
  int P = 0;
  int mu_cap = 0;  // halo columns the window holds
  int gcap = 0;    // all-gather: K elements per rank and parity
  int rcap = 0;    // all-reduce: doubles per rank and parity
  unsigned long long hround = 0, ground = 0, rround = 0;
  char *win = nullptr;  // local window: [halo recv 2 x h x mu_cap K][halo flags nb][gather 2 x P x gcap K][gather flags P][reduce 2 x P x rcap double][reduce flags P]
  size_t off_recv = 0, off_hflag = 0, off_gat = 0, off_gflag = 0, off_red = 0, off_rflag = 0;
  // halo tables (device)
  int *d_enb = nullptr;                        // neighbour index of every map entry
  K **d_peer_base = nullptr;                   // per neighbour: base of ITS halo receive region (mapped here)
  long long *d_peer_stride = nullptr;          // per neighbour: h_peer * mu_cap (slot stride)
  long long *d_peer_off = nullptr;             // per neighbour: offset (entries) of my segment in its layout
  unsigned long long **d_peer_flag = nullptr;  // per neighbour: address of my slot in ITS halo flag array
  // all-to-all tables (device), one entry per rank (self included)
  K **d_peer_gat = nullptr;                    // base of the peer's gather region
  unsigned long long **d_peer_gflag = nullptr; // my slot in the peer's gather flags
  double **d_peer_red = nullptr;
  unsigned long long **d_peer_rflag = nullptr;
  int *h_err = nullptr, *d_err = nullptr;      // mapped pinned host word: a kernel that gave up waiting sets it
  std::vector<void *> opened;
};

namespace {

struct Blob {  // what every rank publishes
  cudaIpcMemHandle_t hwin;
  unsigned long long off_recv, off_hflag, off_gat, off_gflag, off_red, off_rflag;
  int h, nb, mu_cap, gcap, rcap, ok, nsub;
  int ranks[64];
  int ptr[65];
};

constexpr long long SPIN_LIMIT = 40000000LL;  // x 200 ns: seconds -- never hang the GPU if a peer died

__device__ __forceinline__ bool wait_flag(const unsigned long long *flag, unsigned long long round) {
  unsigned long long v = 0;
  long long spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= round) return true;
    __nanosleep(200);
  } while (++spins < SPIN_LIMIT);
  return false;
}

__global__ void kk_pack_p2p(int h, int n, int mu, int parity, const int *__restrict__ map, const int *__restrict__ ebase, const int *__restrict__ esize,
                            const int *__restrict__ enb, const K *__restrict__ x, K *const *__restrict__ peer_base,
                            const long long *__restrict__ peer_stride, const long long *__restrict__ peer_off) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)h * mu) return;
  const int e = (int)(t % h), c = (int)(t / h);
  const int i = enb[e];
  K *dst = peer_base[i] + parity * peer_stride[i] + peer_off[i] * mu + (int64_t)c * esize[e] + (e - ebase[e]);
  *dst = x[map[e] + (int64_t)c * n];  // NVLink store into the neighbour's HBM
}
__global__ void kk_signal_p2p(int nb, unsigned long long round, unsigned long long *const *__restrict__ peer_flag) {
  const int i = threadIdx.x;
  if (i < nb) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag[i]), "l"(round) : "memory");
  }
}
__global__ void kk_unpack_p2p(int nuniq, int n, int mu, int nb, unsigned long long round, const unsigned long long *flags, const int *__restrict__ uidx,
                              const int *__restrict__ useg, const int *__restrict__ upos, const int *__restrict__ ebase, const int *__restrict__ esize,
                              const K *recv, K *x, int *err) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if ((int)threadIdx.x < nb && !wait_flag(flags + threadIdx.x, round)) {
    bad = 1;
    *reinterpret_cast<volatile int *>(err) = 1;
  }
  __syncthreads();
  if (bad) return;
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)nuniq * mu) return;
  const int u = (int)(t % nuniq), c = (int)(t / nuniq);
  K acc = x[uidx[u] + (int64_t)c * n];
  for (int q = useg[u]; q < useg[u + 1]; ++q) {
    const int e = upos[q];
    const K *src = recv + (int64_t)ebase[e] * mu + (int64_t)c * esize[e] + (e - ebase[e]);
#ifdef HB_COMPLEX
    double vr, vi;
    asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(vr), "=d"(vi) : "l"(src));
    acc += mk(vr, vi);
#else
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(src));
    acc += v;
#endif
  }
  x[uidx[u] + (int64_t)c * n] = acc;
}

// all-to-all push: CTA p stores `count` doubles of `src` into slot `me` of peer p's region, then publishes the round to it
__global__ void __launch_bounds__(256) kk_push_all(int me, int count, long long slot_stride, long long parity_off, const double *__restrict__ src,
                                                   double *const *__restrict__ peer_region, unsigned long long *const *__restrict__ peer_flag,
                                                   unsigned long long round) {
  const int p = blockIdx.x;
  double *dst = peer_region[p] + parity_off + (long long)me * slot_stride;
  for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag[p]), "l"(round) : "memory");
}
// wait for every rank's block of this round, then  MODE 0: out[p * count + i] = region[p][i]  (all-gather, rank-major)
//                                                   MODE 1 / 2: out[i] = sum / max over p in rank order  (all-reduce, identical on all ranks)
template <int MODE>
__global__ void __launch_bounds__(256) kk_wait_all(int P, int count, long long slot_stride, const double *region, const unsigned long long *flags,
                                                   unsigned long long round, double *out, int *err) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  for (int p = threadIdx.x; p < P; p += blockDim.x)
    if (!wait_flag(flags + p, round)) {
      bad = 1;
      *reinterpret_cast<volatile int *>(err) = 1;
    }
  __syncthreads();
  if (bad) return;
  if (MODE == 0) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P * count; t += gridDim.x * blockDim.x) {
      const int p = t / count, i = t - p * count;
      double v;
      asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(region + (long long)p * slot_stride + i));
      out[t] = v;
    }
  } else {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
      double acc = 0.0;
      for (int p = 0; p < P; ++p) {
        double v;
        asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(region + (long long)p * slot_stride + i));
        acc = (MODE == 1 || p == 0) ? (p == 0 ? v : acc + v) : fmax(acc, v);
      }
      out[i] = acc;
    }
  }
}

template <class T>
int up_table(const std::vector<T> &v, T **d, cudaStream_t st) {
  HB_CUDA(cudaMalloc(d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) HB_CUDA(cudaMemcpyAsync(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

}  // namespace

void p2p_free(Ctx *c) {
  P2P *p = c->p2p;
  if (!p) return;
  for (void *q : p->opened) cudaIpcCloseMemHandle(q);
  for (void *q : {(void *)p->win, (void *)p->d_enb, (void *)p->d_peer_base, (void *)p->d_peer_stride, (void *)p->d_peer_off, (void *)p->d_peer_flag,
                  (void *)p->d_peer_gat, (void *)p->d_peer_gflag, (void *)p->d_peer_red, (void *)p->d_peer_rflag})
    if (q) cudaFree(q);
  if (p->h_err) cudaFreeHost(p->h_err);
  cudaGetLastError();
  delete p;
  c->p2p = nullptr;
}

static bool fabric_wanted() {
  // HPDDM_B200_HALO=nccl (or HPDDM_B200_FABRIC=0) keeps every collective on NCCL, for A/B measurements
  static const bool off = (getenv("HPDDM_B200_HALO") && !strcmp(getenv("HPDDM_B200_HALO"), "nccl")) || (getenv("HPDDM_B200_FABRIC") && !strcmp(getenv("HPDDM_B200_FABRIC"), "0"));
  return !off;
}

// collective; (re)creates the window for `mu` halo columns and the current coarse layout.  Any failure on any rank -> everybody
// falls back to NCCL (which must then exist: a host-bootstrapped communicator without NCCL reports the error).
int fabric_setup(Ctx *c, int mu) {
  if (c->nproc <= 1) return 0;
  const int P = c->nproc;
  if (c->p2p) {
    cudaStreamSynchronize(c->stream);
    p2p_free(c);
  }
  P2P *p = c->p2p = new P2P;
  p->tried = true;
  p->P = P;
  {  // size the window for the largest request of any rank (ranks may have grown their work space at different times)
    int want[2] = {mu, std::max(c->Lnu, 1)};
    std::vector<int> wants(2 * P);
    HB_CHECK(ctrl_allgather(c, want, wants.data(), sizeof(want)));
    for (int q = 0; q < P; ++q) mu = std::max(mu, wants[2 * q]);
  }
  const bool one_sub = c->subs.size() == 1;
  Sub *s = one_sub ? c->subs[0] : nullptr;
  const int nb = s ? (int)s->nb_rank.size() : 0, h = s ? s->h : 0;
  p->mu_cap = mu;
  p->gcap = std::max(c->Lnu, 1) * mu;
  p->rcap = 2048;
  auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = 0;
  p->off_recv = off;
  off = align(off + (size_t)2 * h * mu * sizeof(K));
  p->off_hflag = off;
  off = align(off + (size_t)std::max(nb, 1) * 8);
  p->off_gat = off;
  off = align(off + (size_t)2 * P * p->gcap * sizeof(K));
  p->off_gflag = off;
  off = align(off + (size_t)P * 8);
  p->off_red = off;
  off = align(off + (size_t)2 * P * p->rcap * sizeof(double));
  p->off_rflag = off;
  off = align(off + (size_t)P * 8);
  Blob mine;
  memset(&mine, 0, sizeof(mine));
  mine.ok = (fabric_wanted() && nb <= 64) ? 1 : 0;
  mine.h = h;
  mine.nb = nb;
  mine.mu_cap = mu;
  mine.gcap = p->gcap;
  mine.rcap = p->rcap;
  mine.nsub = (int)c->subs.size();
  mine.off_recv = p->off_recv;
  mine.off_hflag = p->off_hflag;
  mine.off_gat = p->off_gat;
  mine.off_gflag = p->off_gflag;
  mine.off_red = p->off_red;
  mine.off_rflag = p->off_rflag;
  for (int i = 0; i < nb && i < 64; ++i) mine.ranks[i] = s->nb_rank[i];
  for (int i = 0; i <= nb && i < 65; ++i) mine.ptr[i] = s->nb_ptr[i];
  if (mine.ok) {
    if (cudaMalloc(&p->win, off) != cudaSuccess) mine.ok = 0;
    else if (cudaMemset(p->win, 0, off) != cudaSuccess || cudaIpcGetMemHandle(&mine.hwin, p->win) != cudaSuccess) mine.ok = 0;
    if (mine.ok && (cudaHostAlloc(&p->h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(&p->d_err, p->h_err, 0) != cudaSuccess)) mine.ok = 0;
    if (mine.ok) *p->h_err = 0;
  }
  cudaGetLastError();
  std::vector<Blob> all(P);
  HB_CHECK(ctrl_allgather(c, &mine, all.data(), sizeof(Blob)));
  bool ok = true;
  for (int q = 0; q < P; ++q) ok = ok && all[q].ok && all[q].mu_cap == mu && all[q].gcap == p->gcap;
  bool halo_ok = ok;
  for (int q = 0; q < P; ++q) halo_ok = halo_ok && all[q].nsub == 1;
  std::vector<char *> peer(P, nullptr);
  if (ok)
    for (int q = 0; q < P && ok; ++q) {
      if (q == c->proc_rank) {
        peer[q] = p->win;
        continue;
      }
      void *pw = nullptr;
      if (cudaIpcOpenMemHandle(&pw, all[q].hwin, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      p->opened.push_back(pw);
      peer[q] = static_cast<char *>(pw);
    }
  std::vector<K *> base(nb, nullptr);
  std::vector<long long> stride(nb, 0), offs(nb, 0);
  std::vector<unsigned long long *> pflag(nb, nullptr);
  if (ok && halo_ok)
    for (int i = 0; i < nb; ++i) {
      const int q = s->nb_rank[i];  // one subdomain per process: global rank == process rank
      if (q < 0 || q >= P) {
        halo_ok = false;
        break;
      }
      const Blob &B = all[q];
      int k = -1;
      for (int t = 0; t < B.nb; ++t)
        if (B.ranks[t] == s->grank) k = t;
      if (k < 0 || B.ptr[k + 1] - B.ptr[k] != s->nb_ptr[i + 1] - s->nb_ptr[i]) {
        halo_ok = false;
        break;
      }
      base[i] = reinterpret_cast<K *>(peer[q] + B.off_recv);
      stride[i] = (long long)B.h * mu;
      offs[i] = B.ptr[k];
      pflag[i] = reinterpret_cast<unsigned long long *>(peer[q] + B.off_hflag) + k;
    }
  // agree on the outcome
  int verdict[2] = {ok ? 1 : 0, halo_ok ? 1 : 0};
  std::vector<int> verdicts(2 * P);
  HB_CHECK(ctrl_allgather(c, verdict, verdicts.data(), sizeof(verdict)));
  for (int q = 0; q < P; ++q) {
    ok = ok && verdicts[2 * q];
    halo_ok = halo_ok && verdicts[2 * q + 1];
  }
  if (!ok) {
    if (getenv("HPDDM_B200_DEBUG")) fprintf(stderr, "[hpddm_b200] peer-memory fabric unavailable: using NCCL\n");
    if (!c->nccl) {
      set_error("peer-memory fabric unavailable (a peer GPU is not CUDA-IPC reachable, or HPDDM_B200_HALO=nccl) and no NCCL communicator was initialised");
      return HPDDM_B200_ERR_NCCL;
    }
    return 0;  // p->on stays false
  }
  if (halo_ok) {
    std::vector<int> enb(h);
    for (int i = 0; i < nb; ++i)
      for (int e = s->nb_ptr[i]; e < s->nb_ptr[i + 1]; ++e) enb[e] = i;
    HB_CHECK(up_table(enb, &p->d_enb, c->stream));
    HB_CHECK(up_table(base, &p->d_peer_base, c->stream));
    HB_CHECK(up_table(stride, &p->d_peer_stride, c->stream));
    HB_CHECK(up_table(offs, &p->d_peer_off, c->stream));
    HB_CHECK(up_table(pflag, &p->d_peer_flag, c->stream));
  }
  std::vector<K *> pgat(P);
  std::vector<double *> pred(P);
  std::vector<unsigned long long *> pgf(P), prf(P);
  for (int q = 0; q < P; ++q) {
    pgat[q] = reinterpret_cast<K *>(peer[q] + all[q].off_gat);
    pgf[q] = reinterpret_cast<unsigned long long *>(peer[q] + all[q].off_gflag) + c->proc_rank;
    pred[q] = reinterpret_cast<double *>(peer[q] + all[q].off_red);
    prf[q] = reinterpret_cast<unsigned long long *>(peer[q] + all[q].off_rflag) + c->proc_rank;
  }
  HB_CHECK(up_table(pgat, &p->d_peer_gat, c->stream));
  HB_CHECK(up_table(pgf, &p->d_peer_gflag, c->stream));
  HB_CHECK(up_table(pred, &p->d_peer_red, c->stream));
  HB_CHECK(up_table(prf, &p->d_peer_rflag, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  p->on = true;
  p->halo_on = halo_ok;
  // nobody may push into a window before every rank has finished (re)creating its own: one more control-plane round
  int token = 1;
  std::vector<int> tokens(P);
  HB_CHECK(ctrl_allgather(c, &token, tokens.data(), sizeof(int)));
  return 0;
}

bool fabric_on(Ctx *c) { return c->p2p && c->p2p->on; }

// returns 1 when the exchange was done over peer memory, 0 when the caller must use NCCL, < 0 on error
int p2p_halo(Ctx *c, K *const *x, int mu) {
  P2P *p = c->p2p;
  if (!p || !p->on || !p->halo_on || c->subs.size() != 1) return 0;
  if (mu > p->mu_cap) {
    set_error("peer-memory halo: %d columns but the window holds %d (internal error: ensure_capacity rebuilds it)", mu, p->mu_cap);
    return HPDDM_B200_ERR_STATE;
  }
  Sub *s = c->subs[0];
  if (s->h == 0) return 1;
  const int nb = (int)s->nb_rank.size();
  p->hround++;
  const int parity = (int)(p->hround & 1);
  kk_pack_p2p<<<(unsigned)(((int64_t)s->h * mu + 255) / 256), 256, 0, c->stream>>>(s->h, s->n, mu, parity, s->d_map, s->d_ebase, s->d_esize, p->d_enb, x[0],
                                                                                  p->d_peer_base, p->d_peer_stride, p->d_peer_off);
  kk_signal_p2p<<<1, 64, 0, c->stream>>>(nb, p->hround, p->d_peer_flag);
  kk_unpack_p2p<<<(unsigned)(((int64_t)s->nuniq * mu + 255) / 256), 256, 0, c->stream>>>(
      s->nuniq, s->n, mu, nb, p->hround, reinterpret_cast<unsigned long long *>(p->win + p->off_hflag), s->d_uidx, s->d_useg, s->d_upos, s->d_ebase, s->d_esize,
      reinterpret_cast<K *>(p->win + p->off_recv) + (size_t)parity * s->h * p->mu_cap, x[0], p->d_err);
  c->launches += 3;
  HB_CUDA(cudaGetLastError());
  return 1;
}

const K *p2p_last_halo_window(Ctx *c) {
  P2P *p = c->p2p;
  if (!p || !p->on || !p->halo_on || c->subs.size() != 1 || p->hround == 0) return nullptr;
  return reinterpret_cast<const K *>(p->win + p->off_recv) + (size_t)(p->hround & 1) * c->subs[0]->h * p->mu_cap;
}

// in-place all-gather of `count` K elements per rank (buf = P blocks, rank-major; this rank's block already in place)
int fabric_allgather(Ctx *c, K *buf, int count) {
  P2P *p = c->p2p;
  if (!p || !p->on) return 0;
  if (count > p->gcap) {
    set_error("peer-memory all-gather: %d elements per rank but the window holds %d", count, p->gcap);
    return HPDDM_B200_ERR_STATE;
  }
  if (count == 0) return 1;
  p->ground++;
  const int parity = (int)(p->ground & 1);
  const long long slot = (long long)p->gcap * KD, poff = (long long)parity * p->P * slot;
  kk_push_all<<<p->P, 256, 0, c->stream>>>(c->proc_rank, count * KD, slot, poff, reinterpret_cast<const double *>(buf + (size_t)c->proc_rank * count),
                                           reinterpret_cast<double *const *>(p->d_peer_gat), p->d_peer_gflag, p->ground);
  kk_wait_all<0><<<1, 256, 0, c->stream>>>(p->P, count * KD, slot, reinterpret_cast<const double *>(p->win + p->off_gat) + poff,
                                           reinterpret_cast<const unsigned long long *>(p->win + p->off_gflag), p->ground, reinterpret_cast<double *>(buf), p->d_err);
  c->launches += 2;
  HB_CUDA(cudaGetLastError());
  return 1;
}

// in-place all-reduce of `count` doubles: op 0 = sum (in rank order, identical on every rank), 1 = max
int fabric_allreduce(Ctx *c, double *buf, int count, int op) {
  P2P *p = c->p2p;
  if (!p || !p->on) return 0;
  for (int done = 0; done < count; done += p->rcap) {
    const int cnt = std::min(p->rcap, count - done);
    p->rround++;
    const int parity = (int)(p->rround & 1);
    const long long slot = p->rcap, poff = (long long)parity * p->P * slot;
    kk_push_all<<<p->P, 256, 0, c->stream>>>(c->proc_rank, cnt, slot, poff, buf + done, p->d_peer_red, p->d_peer_rflag, p->rround);
    const double *region = reinterpret_cast<const double *>(p->win + p->off_red) + poff;
    const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(p->win + p->off_rflag);
    if (op == 0) kk_wait_all<1><<<1, 256, 0, c->stream>>>(p->P, cnt, slot, region, flags, p->rround, buf + done, p->d_err);
    else kk_wait_all<2><<<1, 256, 0, c->stream>>>(p->P, cnt, slot, region, flags, p->rround, buf + done, p->d_err);
    c->launches += 2;
  }
  HB_CUDA(cudaGetLastError());
  return 1;
}

// after a stream synchronisation: did a kernel give up waiting for a peer?  (the flag lives in mapped host memory: no copy)
int p2p_check(Ctx *c) {
  if (!c->p2p || !c->p2p->on || !c->p2p->h_err) return 0;
  if (*reinterpret_cast<volatile int *>(c->p2p->h_err)) {
    set_error("peer-memory fabric: timed out waiting for a peer's data (a rank died or left the collective sequence)");
    return HPDDM_B200_ERR_NCCL;
  }
  return 0;
}

}  // namespace hb
