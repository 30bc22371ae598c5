// Halo exchange over NVLink peer memory: the pack kernel stores every shared value straight into the
// neighbour's receive buffer (CUDA-IPC mapped, one process per GPU), a one-warp kernel publishes a
// round number, and the unpack-add kernel of the neighbour waits for it -- no NCCL call, no staging copy
// on the hot path (NCCL stays in charge of the bootstrap, the coarse all-gather and the Krylov reductions).
//
// Replaces the MPI_Irecv / gthr / MPI_Isend / MPI_Waitany loop of Subdomain::exchange
// (include/HPDDM_subdomain.hpp:115-130) for decompositions with one subdomain per process.
// Receive buffers are double-buffered by round parity: a rank can be at most one exchange ahead of a
// neighbour (it needs that neighbour's data of round r to finish round r), so two slots never collide.
#include <cstring>
#include <map>

#include "hb_internal.h"

namespace hb {

struct P2P {
  bool tried = false, on = false;
  int mu_cap = 0;
  unsigned long long round = 0;
  K *recv2 = nullptr;                   // 2 * h * mu_cap
  unsigned long long *flags = nullptr;  // one per neighbour, written remotely
  int *d_enb = nullptr;                 // neighbour index of every map entry
  K **d_peer_base = nullptr;       // per neighbour: base of ITS receive buffer (mapped here)
  long long *d_peer_stride = nullptr;   // per neighbour: h_peer * mu_cap (slot stride)
  long long *d_peer_off = nullptr;      // per neighbour: offset (entries) of my segment in its layout
  unsigned long long **d_peer_flag = nullptr;  // per neighbour: address of my slot in ITS flag array
  int *d_err = nullptr;
  std::vector<void *> opened;
};

namespace {

struct Blob {  // what every rank publishes
  cudaIpcMemHandle_t hrecv, hflag;
  int h, nb, mu_cap, ok;
  int ranks[64];
  int ptr[65];
};

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
  if ((int)threadIdx.x < nb) {
    unsigned long long v = 0;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
      if (v >= round) break;
      __nanosleep(200);
    } while (++spins < 20000000LL);  // ~ seconds: never hang the GPU if a neighbour died
    if (v < round) {
      bad = 1;
      atomicExch(err, 1);
    }
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

}  // namespace

void p2p_free(Ctx *c) {
  P2P *p = c->p2p;
  if (!p) return;
  for (void *q : p->opened) cudaIpcCloseMemHandle(q);
  for (void *q : {(void *)p->recv2, (void *)p->flags, (void *)p->d_enb, (void *)p->d_peer_base, (void *)p->d_peer_stride, (void *)p->d_peer_off,
                  (void *)p->d_peer_flag, (void *)p->d_err})
    if (q) cudaFree(q);
  delete p;
  c->p2p = nullptr;
}

// collective; (re)creates the mapped buffers for `mu` columns.  Any failure on any rank -> everybody falls back to NCCL.
static int p2p_setup(Ctx *c, int mu) {
  Sub *s = c->subs[0];
  const int nb = (int)s->nb_rank.size(), P = c->nproc;
  if (c->p2p) {
    cudaStreamSynchronize(c->stream);
    p2p_free(c);
  }
  P2P *p = c->p2p = new P2P;
  p->tried = true;
  Blob mine;
  memset(&mine, 0, sizeof(mine));
  mine.h = s->h;
  mine.nb = nb;
  mine.mu_cap = mu;
  mine.ok = (nb <= 64) ? 1 : 0;
  for (int i = 0; i < nb && i < 64; ++i) mine.ranks[i] = s->nb_rank[i];
  for (int i = 0; i <= nb && i < 65; ++i) mine.ptr[i] = s->nb_ptr[i];
  if (cudaMalloc(&p->recv2, std::max<size_t>((size_t)2 * s->h * mu, 1) * sizeof(K)) != cudaSuccess) mine.ok = 0;
  if (cudaMalloc(&p->flags, std::max(nb, 1) * sizeof(unsigned long long)) != cudaSuccess) mine.ok = 0;
  if (mine.ok) {
    cudaMemset(p->flags, 0, std::max(nb, 1) * sizeof(unsigned long long));
    if (cudaIpcGetMemHandle(&mine.hrecv, p->recv2) != cudaSuccess || cudaIpcGetMemHandle(&mine.hflag, p->flags) != cudaSuccess) mine.ok = 0;
  }
  cudaGetLastError();
  // all-gather the blobs
  std::vector<Blob> all(P);
  char *dbuf = nullptr;
  HB_CUDA(cudaMalloc(&dbuf, (size_t)P * sizeof(Blob)));
  HB_CUDA(cudaMemcpyAsync(dbuf + (size_t)c->proc_rank * sizeof(Blob), &mine, sizeof(Blob), cudaMemcpyHostToDevice, c->stream));
  HB_CHECK(nccl_allgather_bytes(c, dbuf + (size_t)c->proc_rank * sizeof(Blob), dbuf, sizeof(Blob)));
  HB_CUDA(cudaMemcpyAsync(all.data(), dbuf, (size_t)P * sizeof(Blob), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(dbuf);
  bool ok = true;
  for (int q = 0; q < P; ++q) ok = ok && all[q].ok && all[q].mu_cap == mu;
  std::vector<K *> base(nb, nullptr);
  std::vector<long long> stride(nb, 0), off(nb, 0);
  std::vector<unsigned long long *> pflag(nb, nullptr);
  if (ok) {
    std::map<int, std::pair<void *, void *>> open;  // rank -> (recv, flags)
    for (int i = 0; i < nb && ok; ++i) {
      const int q = s->nb_rank[i];  // one subdomain per process: global rank == process rank
      const Blob &B = all[q];
      int k = -1;
      for (int t = 0; t < B.nb; ++t)
        if (B.ranks[t] == s->grank) k = t;
      if (k < 0 || B.ptr[k + 1] - B.ptr[k] != s->nb_ptr[i + 1] - s->nb_ptr[i]) {
        ok = false;
        break;
      }
      if (!open.count(q)) {
        void *pr = nullptr, *pf = nullptr;
        if (cudaIpcOpenMemHandle(&pr, B.hrecv, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess || cudaIpcOpenMemHandle(&pf, B.hflag, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = false;
          break;
        }
        p->opened.push_back(pr);
        p->opened.push_back(pf);
        open[q] = {pr, pf};
      }
      base[i] = static_cast<K *>(open[q].first);
      stride[i] = (long long)B.h * mu;
      off[i] = B.ptr[k];
      pflag[i] = static_cast<unsigned long long *>(open[q].second) + k;
    }
  }
  // agree on the outcome
  double *dflag = nullptr;
  HB_CUDA(cudaMalloc(&dflag, sizeof(double)));
  const double fail = ok ? 0.0 : 1.0;
  HB_CUDA(cudaMemcpyAsync(dflag, &fail, sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HB_CHECK(nccl_allreduce_sum(c, dflag, 1));
  double tot = 0.0;
  HB_CUDA(cudaMemcpyAsync(&tot, dflag, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(dflag);
  if (tot != 0.0) {
    if (getenv("HPDDM_B200_DEBUG")) fprintf(stderr, "[hpddm_b200] peer-memory halo unavailable on %d rank(s): using NCCL send/recv\n", (int)tot);
    return 0;  // p->on stays false
  }
  std::vector<int> enb(s->h);
  for (int i = 0; i < nb; ++i)
    for (int e = s->nb_ptr[i]; e < s->nb_ptr[i + 1]; ++e) enb[e] = i;
  HB_CUDA(cudaMalloc(&p->d_enb, std::max<size_t>(s->h, 1) * sizeof(int)));
  HB_CUDA(cudaMalloc(&p->d_peer_base, std::max(nb, 1) * sizeof(K *)));
  HB_CUDA(cudaMalloc(&p->d_peer_stride, std::max(nb, 1) * sizeof(long long)));
  HB_CUDA(cudaMalloc(&p->d_peer_off, std::max(nb, 1) * sizeof(long long)));
  HB_CUDA(cudaMalloc(&p->d_peer_flag, std::max(nb, 1) * sizeof(unsigned long long *)));
  HB_CUDA(cudaMalloc(&p->d_err, sizeof(int)));
  HB_CUDA(cudaMemsetAsync(p->d_err, 0, sizeof(int), c->stream));
  HB_CUDA(cudaMemcpyAsync(p->d_enb, enb.data(), s->h * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(cudaMemcpyAsync(p->d_peer_base, base.data(), nb * sizeof(K *), cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(cudaMemcpyAsync(p->d_peer_stride, stride.data(), nb * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(cudaMemcpyAsync(p->d_peer_off, off.data(), nb * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(cudaMemcpyAsync(p->d_peer_flag, pflag.data(), nb * sizeof(unsigned long long *), cudaMemcpyHostToDevice, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  p->mu_cap = mu;
  p->round = 0;
  p->on = true;
  return 0;
}

// returns 1 when the exchange was done over peer memory, 0 when the caller must use NCCL, < 0 on error
int p2p_halo(Ctx *c, K *const *x, int mu) {
  // opt-in for now (HPDDM_B200_HALO=p2p): verified against the oracle on 2 GPUs this round, not yet on 8
  static const bool enabled = getenv("HPDDM_B200_HALO") && !strcmp(getenv("HPDDM_B200_HALO"), "p2p");
  if (!enabled || c->nproc <= 1 || c->subs.size() != 1) return 0;
  // first use, or more columns than the mapped buffers hold: (re)build collectively; a failed attempt is not retried
  if (!c->p2p || (c->p2p->on && mu > c->p2p->mu_cap)) HB_CHECK(p2p_setup(c, mu));
  P2P *p = c->p2p;
  if (!p || !p->on) return 0;
  Sub *s = c->subs[0];
  if (s->h == 0) return 1;
  const int nb = (int)s->nb_rank.size();
  p->round++;
  const int parity = (int)(p->round & 1);
  kk_pack_p2p<<<(unsigned)(((int64_t)s->h * mu + 255) / 256), 256, 0, c->stream>>>(s->h, s->n, mu, parity, s->d_map, s->d_ebase, s->d_esize, p->d_enb, x[0],
                                                                                  p->d_peer_base, p->d_peer_stride, p->d_peer_off);
  kk_signal_p2p<<<1, 64, 0, c->stream>>>(nb, p->round, p->d_peer_flag);
  kk_unpack_p2p<<<(unsigned)(((int64_t)s->nuniq * mu + 255) / 256), 256, 0, c->stream>>>(s->nuniq, s->n, mu, nb, p->round, p->flags, s->d_uidx, s->d_useg,
                                                                                          s->d_upos, s->d_ebase, s->d_esize,
                                                                                          p->recv2 + (size_t)parity * s->h * p->mu_cap, x[0], p->d_err);
  c->launches += 3;
  HB_CUDA(cudaGetLastError());
  return 1;
}

int p2p_check(Ctx *c) {  // called at synchronisation points: did an unpack give up waiting?
  if (!c->p2p || !c->p2p->on) return 0;
  int e = 0;
  HB_CUDA(cudaMemcpyAsync(&e, c->p2p->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  HB_CUDA(cudaStreamSynchronize(c->stream));
  if (e) {
    set_error("peer-memory halo: timed out waiting for a neighbour's data");
    return HPDDM_B200_ERR_NCCL;
  }
  return 0;
}

}  // namespace hb
