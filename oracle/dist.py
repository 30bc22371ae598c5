"""Distributed (one process per rank) restatement of the communication steps, over
torch.distributed (any backend; the CPU tests use gloo).  TEST INFRASTRUCTURE ONLY.

Subdomain::exchange (include/HPDDM_subdomain.hpp:115-130): per neighbour Irecv / gather /
Isend, then add in completion order -- here in neighbour-rank order."""
import numpy as np
import torch
import torch.distributed as dist


class RankSubdomain:
    def __init__(self, part, d):
        order = sorted(range(len(part["o"])), key=lambda q: part["o"][q])
        self.map = [(int(part["o"][q]), np.asarray(part["mapping"][q], dtype=np.int64)) for q in order if len(part["mapping"][q]) > 0]
        self.d = np.asarray(d, dtype=np.float64)

    def subdomain_exchange(self, x):
        """x: (n, mu) array of this rank, modified in place."""
        mu = x.shape[1]
        recv = [torch.empty(idx.size * mu, dtype=torch.float64) for _, idx in self.map]
        send = [torch.from_numpy(np.ascontiguousarray(x[idx, :].T).reshape(-1).copy()) for _, idx in self.map]
        reqs = []
        for (nb, _), r, s in zip(self.map, recv, send):
            reqs.append(dist.irecv(r, src=nb))
            reqs.append(dist.isend(s, dst=nb))
        for q in reqs:
            q.wait()
        for (nb, idx), r in zip(self.map, recv):
            np.add.at(x, (idx, slice(None)), r.numpy().reshape(mu, idx.size).T)
        return x

    def exchange(self, x):
        x *= self.d[:, None]
        return self.subdomain_exchange(x)

    def dot(self, x, y):
        t = torch.from_numpy((self.d[:, None] * x * y).sum(axis=0).copy())
        dist.all_reduce(t)
        return t.numpy()
