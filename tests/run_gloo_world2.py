"""world_size-2 CPU check of the N>1 host logic (launched by tests/test_multiproc_cpu.py with torchrun, gloo):
per-rank generation, neighbour-map agreement between ranks, the distributed halo sum and D-weighted dots
against the in-process oracle, the communicator bootstrap broadcast used by Decomposition.comm_init_torch, and -- through the
host-only entry points hpddm_b200_debug_halo_schedule / _coarse_layout -- the product's own message ordering and coarse layout."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpddm_b200.examples.generate import generate3d, generate_world, split_grid_3d  # noqa: E402
from oracle.dist import RankSubdomain  # noqa: E402
from oracle.schwarz import SchwarzWorld  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    grid = split_grid_3d(world)
    N = tuple(g * 6 for g in grid)
    mine = generate3d(rank, world, N=N, overlap=2, mu=3, grid=grid)
    # every rank also builds the whole world in-process as the checker
    parts = generate_world(world, dim=3, N=N, overlap=2, mu=3, grid=grid)
    w = SchwarzWorld(parts)
    w.multiplicity_scaling()
    # 1. neighbour lists agree across ranks: sizes of the shared index sets
    sizes = torch.zeros(world, world, dtype=torch.int64)
    for nb, m in zip(mine["o"], mine["mapping"]):
        sizes[rank, nb] = len(m)
    dist.all_reduce(sizes)
    assert torch.equal(sizes, sizes.T), sizes
    # 2. distributed halo sum == in-process oracle
    sub = RankSubdomain(mine, w.d[rank])
    rs = np.random.RandomState(5)
    xs = [np.asfortranarray(rs.standard_normal(p["f"].shape)) for p in parts]
    ref = w.exchange([v.copy() for v in xs])[rank]
    got = sub.exchange(xs[rank].copy())
    assert np.abs(got - ref).max() < 1e-14
    # 3. D-weighted dots
    assert np.abs(sub.dot(xs[rank], xs[rank]) - w.dot(xs, xs)).max() < 1e-10
    # 4. the 128-byte id broadcast of Decomposition.comm_init_torch (payload only; NCCL itself needs GPUs)
    payload = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(payload, 0)
    assert bytes(payload.numpy().tobytes()) == bytes(range(128))
    # 5. the PRODUCT's own N > 1 host logic (libhpddm_b200.so, host-only entry points -- no GPU needed): the NCCL message schedule
    #    of a halo round with two subdomains per process, and the padded coarse layout for a non-uniform coarse space
    import ctypes as C
    from hpddm_b200 import capi
    L = capi.lib()
    P4 = 2 * world
    grid4 = split_grid_3d(P4)
    parts4 = [generate3d(r, P4, N=tuple(g * 5 for g in grid4), overlap=1, mu=1, grid=grid4) for r in range(P4)]
    mine4 = [2 * rank, 2 * rank + 1]                      # contiguous global ranks per process (include/hpddm_b200.h conventions)
    granks = np.array(mine4, dtype=np.int32)
    nbc = np.array([len(parts4[g]["o"]) for g in mine4], dtype=np.int32)
    nbr = np.array([nb for g in mine4 for nb in sorted(parts4[g]["o"])], dtype=np.int32)
    nmsg = C.c_int(0)
    capi.check(L.hpddm_b200_debug_halo_schedule(2, capi.ptr(granks), capi.ptr(nbc), capi.ptr(nbr), None, None, C.byref(nmsg)))
    sends = np.zeros(4 * nmsg.value, dtype=np.int32)
    recvs = np.zeros(4 * nmsg.value, dtype=np.int32)
    capi.check(L.hpddm_b200_debug_halo_schedule(2, capi.ptr(granks), capi.ptr(nbc), capi.ptr(nbr), capi.ptr(sends), capi.ptr(recvs), C.byref(nmsg)))
    sends, recvs = sends.reshape(-1, 4), recvs.reshape(-1, 4)
    remote = sum(1 for g in mine4 for nb in parts4[g]["o"] if nb // 2 != rank)
    assert nmsg.value == remote and len(sends) == len(recvs) == remote
    pad = 64
    buf = torch.full((2, pad, 2), -1, dtype=torch.int64)
    buf[0, :len(sends)] = torch.from_numpy(sends[:, :2].astype(np.int64))
    buf[1, :len(recvs)] = torch.from_numpy(recvs[:, :2].astype(np.int64))
    allb = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(allb, buf)
    for a in range(world):                                 # the k-th send a -> b is the k-th receive b posts for a: same (dst, src) sequence
        for b_ in range(world):
            if a == b_:
                continue
            s_ab = [tuple(v.tolist()) for v in allb[a][0] if v[0] >= 0 and int(v[0]) // 2 == b_]
            r_ba = [tuple(v.tolist()) for v in allb[b_][1] if v[0] >= 0 and int(v[1]) // 2 == a]
            assert s_ab == r_ba and len(s_ab) > 0, (a, b_, s_ab, r_ba)
    rows = np.array([3 + 2 * p for p in range(world)], dtype=np.int32)   # non-uniform coarse rows per process
    off = np.zeros(world + 1, dtype=np.int32)
    lmax = C.c_int(0)
    capi.check(L.hpddm_b200_debug_coarse_layout(world, capi.ptr(rows), capi.ptr(off), C.byref(lmax)))
    assert off.tolist() == [0] + np.cumsum(rows).tolist() and lmax.value == int(rows.max())
    dist.barrier()
    if rank == 0:
        print("gloo world-2 OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
