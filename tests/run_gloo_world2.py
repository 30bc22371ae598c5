"""world_size-2 CPU check of the N>1 host logic (launched by tests/test_multiproc_cpu.py with torchrun, gloo):
per-rank generation, neighbour-map agreement between ranks, the distributed halo sum and D-weighted dots
against the in-process oracle, and the communicator bootstrap broadcast used by Decomposition.comm_init_torch."""
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
    dist.barrier()
    if rank == 0:
        print("gloo world-2 OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
