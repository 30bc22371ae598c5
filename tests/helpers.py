"""Shared test plumbing: build the GPU decomposition from oracle-generated parts."""
import numpy as np

from hpddm_b200 import Decomposition


def build_gpu_decomposition(parts, world=None, two_level=False, device=0, grid_hint=True, own_scaling=False,
                            own_coarse=True, ranks=None, deco=None, method="ras"):
    """parts: oracle.generate output for ALL ranks; `ranks` = the global ranks hosted here."""
    ranks = list(range(len(parts))) if ranks is None else list(ranks)
    dtype = np.complex128 if any(np.iscomplexobj(parts[r]["Mat"].data) for r in ranks) else np.float64
    deco = Decomposition(device, dtype=dtype) if deco is None else deco
    for r in ranks:
        p = parts[r]
        s = deco.add(r)
        s.initialize(p["Mat"], p["o"], p["mapping"], sym=p.get("sym", False))
        if grid_hint and "dims" in p:
            s.setGridHint(*p["dims"])
    if own_scaling or world is None:
        deco.multiplicityScaling([parts[r]["d"] for r in ranks])
    else:
        for s, r in zip(deco.subs, ranks):
            s.setScaling(world.d[r])
    for s in deco.subs:
        s.callNumfact(method=method)
    if two_level:
        for s, r in zip(deco.subs, ranks):
            s.setVectors(world.Z[r])
        if own_coarse:
            deco.buildTwo()
        else:
            deco.setCoarse(world.E)
    return deco


def relerr(got, ref):
    num = max(np.abs(g - r).max() for g, r in zip(got, ref))
    den = max(np.abs(r).max() for r in ref)
    return num / den
