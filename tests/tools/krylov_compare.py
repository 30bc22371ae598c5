"""Full Krylov solves on one B200 hosting a 2 x 2 x 2 decomposition (8 subdomains of M^3 cells + overlap, two-level deflated RAS,
nu cosine modes per subdomain), MU random right-hand sides: host-driven GMRES over the C ABI (host vectors, what an unchanged
Krylov driver does) vs the device-resident drivers hpddm_b200_solve (GMRES, all columns together), hpddm_b200_solve_bgmres and --
for the symmetric one-level ASM variant -- hpddm_b200_solve_cg.   usage: python tests/tools/krylov_compare.py [M] [MU] [nu]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from hpddm_b200 import Decomposition, KrylovOperator
from hpddm_b200.examples.generate import generate3d
from bench import cosine_modes
from oracle.krylov import gmres   # the restated reference driver, used here only as the HOST-driven caller

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
MU = int(sys.argv[2]) if len(sys.argv) > 2 else 4
NU = int(sys.argv[3]) if len(sys.argv) > 3 else 8
P, grid = 8, (2, 2, 2)
parts = [generate3d(r, P, N=(2 * M,) * 3, overlap=1, mu=MU, grid=grid) for r in range(P)]
out = {"subdomains": P, "cells_per_subdomain_edge": M, "n_loc": parts[0]["ndof"], "rhs": MU, "nu": NU}
for method in ("ras", "asm"):
    deco = Decomposition(0)
    for r, p in enumerate(parts):
        s = deco.add(r)
        s.initialize(p["Mat"], p["o"], p["mapping"])
        s.setGridHint(*p["dims"])
    deco.multiplicityScaling([p["d"] for p in parts])
    for s in deco.subs:
        s.callNumfact(method=method)
    corr = None
    if method == "ras":
        for s, p in zip(deco.subs, parts):
            s.setVectors(cosine_modes(p["dims"], NU))
        deco.buildTwo()
        corr = "deflated"
    b = deco.exchange([p["f"] for p in parts], scaled=True)
    runs = [("gmres_device", lambda: deco.solve(b, correction=corr)), ("bgmres_device", lambda: deco.solve_bgmres(b, correction=corr))] if method == "ras" else \
           [("cg_device", lambda: deco.solve_cg(b, correction=corr)), ("gmres_device", lambda: deco.solve(b, correction=corr))]
    res = {}
    for name, fn in runs:
        fn()
        deco.synchronize()
        t0 = time.time()
        it, x, rr = fn()
        res[name] = {"seconds": round(time.time() - t0, 4), "iterations": it, "max_rel_residual": float(np.max(rr))}
    if method == "ras":
        t0 = time.time()
        it, x, applies = gmres(KrylovOperator(deco, corr), b)
        res["gmres_host_driven"] = {"seconds": round(time.time() - t0, 4), "iterations": it}
    out["two-level deflated RAS" if method == "ras" else "one-level ASM"] = res
    deco.close()
print(json.dumps(out))
