"""Runs the device-resident GCRO-DR / BGCRO-DR driver (hpddm_b200[z]_solve_gcrodr / _solve_bgcrodr, by the golden's Krylov method) on one golden of the unmodified reference and prints one
JSON line: iteration counts of every solve of the sequence next to the reference's, the relative error of every solution, the
dimension of the recycled pair kept in the context, kernel launches.  Run by tests/test_gpu_zz_gcrodr_device.py in its own
process (a fault in the driver must not take the CUDA context of the test session with it).

    python tests/tools/run_gcrodr_device.py small_40x40_p4_gcrodr_m8_k4_solves3
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.golden_util import load  # noqa: E402
from tests.helpers import build_gpu_decomposition  # noqa: E402


def main(name):
    parts, ref, meta = load(name)
    P = meta["P"]
    for p in parts:
        p["dims"] = None
    deco = build_gpu_decomposition(parts, None, own_scaling=True, grid_hint=False, method=meta["method"])
    corr = None
    if meta["nu"] > 0:
        for s, r in zip(deco.subs, range(P)):
            s.setVectors(ref[r]["Z"].reshape(meta["nu"], -1).T)
        deco.buildTwo()
        corr = "deflated"
    block = meta["krylov"] == "bgcrodr"
    solver = deco.solve_bgcrodr if block else deco.solve_gcrodr
    out = dict(case=name, its=[], ref=[], err=[], res=[])
    for s in range(1, meta["solves"] + 1):
        tag = "" if s == 1 else str(s)
        b = [parts[r]["f"] if s == 1 else ref[r]["f" + tag] for r in range(P)]
        it, x, res = solver(b, correction=corr, restart=meta["restart"], recycle=meta["recycle"], max_it=meta["max_it"], tol=meta["tol"],
                                       target=meta["recycle_target"], same_system=min(s, 2) if meta["same_system"] else 0)
        gold = [ref[r]["sol" + tag] for r in range(P)]
        out["its"].append(int(it))
        out["ref"].append(int(ref[0]["iterations" + tag][0]))
        out["err"].append(float(max(np.abs(x[r] - gold[r]).max() / np.abs(gold[r]).max() for r in range(P))))
        out["res"].append(float(np.max(res)))
    out["recycled_dim"] = deco.recycle_dim()
    deco.recycle_destroy()
    out["recycled_dim_after_destroy"] = deco.recycle_dim()
    # recycle = 0 is GMRES (GCRODR.hpp:50-55): same count as the plain device driver
    b = [parts[r]["f"] for r in range(P)]
    plain = deco.solve_bgmres if block else deco.solve
    out["gmres_fallback"] = [int(solver(b, correction=corr, restart=meta["restart"], recycle=0, max_it=meta["max_it"], tol=meta["tol"])[0]),
                             int(plain(b, correction=corr, restart=meta["restart"], max_it=meta["max_it"], tol=meta["tol"])[0])]
    out["launches"] = int(deco.launches)
    deco.close()
    print(json.dumps(out))
    return out


if __name__ == "__main__":
    main(sys.argv[1])
