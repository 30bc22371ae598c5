"""numfact + solve residual sweep over subdomain sizes: python profiles/sweep.py m1 m2 ..."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from hpddm_b200 import Decomposition
from hpddm_b200.examples.generate import generate3d
for arg in sys.argv[1:]:
    dims = [int(v) for v in arg.split("x")]
    if len(dims) == 1: dims = dims * 3
    part = generate3d(0, 1, N=tuple(dims), overlap=1, mu=1, grid=(1, 1, 1))
    deco = Decomposition(0)
    s = deco.add(0)
    s.initialize(part["Mat"], part["o"], part["mapping"]); s.setGridHint(*part["dims"])
    deco.multiplicityScaling([part["d"]])
    t0 = time.time()
    try:
        s.callNumfact()
    except Exception as e:
        print(arg, "EXC", e, flush=True); deco.close(); continue
    t1 = time.time()
    b = part["f"]; x = s.solve(b); st = s.statistics()
    r = np.linalg.norm(part["Mat"] @ x - b) / np.linalg.norm(b)
    print(arg, "sym", st["symmetric"], "fronts", st["fronts"], "levels", st["levels"], "nnzL %.3g" % st["nnz_factor"], "numfact %.2fs" % (t1 - t0), "resid %.2e" % r, flush=True)
    deco.close()
