"""Profiling driver: build one m^3 Poisson subdomain, run a few local solves / applies.
usage: python profiles/run_solve.py M [nsolve] [napply] [d|z] [mu]     (z: complex Helmholtz subdomain, hpddm_b200z_*; mu right-hand sides)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from hpddm_b200 import Decomposition, capi
from hpddm_b200.examples.generate import generate3d
from bench import cosine_modes

m = int(sys.argv[1]); nsolve = int(sys.argv[2]) if len(sys.argv) > 2 else 3; napply = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cplx = len(sys.argv) > 4 and sys.argv[4] == "z"
mu = int(sys.argv[5]) if len(sys.argv) > 5 else 1
if cplx:
    from hpddm_b200.examples.generate import generate_helmholtz3d
    part = generate_helmholtz3d(0, 1, N=(m, m, m), overlap=1, mu=1, grid=(1, 1, 1), k=2.0, nu=20)
else:
    part = generate3d(0, 1, N=(m, m, m), overlap=1, mu=1, grid=(1, 1, 1))
deco = Decomposition(0, dtype=np.complex128 if cplx else np.float64)
s = deco.add(0)
s.initialize(part["Mat"], part["o"], part["mapping"]); s.setGridHint(*part["dims"])
deco.multiplicityScaling([part["d"]])
t0 = time.time(); s.callNumfact(); deco.synchronize(); t1 = time.time()
st = s.statistics()
print("numfact %.2fs" % (t1 - t0), st, flush=True)
n = part["ndof"]
x = torch.rand(n * mu, dtype=torch.complex128 if cplx else torch.float64, device="cuda"); y = torch.empty_like(x)
for _ in range(nsolve):
    deco.api.check(deco.api.sub_solve(s.h, x.data_ptr(), y.data_ptr(), mu, capi.DEVICE))
deco.synchronize()
# residual check of the last solve
A = part["Mat"]; xs = x.cpu().numpy()[-n:]; ys = y.cpu().numpy()[-n:]
print("solve residual", np.linalg.norm(A @ ys - xs) / np.linalg.norm(xs), flush=True)
if napply:
    s.setVectors(part["Z"] if cplx else cosine_modes(part["dims"], 20)); deco.buildTwo()
    for _ in range(napply):
        deco.apply_device([x], [y], 1, "deflated")
    deco.synchronize()
deco.close()
