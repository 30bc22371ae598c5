import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.generate import generate_world
from oracle.schwarz import SchwarzWorld
from tests.helpers import build_gpu_decomposition, relerr
def run(tag, parts, w, env):
    for k in ("HPDDM_B200_FORCE_LU","HPDDM_B200_SMALL","HPDDM_B200_LEAF"): os.environ.pop(k, None)
    os.environ.update(env)
    try:
        deco = build_gpu_decomposition(parts, w)
        errs=[]
        for r,s in enumerate(deco.subs):
            b = parts[r]["f"][:, :1]
            errs.append(relerr([s.solve(b)], [w.solver[r].solve(b)]))
        st = deco.subs[0].statistics()
        print(tag, env, "sym", st["symmetric"], "fronts", st["fronts"], "levels", st["levels"], "err", max(errs), flush=True)
        deco.close()
    except Exception as e:
        print(tag, env, "EXC", e, flush=True)
parts = generate_world(2, dim=3, mu=1, N=(12, 12, 12), overlap=1); w = SchwarzWorld(parts); w.multiplicity_scaling(); w.numfact()
for env in ({}, {"HPDDM_B200_SMALL":"0"}, {"HPDDM_B200_FORCE_LU":"1"}, {"HPDDM_B200_FORCE_LU":"1","HPDDM_B200_SMALL":"0"}, {"HPDDM_B200_FORCE_LU":"1","HPDDM_B200_SMALL":"100000"}):
    run("sym3d", parts, w, env)
parts = generate_world(4, dim=2, mu=0, Nx=100, Ny=100, overlap=1); w = SchwarzWorld(parts); w.multiplicity_scaling(); w.numfact()
for env in ({}, {"HPDDM_B200_SMALL":"0"}, {"HPDDM_B200_SMALL":"100000"}, {"HPDDM_B200_LEAF":"100000"}, {"HPDDM_B200_LEAF":"100000","HPDDM_B200_SMALL":"0"}):
    run("quirk2d", parts, w, env)
parts = generate_world(4, dim=2, mu=0, Nx=24, Ny=24, overlap=1); w = SchwarzWorld(parts); w.multiplicity_scaling(); w.numfact()
for env in ({}, {"HPDDM_B200_SMALL":"0"}, {"HPDDM_B200_LEAF":"100000"}, {"HPDDM_B200_LEAF":"8"}):
    run("quirk2d-small", parts, w, env)
