"""GPU: the device-resident GCRO-DR and BGCRO-DR drivers (hpddm_b200[z]_solve_gcrodr / _solve_bgcrodr: Krylov basis and recycled pair
(U, C) in HBM) on the goldens of the unmodified reference's IterativeMethod::GCRODR / BGCRODR -- every solve of a sequence that shares the recycled pair must give the
reference's iteration count and solution.

Status of this file (round 2): the driver's logic (hb_gcrodr.cpp) is verified on the CPU against the same goldens through a host
vector backend (tests/test_cpu_gcrodr.py); its DEVICE backend (hb_krylov.cu: DeviceBackend, a thin layer over the kernels of the
GMRES / BGMRES drivers) and the exported entry points are verified on the CPU against a host stand-in for the CUDA runtime and the
kernel launchers (tests/test_cpu_krylov_mock.py: the reference's own 40X known-answer test, block splits, complex scalars) -- but
they were written after this round's GPU budget was spent and have not run on hardware yet.  Each case therefore
runs in its own process with a timeout and is marked xfail(strict=False): an XPASS in the report means the device path reproduced the
reference on the B200; a failure is recorded without stopping the (verified) rest of the suite, which is why the file sorts last.
The host-driven GCRO-DR on top of the GPU apply is a gating test (tests/test_gpu_golden.py)."""
import json
import os
import subprocess
import sys

import pytest

from tests.golden_util import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="device backend of the GCRO-DR driver not yet run on hardware (CPU-verified logic, see module docstring)")
@pytest.mark.parametrize("name", [n for n in cases() if "_gcrodr_" in n or "_bgcrodr_" in n])
def test_device_gcrodr_reproduces_the_reference(name):
    if os.environ.get("HPDDM_B200_TEST_STANDIN") == "1":   # CPU stand-in below the C ABI (tests/tools/run_gpu_tests_on_stand_in.py): same process
        from tests.tools.run_gcrodr_device import main as run_case
        out = run_case(name)
    else:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "run_gcrodr_device.py"), name], capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
        out = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
        print(out)
    assert out["its"] == out["ref"], out
    assert max(out["err"]) < 1e-7, out
    assert out["recycled_dim"] > 0 and out["recycled_dim_after_destroy"] == 0
    assert out["gmres_fallback"][0] == out["gmres_fallback"][1]
    assert out["launches"] > 0


@pytest.mark.xfail(strict=False, reason="device backend of the GCRO-DR / BGCRO-DR drivers not yet run on hardware (see module docstring)")
def test_recycled_pair_life_cycle():
    """The pair kept in the context across what can happen to it: a second solve, a new deflation space + coarse operator (C = A M^-1 U is
    recomputed when a solve starts), another number of right-hand sides (dropped and rebuilt), the block driver and back (pairs of the
    other kind are dropped), restart 2 / recycle >= restart / iteration limit inside the first cycle / zero right-hand side."""
    import numpy as np
    from oracle.generate import generate_world
    from oracle.schwarz import SchwarzWorld
    from tests.helpers import build_gpu_decomposition
    parts = generate_world(8, dim=3, N=(10, 10, 10), overlap=1, mu=1, neumann=True)
    w = SchwarzWorld(parts)
    w.multiplicity_scaling()
    w.numfact()
    w.solve_gevp([p["MatNeumann"] for p in parts], nu=3)
    w.build_coarse()
    deco = build_gpu_decomposition(parts, w, two_level=True)

    def converged(x, rhs, tol):
        res = w.compute_residual(x, rhs)
        return bool(np.all(res[:, 1] <= 10 * tol * res[:, 0]))

    b = [np.asfortranarray(p["f"][:, :1]) for p in parts]
    kw = dict(correction="deflated", restart=6, recycle=2, tol=1e-9)
    for _ in range(2):
        it, x, _ = deco.solve_gcrodr(b, **kw)
        assert converged(x, b, 1e-9) and deco.recycle_dim() == 2
    w.set_vectors([z[:, :2] for z in w.Z])
    w.build_coarse()
    for s, z in zip(deco.subs, w.Z):
        s.setVectors(z)
    deco.buildTwo()
    it, x, _ = deco.solve_gcrodr(b, **kw)
    assert converged(x, b, 1e-9)
    b2 = [np.asfortranarray(np.hstack([v, 2.0 * v[::-1]])) for v in b]
    for fn in (deco.solve_gcrodr, deco.solve_bgcrodr, deco.solve_gcrodr):
        it, x, _ = fn(b2, **kw)
        assert converged(x, b2, 1e-9) and deco.recycle_dim() == 2, fn.__name__
    for fn in (deco.solve_gcrodr, deco.solve_bgcrodr):
        deco.recycle_destroy()
        it, x, _ = fn(b2, correction="deflated", restart=2, recycle=5, tol=1e-6)
        assert converged(x, b2, 1e-6) and deco.recycle_dim() == 1
        deco.recycle_destroy()
        it, x, _ = fn(b2, correction="deflated", restart=3, recycle=2, max_it=2)
        assert it == 2
        it, x, _ = fn([0.0 * v for v in b2], correction="deflated", restart=5, recycle=2)
        assert it == 0 and all(np.all(v == 0) for v in x)
    deco.close()
