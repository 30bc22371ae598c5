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
