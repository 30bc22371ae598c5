"""GPU: the reference's OWN driver (examples/schwarz.cpp + generate.cpp, unmodified, built by
oracle/ref_build) with its SUBDOMAIN plugin replaced by HPDDM::B200Sub -- 4 forked ranks, local
factorisations and triangular solves on the GPU, everything else reference code."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "schwarz_b200")
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/schwarz_b200 not built (needs /root/reference at build time)")
def test_reference_driver_with_b200_subdomain_solver(tmp_path):
    env = dict(os.environ, HPDDM_SHIM_NP="4")
    res = subprocess.run([BIN, "-hpddm_schwarz_method", "ras", "-hpddm_gmres_restart=25", "-hpddm_max_it", "80", "-hpddm_verbosity", "1"],
                         env=env, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-2000:]            # examples/schwarz.cpp:140-144 pass/fail
    m = re.search(r"converges after\s+(\d+)\s+iteration", out)
    assert m, out[-2000:]
    golden = np.load(os.path.join(ROOT, "tests", "golden", "config1_100x100_p4_ras.npz"))
    assert int(m.group(1)) == int(golden["r0_iterations"][0])   # identical to the all-CPU reference run (33)
