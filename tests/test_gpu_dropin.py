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


BINZ = os.path.join(ROOT, "oracle", "_ref", "schwarz_b200_z")


@pytest.mark.skipif(not os.path.exists(BINZ), reason="oracle/_ref/schwarz_b200_z not built (needs /root/reference at build time)")
def test_reference_driver_complex_build_with_b200_subdomain_solver(tmp_path):
    """The same unmodified driver compiled with -DFORCE_COMPLEX (K = std::complex<double>, examples/schwarz.hpp:56-60):
    SUBDOMAIN = HPDDM::B200Sub<std::complex<double>> -> hpddm_b200z_* (complex no-pivot LU + complex SpTRSV on the GPU).
    The generator's matrix and right-hand side are real-valued, so complex GMRES must take exactly the iterations of
    the real run (33, the all-CPU reference golden)."""
    env = dict(os.environ, HPDDM_SHIM_NP="4")
    res = subprocess.run([BINZ, "-hpddm_schwarz_method", "ras", "-hpddm_gmres_restart=25", "-hpddm_max_it", "80", "-hpddm_verbosity", "1"],
                         env=env, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-2000:]
    m = re.search(r"converges after\s+(\d+)\s+iteration", out)
    assert m, out[-2000:]
    golden = np.load(os.path.join(ROOT, "tests", "golden", "config1_100x100_p4_ras.npz"))
    assert int(m.group(1)) == int(golden["r0_iterations"][0])


FULL = os.path.join(ROOT, "oracle", "_ref", "b200_full_driver")
REFDRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver")


def _iterations(out, tag):
    m = re.search(tag + r".*?\bit (\d+)", out)
    assert m, out[-1500:]
    return int(m.group(1))


@pytest.mark.skipif(not os.path.exists(FULL), reason="oracle/_ref/b200_full_driver not built")
@pytest.mark.parametrize("extra", [[], ["-hpddm_krylov_method", "cg"], ["-deflation_vectors", "3"],
                                   ["-hpddm_krylov_method", "bgmres", "-generate_random_rhs", "4"],    # the reference's BGMRES, block apply with mu = 4
                                   ["-hpddm_krylov_method", "gcrodr", "-hpddm_recycle", "5"]])         # recycling driver (the Krylov method of config 5)
def test_reference_krylov_drivers_on_b200schwarz_single_gpu(tmp_path, extra):
    """IterativeMethod::solve (GMRES / CG, unmodified reference code) driving HPDDM::B200Schwarz, 1 rank."""
    env = dict(os.environ, HPDDM_SHIM_NP="1", HPDDM_B200_NDEV="1")
    res = subprocess.run([FULL, "-hpddm_schwarz_method", "ras", "-Nx", "60", "-Ny", "60", "-hpddm_verbosity", "1"] + extra,
                         env=env, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-2000:]
    assert _iterations(out, "b200_full_driver") <= 3      # one subdomain: the preconditioner is the exact inverse


@pytest.mark.skipif(not os.path.exists(FULL) or not os.path.exists(REFDRV), reason="oracle/_ref drivers not built")
@pytest.mark.parametrize("krylov", [[], ["-hpddm_krylov_method", "gcrodr", "-hpddm_recycle", "5", "-hpddm_gmres_restart", "10"]])
def test_reference_krylov_drivers_on_b200schwarz_multi_gpu(tmp_path, krylov):
    """One rank per GPU (NCCL halo + coarse gather), reference GMRES on top; iteration count must equal
    the all-CPU reference run of the same case (oracle/_ref/ref_driver)."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    P = 4 if ngpu >= 4 else 2
    args = ["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3", "-Nx", "60", "-Ny", "60", "-hpddm_verbosity", "1"] + krylov
    env = dict(os.environ, HPDDM_SHIM_NP=str(P), HPDDM_B200_NDEV=str(ngpu), HPDDM_REF_DUMP=str(tmp_path / "g"))
    ref = subprocess.run([REFDRV] + args, env=env, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    got = subprocess.run([FULL] + args, env=env, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert got.returncode == 0, (got.stdout + got.stderr)[-2000:]
    assert _iterations(got.stdout, "b200_full_driver") == _iterations(ref.stdout, "ref_driver")
