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


# ------------------------------------------------------------------------------------------------------------------------------
# FULL seam: the unmodified examples/schwarz.cpp instantiating HPDDM::Schwarz<HPDDM::B200Sub, ...> (hpddm_b200/host/
# HPDDM_B200_schwarz.hpp) -- the whole apply / deflation / exchange / GMV / numfact / coarse operator / GenEO on the GPU, the Krylov
# drivers reference host code -- against the pure-reference build of the same file (oracle/_ref/schwarz_ref, dense LAPACK solvers).
# The argument lists are the reference's own test lines for schwarz_cpp (Makefile:310-347 of hpddm/hpddm).
FULLBIN = os.path.join(ROOT, "oracle", "_ref", "schwarz_b200_full")
FULLBINZ = os.path.join(ROOT, "oracle", "_ref", "schwarz_b200_full_z")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "schwarz_ref")
needs_full = pytest.mark.skipif(not (os.path.exists(FULLBIN) and os.path.exists(REFBIN)), reason="oracle/_ref/schwarz_b200_full / schwarz_ref not built (need /root/reference at build time)")


def _driver(binary, nranks, args, tmp_path, debug=True):
    env = dict(os.environ, HPDDM_SHIM_NP=str(nranks))
    if debug:
        env["HPDDM_B200_DEBUG"] = "1"
    res = subprocess.run([binary] + args, env=env, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    out = res.stdout + res.stderr
    m = re.search(r"converges after\s+(\d+)\s+iteration", out)
    return res.returncode, (int(m.group(1)) if m else None), out


def _on_gpu(out, nranks):
    """every rank reports its kernel launches at destruction (HPDDM_B200_DEBUG): the path really ran on the device, and with
    more than one rank the collectives went over the peer-memory fabric (transport 2)"""
    lines = re.findall(r"HPDDM::Schwarz<B200Sub>: (\d+) kernel launches, transport (\d)", out)
    assert len(lines) == nranks, out[-2000:]
    assert all(int(n) > 50 for n, _ in lines) and all(int(t) == (2 if nranks > 1 else 0) for _, t in lines), lines


@needs_full
@pytest.mark.parametrize("args", [
    ["-hpddm_verbosity=1", "--hpddm_gmres_restart=25", "-hpddm_max_it", "80"],                                                   # BASELINE config 1 (33 iterations)
    ["-hpddm_verbosity=1", "-hpddm_schwarz_method", "asm", "-hpddm_krylov_method", "cg", "-Nx", "40", "-Ny", "40"],
    ["-hpddm_verbosity=1", "-hpddm_schwarz_coarse_correction", "deflated", "-hpddm_geneo_nu=0", "-Nx", "60", "-Ny", "60"],      # constant deflation vector (schwarz.cpp:117-122)
    ["-hpddm_verbosity=1", "-hpddm_schwarz_coarse_correction", "additive", "-hpddm_geneo_nu=0", "-symmetric_csr", "-Nx", "60", "-Ny", "60"],
    ["-hpddm_verbosity=1", "-hpddm_schwarz_coarse_correction", "balanced", "-hpddm_geneo_nu=0", "-Nx", "60", "-Ny", "60", "-overlap", "2"],
    ["-hpddm_verbosity=1", "-hpddm_schwarz_coarse_correction", "deflated", "-hpddm_geneo_nu=0", "-hpddm_krylov_method", "gcrodr", "-hpddm_recycle", "5", "-hpddm_gmres_restart", "10", "-Nx", "60", "-Ny", "60"],
])
def test_unmodified_driver_on_the_full_gpu_path_matches_the_pure_reference(tmp_path, args):
    rc_ref, it_ref, out_ref = _driver(REFBIN, 4, args, tmp_path, debug=False)
    rc, it, out = _driver(FULLBIN, 4, args, tmp_path)
    assert it_ref is not None, out_ref[-1500:]
    assert rc == rc_ref, out[-3000:]                 # the driver's own verdict (it <= 45, relative residual <= 1e-2: schwarz.cpp:140-144), same as the reference's
    assert it == it_ref, (it, it_ref)                # identical Krylov iteration count
    _on_gpu(out, 4)


@needs_full
@pytest.mark.parametrize("nranks,args", [
    # reference Makefile:312: several right-hand sides (random: only the driver's verdict can be compared), MGS
    (4, ["-hpddm_verbosity=1", "--hpddm_gmres_restart=25", "-hpddm_max_it", "80", "-generate_random_rhs", "4", "-hpddm_orthogonalization=mgs"]),
    # Makefile:314
    (4, ["-hpddm_verbosity=1", "--hpddm_gmres_restart=25", "-hpddm_max_it", "80", "-generate_random_rhs", "4", "-hpddm_schwarz_coarse_correction", "deflated", "-hpddm_geneo_nu=0"]),
    # EIGENSOLVER lines (GenEO on the GPU): Makefile:319,320,327,334,340
    (2, ["-hpddm_schwarz_coarse_correction", "deflated", "-hpddm_geneo_nu=2", "-hpddm_verbosity=2", "-symmetric_csr", "--hpddm_gmres_restart", "20"]),
    (4, ["-hpddm_schwarz_coarse_correction", "deflated", "-hpddm_geneo_nu=10", "-hpddm_verbosity=4", "--hpddm_gmres_restart=15", "-hpddm_max_it", "80"]),
    (4, ["-hpddm_schwarz_coarse_correction", "deflated", "-hpddm_geneo_nu=10", "-hpddm_verbosity=2", "-nonuniform", "-Nx", "50", "-Ny", "50", "-symmetric_csr", "-hpddm_level_2_p", "2", "-hpddm_gmres_restart=25"]),
    (4, ["-hpddm_schwarz_coarse_correction", "deflated", "-hpddm_geneo_nu=10", "-hpddm_verbosity=2", "-nonuniform", "-Nx", "50", "-Ny", "50", "-symmetric_csr", "-hpddm_level_2_p", "2",
         "-generate_random_rhs", "8", "-hpddm_krylov_method=bgmres", "-hpddm_gmres_restart=10", "-hpddm_deflation_tol=1e-4", "-hpddm_gmres_restart=25"]),
    (4, ["-hpddm_schwarz_coarse_correction", "additive", "-hpddm_geneo_nu=1", "-hpddm_verbosity=2", "-Nx", "20", "-Ny", "20", "-symmetric_csr", "-hpddm_level_2_p", "2", "-generate_random_rhs", "4",
         "-hpddm_krylov_method=bfbcg", "-hpddm_deflation_tol=1e-4", "-hpddm_schwarz_method", "asm", "-hpddm_geneo_threshold", "1e+1"]),
    (5, ["-hpddm_myPrefix_schwarz_coarse_correction", "deflated", "-hpddm_myPrefix_geneo_nu=10", "-hpddm_myPrefix_verbosity=2", "-nonuniform", "-Nx", "50", "-Ny", "50", "-symmetric_csr",
         "-hpddm_myPrefix_gmres_restart=25", "-hpddm_verbosity=2", "-prefix=myPrefix_", "-hpddm_myPrefix_geneo_threshold", "0.2"]),
])
def test_reference_test_lines_of_schwarz_cpp_pass_on_the_full_gpu_path(tmp_path, nranks, args):
    rc, it, out = _driver(FULLBIN, nranks, args, tmp_path)
    assert rc == 0, out[-3000:]                      # examples/schwarz.cpp:140-144 (it <= 45 / 60, relative residual <= 1e-2)
    _on_gpu(out, nranks)


@pytest.mark.skipif(not os.path.exists(FULLBINZ), reason="oracle/_ref/schwarz_b200_full_z not built")
def test_unmodified_driver_complex_build_on_the_full_gpu_path(tmp_path):
    """-DFORCE_COMPLEX build of the same file: HPDDM::Schwarz<B200Sub, ..., std::complex<double>> -> hpddm_b200z_*; real-valued data,
    so exactly the 33 iterations of the real run"""
    rc, it, out = _driver(FULLBINZ, 4, ["-hpddm_verbosity=1", "--hpddm_gmres_restart=25", "-hpddm_max_it", "80"], tmp_path)
    assert rc == 0 and it == 33, out[-3000:]
    _on_gpu(out, 4)
