"""GPU: the device-resident Krylov route of the C++ mirror -- HPDDM::B200Schwarz::solve -> hpddm_b200_solve / _solve_bgmres /
_solve_gcrodr / _solve_bgcrodr by -hpddm_krylov_method, read like the reference's drivers read their options -- under the reference's
generator and MPI ranks (oracle/_ref/b200_full_driver -device_krylov 1, 4 ranks sharing GPU 0, control plane over MPI_Allgather:
B200Schwarz::setCommunicatorHost), two successive solves, against the UNMODIFIED reference's own driver on the same arguments
(oracle/_ref/ref_driver): identical iteration counts for both solves.

Non-gating like tests/test_gpu_zz_gcrodr_device.py (added after the round's GPU budget was spent; xfail(strict=False): an XPASS means it
ran green on the B200).  On the CPU the same file runs against the host stand-in of the device layer and must pass
(tests/test_cpu_stand_in.py)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.path.join(ROOT, "oracle", "_ref", "b200_full_driver")
REFDRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
pytestmark = pytest.mark.gpu


@pytest.mark.xfail(strict=False, reason="device-resident route of the C++ mirror under MPI ranks: not yet run on hardware (see module docstring)")
@pytest.mark.skipif(not os.path.exists(FULL) or not os.path.exists(REFDRV), reason="oracle/_ref drivers not built (need /root/reference at build time)")
@pytest.mark.parametrize("krylov", [["-hpddm_krylov_method", "gmres", "-hpddm_gmres_restart", "8"],
                                    ["-hpddm_krylov_method", "bgmres", "-hpddm_gmres_restart", "8"],
                                    ["-hpddm_krylov_method", "gcrodr", "-hpddm_recycle", "3", "-hpddm_gmres_restart", "8"],
                                    ["-hpddm_krylov_method", "bgcrodr", "-hpddm_recycle", "3", "-hpddm_gmres_restart", "8"]])
def test_cpp_mirror_device_resident_solves_match_the_reference(tmp_path, krylov):
    args = ["-hpddm_schwarz_method", "ras", "-hpddm_schwarz_coarse_correction", "deflated", "-deflation_vectors", "3", "-Nx", "60", "-Ny", "60", "-hpddm_verbosity", "1",
            "-solves", "2"] + krylov
    env = dict(os.environ, HPDDM_SHIM_NP="4", HPDDM_B200_NDEV="1", HPDDM_B200_BOOT="host", HPDDM_REF_DUMP=str(tmp_path / "g"))
    ref = subprocess.run([REFDRV] + args, env=env, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    got = subprocess.run([FULL] + args + ["-device_krylov", "1"], env=env, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert got.returncode == 0, (got.stdout + got.stderr)[-2000:]
    want = [int(v) for v in re.findall(r"ref_driver:.*?\bit (\d+)", ref.stdout)]
    have = [int(re.search(r"b200_full_driver: 4 ranks, it (\d+)", got.stdout).group(1)), int(re.search(r"b200_full_driver: solve 2, it (\d+)", got.stdout).group(1))]
    assert len(want) == 2 and have == want, (have, want)


# ---- the golden-vector driver itself on the FULL seam ------------------------------------------------------------------------------
# oracle/ref_build/ref_driver.cpp -- the program that produced tests/golden/*.npz from the UNMODIFIED reference -- compiled a second
# time with -DB200SUB -DB200SCHWARZ, i.e. on HPDDM::Schwarz<HPDDM::B200Sub, ...> (oracle/_ref/ref_driver_b200_full[_z]): every dump
# (partition of unity, Subdomain::exchange, GMV, one-level apply, deflation, the three corrections, solutions, iteration counts of
# every solve of a sequence) now comes from the CUDA library through the C++ seam and is compared with the golden FIELD BY FIELD.
# Cases with random right-hand sides are not reproducible (std::random_device in the reference's generator) and are left out; so are the
# 6-rank goldens, whose two-level outputs carry the reference's dense-LAPACK-plugin quirk (E^T y = r for a sparse coarse pattern, DESIGN.md
# section 6) that the library deliberately does not reproduce.
FULLDRV = os.path.join(ROOT, "oracle", "_ref", "ref_driver_b200_full")
SEAM_CASES = [("small_40x40_p4_ras", []), ("small_40x40_p4_twolevel_nu3", []), ("small_36x36_p4_symcsr_twolevel_nu2", []),
              ("small_40x40_p4_penalised_ras", []), ("small_40x40_p4_asm_cg", []), ("complex_40x40_p4_twolevel_nu3", []), ("complex_40x40_p4_penalised_ras", []),
              ("small_40x40_p4_gcrodr_m8_k4_solves3", ["-device_krylov", "1"]), ("small_40x40_p4_gcrodr_m6_k2_twolevel_solves2", ["-device_krylov", "1"]),
              ("small_40x40_p4_gcrodr_m12_k4_same_solves3", ["-device_krylov", "1"]), ("small_40x40_p4_bgcrodr_m8_k4_solves2", ["-device_krylov", "1"]),
              ("complex_40x40_p4_gcrodr_m8_k3_solves2", ["-device_krylov", "1"]), ("small_40x40_p4_twolevel_nu3", ["-device_krylov", "1"])]


# on the GPU box six of them (every run starts 4 processes with a CUDA context each); the host stand-in run takes them all
if os.environ.get("HPDDM_B200_TEST_STANDIN") != "1":
    SEAM_CASES = [c for c in SEAM_CASES if c[0] in ("small_40x40_p4_penalised_ras", "complex_40x40_p4_twolevel_nu3", "small_40x40_p4_bgcrodr_m8_k4_solves2",
                                                    "small_40x40_p4_gcrodr_m6_k2_twolevel_solves2", "complex_40x40_p4_gcrodr_m8_k3_solves2") or c == ("small_40x40_p4_twolevel_nu3", [])]


@pytest.mark.xfail(strict=False, reason="golden driver on the full seam: not yet run on hardware (see module docstring)")
@pytest.mark.skipif(not os.path.exists(FULLDRV), reason="oracle/_ref/ref_driver_b200_full not built (needs /root/reference at build time)")
@pytest.mark.parametrize("name,extra", SEAM_CASES)
def test_golden_driver_on_the_full_seam_reproduces_every_dump(tmp_path, name, extra):
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle.ref_build.make_goldens import read_dump
    gold = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    P = int(gold["np"])
    args = str(gold["args"]).split()
    binary = FULLDRV + ("_z" if name.startswith("complex") else "")
    env = dict(os.environ, HPDDM_SHIM_NP=str(P), HPDDM_REF_DUMP=str(tmp_path / "g"))
    res = subprocess.run([binary] + args + extra, env=env, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout + res.stderr)[-2000:]
    checked = 0
    for r in range(P):
        got = read_dump(str(tmp_path / f"g_{r}.bin"))
        for key, val in got.items():
            want = gold[f"r{r}_{key}"]
            if key.startswith("iterations") or key in ("header", "ia", "ja", "o") or key.startswith("mapping"):
                assert np.array_equal(val, want), (r, key, val, want)
            else:
                tol = 1e-6 if key.startswith("sol") or key.startswith("residual") else 1e-10     # solutions: both are iterates at tol 1e-6 .. 1e-9
                assert np.abs(val - want).max() <= tol * max(np.abs(want).max(), 1e-300), (r, key)
            checked += 1
    assert checked >= 15 * P
