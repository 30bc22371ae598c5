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
