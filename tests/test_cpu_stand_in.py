"""CPU: the `-m gpu` test files that fit run against a STAND-IN library: the orchestration layer of libhpddm_b200.so -- hb_api.cu,
hb_krylov.cu, hb_geneo.cu, hb_gcrodr.cpp, hb_symbolic.cpp, i.e. every exported entry point and all of its host logic -- compiled with
g++ and linked with tests/native/device_mock.cpp, a host implementation of what lies below (CUDA runtime calls, kernel launchers,
local factorisation / triangular solves).  Pins, without a GPU: the Python mirror, the C ABI, halo planning and ordering, coarse layout,
apply / deflation orchestration of every correction and Prcndtnr, staging of caller memory, the device Krylov drivers, the C++ seams
under the UNMODIFIED examples/schwarz.cpp with 4-5 MPI-shim ranks, and the GPU test code itself.  Does not pin the CUDA kernels: those
are what the `-m gpu` run on the B200 is for.  See tests/tools/run_gpu_tests_on_stand_in.py (`--asan` runs under ASan + UBSan)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_test_files_pass_on_the_host_stand_in():
    """goldens of the unmodified reference through the C ABI, every entry point against the oracle, the complex instantiation, the
    device GCRO-DR / BGCRO-DR drivers (non-gating on the GPU: all 12 cases must pass here)"""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "run_gpu_tests_on_stand_in.py")], capture_output=True, text=True, timeout=1800, cwd=ROOT)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    m = re.search(r"(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= 90 and "failed" not in res.stdout.splitlines()[-1], tail
    x = re.search(r"(\d+) xpassed", res.stdout)
    assert x and int(x.group(1)) == 12 and "xfailed" not in res.stdout.splitlines()[-1], tail


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "schwarz_b200_full")), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_unmodified_reference_driver_on_both_seams_on_the_host_stand_in(tmp_path):
    """tests/test_gpu_dropin.py with the drivers' libhpddm_b200.so resolved to the stand-in (LD_LIBRARY_PATH precedes their RUNPATH):
    the reference's own examples/schwarz.cpp, unmodified, on HPDDM::B200Sub and on HPDDM::Schwarz<B200Sub, ...>, 4-5 forked ranks whose
    collectives go through MPI_Allgather of the shim -- same iteration counts as the pure-reference build, the reference's test lines"""
    sys.path.insert(0, ROOT)
    from tests.tools.run_gpu_tests_on_stand_in import build
    so = build(str(tmp_path), [])
    libdir = tmp_path / "lib"
    libdir.mkdir()
    shutil.copy(so, libdir / "libhpddm_b200.so")
    env = dict(os.environ, HPDDM_B200_TEST_STANDIN="1", LD_LIBRARY_PATH=str(libdir) + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_dropin.py"), os.path.join(ROOT, "tests", "test_gpu_zz_cpp_device_krylov.py"),
                          "-q", "-m", "gpu", "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=1800, cwd=ROOT)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    m = re.search(r"(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= 20 and "failed" not in res.stdout.splitlines()[-1], tail
    # the device-resident Krylov route of the C++ mirror under 4 MPI-shim ranks (GMRES, BGMRES, GCRO-DR, BGCRO-DR; two solves each):
    # non-gating on the GPU, all four must pass here
    x = re.search(r"(\d+) xpassed", res.stdout)
    assert x and int(x.group(1)) == 4 + 13 and "xfailed" not in res.stdout.splitlines()[-1], tail   # + the golden driver on the full seam, 13 cases
