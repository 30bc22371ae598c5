"""Runs `-m gpu` test files on a machine WITHOUT a GPU against a stand-in library: the orchestration layer of libhpddm_b200.so
(hb_api.cu, hb_krylov.cu, hb_geneo.cu, hb_gcrodr.cpp, hb_symbolic.cpp -- every exported entry point) compiled with g++ and linked with
tests/native/device_mock.cpp, a host implementation of what lies below it (CUDA runtime calls, kernel launchers, local factorisation /
triangular solves, dense coarse inverse).  What this exercises: the Python mirror, the C ABI, all host logic of the library and the
test code itself.  What it cannot: the CUDA kernels (hb_kernels.cu, hb_solve.cu, hb_numfact.cu, hb_p2p.cu) -- those need the B200.
The product never loads this library: the path is patched into hpddm_b200.capi from here, for this process only.

    python tests/tools/run_gpu_tests_on_stand_in.py [--asan] [pytest arguments ...]        (default: the single-process test files)
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SOURCES = ["hb_api.cu", "hb_krylov.cu", "hb_geneo.cu", "hb_gcrodr.cpp", "hb_symbolic.cpp"]
DEFAULT = ["tests/test_gpu_golden.py", "tests/test_gpu_parity.py", "tests/test_gpu_zcomplex.py", "tests/test_gpu_zz_gcrodr_device.py"]
# tests of the default files that need the real device layer: subdomains beyond the dense-band LU of the stand-in (order > 11 585), or
# behaviour of the real multifrontal factorisation itself (Cholesky breakdown -> LU fallback inside the fronts)
NEEDS_DEVICE = ["test_single_subdomain_direct_solve_residual", "test_pageable_host_vectors_are_pinned_in_place_inside_a_start_end_bracket",
                "test_cholesky_breakdown_in_any_large_front_of_a_level_falls_back_to_lu", "test_complex_solve_round_trip_at_scale",
                "test_complex_apply_is_complex_linear_and_blocks_equal_columns", "test_complex_two_level_apply_of_A_times_coarse_vector_returns_it"]


def build(out_dir, extra):
    """Compiles the stand-in (both scalar builds) and returns its path.  The result is cached per content of the sources and flags under
    the system's temporary directory (three test modules need it); `out_dir` receives a copy."""
    import hashlib
    import shutil
    csrc = os.path.join(ROOT, "hpddm_b200", "csrc")
    h = hashlib.sha256(" ".join(extra).encode())
    for d in (csrc, os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "native")):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cpp", ".h", ".hpp")):
                h.update(f.encode())
                h.update(open(os.path.join(d, f), "rb").read())
    cache = os.path.join(tempfile.gettempdir(), "hpddm_b200_standin_" + h.hexdigest()[:16])
    cached = os.path.join(cache, "libhpddm_b200_standin.so")
    if not os.path.exists(cached):
        work = tempfile.mkdtemp(prefix="hpddm_b200_standin_build_")
        so = _compile(work, extra)
        os.makedirs(cache, exist_ok=True)
        os.replace(so, cached + f".{os.getpid()}")
        os.replace(cached + f".{os.getpid()}", cached)     # atomic: concurrent builders race benignly
        shutil.rmtree(work, ignore_errors=True)
    out = os.path.join(out_dir, "libhpddm_b200_standin.so")
    shutil.copy(cached, out)
    return out


def _compile(out_dir, extra):
    csrc = os.path.join(ROOT, "hpddm_b200", "csrc")
    objs = []
    for sfx, flags in (("d", []), ("z", ["-DHB_COMPLEX"])):
        for src in [os.path.join(csrc, s) for s in SOURCES] + [os.path.join(ROOT, "tests", "native", "device_mock.cpp")]:
            obj = os.path.join(out_dir, os.path.basename(src).rsplit(".", 1)[0] + f".{sfx}.o")
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-I/usr/local/cuda/include", "-I", os.path.join(ROOT, "include")] + extra + flags +
                                  ["-x", "c++", "-c", src, "-o", obj])
            objs.append(obj)
    so = os.path.join(out_dir, "libhpddm_b200_standin.so")
    subprocess.check_call(["g++", "-shared", "-Wl,-Bsymbolic"] + extra + ["-o", so] + objs + ["-ldl"])
    return so


def main(argv):
    asan = "--asan" in argv
    argv = [a for a in argv if a != "--asan"]
    if asan and "libasan" not in os.environ.get("LD_PRELOAD", ""):
        lib = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, LD_PRELOAD=lib, ASAN_OPTIONS="detect_leaks=0")
        raise SystemExit(subprocess.call([sys.executable, os.path.abspath(__file__), "--asan"] + argv, env=env))
    with tempfile.TemporaryDirectory() as tmp:
        so = build(tmp, ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-g", "-O1"] if asan else [])
        os.environ["HPDDM_B200_TEST_STANDIN"] = "1"     # tests/conftest.py: do not skip the gpu-marked tests in this process
        from hpddm_b200 import capi
        capi.LIB_PATH = so
        import pytest
        if not argv:
            argv = DEFAULT + ["-k", " and ".join("not " + t for t in NEEDS_DEVICE)]
        return pytest.main(["-q", "-m", "gpu", "-p", "no:cacheprovider"] + argv)


if __name__ == "__main__":
    raise SystemExit(main(sys.argv[1:]))
