"""CPU-side checks: the C-ABI library loads, exports every symbol include/hpddm_b200.h and
include/hpddm_b200z.h (complex instantiation) declare, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from hpddm_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    out = set()
    for header in ("hpddm_b200.h", "hpddm_b200z.h"):
        src = open(os.path.join(ROOT, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        out |= set(re.findall(r"\b(hpddm_b200z?_[a-z0-9_]+)\s*\(", src))
    return sorted(out)


def test_complex_header_mirrors_real_header():
    names = declared_symbols()
    real = {n[len("hpddm_b200_"):] for n in names if n.startswith("hpddm_b200_")}
    cplx = {n[len("hpddm_b200z_"):] for n in names if n.startswith("hpddm_b200z_")}
    assert real == cplx and len(real) > 30


def test_header_and_binding_agree():
    assert declared_symbols() == capi.EXPORTS


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert b"sm_100a" in L.hpddm_b200_version() and b"complex" in L.hpddm_b200z_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = capi.lib().hpddm_b200_ctx_create(0, C.byref(h))
    assert rc < 0
    assert b"no CPU fallback" in capi.lib().hpddm_b200_last_error()
    with pytest.raises(capi.HpddmB200Error):
        capi.check(rc)
    rc = capi.COMPLEX.ctx_create(0, C.byref(h))
    assert rc < 0 and b"no CPU fallback" in capi.COMPLEX.last_error()
    with pytest.raises(capi.HpddmB200Error):
        capi.COMPLEX.check(rc)


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = "import sys; import hpddm_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hpddm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_host_cpp_header_compiles_and_links():
    """hpddm_b200/host/HPDDM_B200.hpp: every member of B200Sub / B200Schwarz instantiates and the
    symbols resolve against libhpddm_b200.so (no reference tree needed)."""
    import subprocess
    import tempfile
    src = os.path.join(ROOT, "tests", "native", "test_host_header.cpp")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "t")
        subprocess.check_call(["g++", "-std=c++11", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "hpddm_b200", "host"), src,
                               "-L", os.path.join(ROOT, "hpddm_b200", "lib"), "-lhpddm_b200", "-Wl,-rpath," + os.path.join(ROOT, "hpddm_b200", "lib"), "-o", exe])
        subprocess.check_call([exe])


@pytest.mark.parametrize("flags", [[], ["-DFORCE_COMPLEX"]])
def test_full_seam_instantiates_its_device_resident_solve(flags, tmp_path):
    """hpddm_b200/host/HPDDM_B200_schwarz.hpp against the reference's headers: the unmodified driver never calls solveOnDevice (it goes
    through IterativeMethod::solve), so this member -- the device-resident GMRES / BGMRES / CG / GCRO-DR dispatch on -hpddm_krylov_method --
    is instantiated here, for both scalar types.  Needs the reference tree (absent on the GPU box: skipped there)."""
    import subprocess
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "include")):
        pytest.skip("reference tree not present")
    src = tmp_path / "seam_inst.cpp"
    src.write_text('#include "schwarz.hpp"\n'
                   "template <class T> int inst(T &A) { K *f = nullptr, *x = nullptr; A.destroy(); return A.solveOnDevice(f, x, 1); }\n"
                   "int main() { volatile bool run = false; if (run) { HPDDM::Schwarz<SUBDOMAIN, COARSEOPERATOR, symCoarse, K> A; return inst(A); } return 0; }\n")
    host = os.path.join(ROOT, "hpddm_b200", "host")
    subprocess.check_call(["g++", "-O0", "-std=c++11", "-w", "-DDLAPACK", "-DB200SUB", "-DB200SCHWARZ", "-DGENERAL_CO", "-DHPDDM_NUMBERING='C'"] + flags +
                          ["-I", os.path.join(ROOT, "oracle", "ref_build"), "-I", os.path.join(ref, "include"), "-I", os.path.join(ref, "examples"), "-I", os.path.join(ROOT, "include"),
                           "-I", host, "-include", os.path.join(host, "HPDDM_B200.hpp"), "-include", os.path.join(host, "HPDDM_B200_schwarz.hpp"), "-c", str(src),
                           "-o", str(tmp_path / "seam_inst.o")])


def test_ctypes_signatures_match_the_headers():
    """every declaration of include/hpddm_b200.h / hpddm_b200z.h: same number of arguments and same scalar/pointer kind as
    the argtypes of hpddm_b200/capi.py (a wrong binding would corrupt the stack silently)"""
    import ctypes as C
    for header in ("hpddm_b200.h", "hpddm_b200z.h"):
        src = open(os.path.join(ROOT, "include", header)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"\b(hpddm_b200z?_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
            name, args = m.group(1), m.group(2).strip()
            if name not in capi._SIGS:
                continue
            params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
            _, argtypes = capi._SIGS[name]
            assert len(params) == len(argtypes), (name, params, argtypes)
            for prm, at in zip(params, argtypes):
                is_ptr = "*" in prm or "_fn " in prm   # (function-pointer typedefs, e.g. hpddm_b200_allgather_fn)
                if is_ptr:
                    assert at in (C.c_void_p, C.c_char_p) or hasattr(at, "contents") or at is capi._P, (name, prm, at)
                elif prm.startswith("double"):
                    assert at is C.c_double, (name, prm, at)
                elif prm.startswith("char"):
                    assert at is C.c_char, (name, prm, at)
                elif prm.startswith("size_t"):
                    assert at is C.c_size_t, (name, prm, at)
                else:
                    assert at is C.c_int, (name, prm, at)


@pytest.mark.parametrize("scalar", ["real", "complex"])
def test_symbolic_phase_and_work_item_tiling_on_the_host(scalar, tmp_path):
    """hb_symbolic.cpp on the host (tests/native/test_symbolic.cpp): valid permutation, level property of the assembly tree, a dense
    multifrontal Cholesky driven by the same structures succeeds, and the forward / backward SpTRSV work items tile every stored panel
    entry exactly once -- for both scalar builds (the chunk widths FCH / BCH depend on sizeof(K)), geometric and algebraic ordering."""
    import subprocess
    exe = str(tmp_path / "tsym")
    cmd = ["nvcc", "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "hpddm_b200", "csrc"), "-x", "cu",
           "-o", exe, os.path.join(ROOT, "tests", "native", "test_symbolic.cpp"), os.path.join(ROOT, "hpddm_b200", "csrc", "hb_symbolic.cpp")]
    if scalar == "complex":
        cmd.insert(1, "-DHB_COMPLEX")
    subprocess.check_call(cmd)
    for args in (["12", "12", "12", "1"], ["10", "8", "6", "0"], ["24", "24", "24", "1"], ["30", "30", "1", "0"], ["9", "7", "5", "1", "8"]):
        out = subprocess.run([exe] + args, capture_output=True, text=True)
        assert out.returncode == 0 and "factor ok" in out.stdout, (args, out.stdout[-300:])


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs next to the CUDA arm): one JSON line with the contract's keys"""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "24", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-1000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["value"] > 0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]
    # both arms describe the SAME workload: the config object is a function of (cells, gpus, nu, mu) only and the CPU sample is never shrunk
    import bench
    wl, par = bench.workload_name(24, 1, 20, 1)
    assert d["config"] == {"workload": wl, "parallelism": par, "cells": 24, "nu": 20, "mu": 1}
    assert d["cpu_baseline"]["cells"] == 24 and "24^3" in d["cpu_baseline"]["sample"]


def test_bench_default_size_rule_is_shared_by_both_arms(monkeypatch):
    import bench
    monkeypatch.delenv("HPDDM_B200_BENCH_M", raising=False)
    monkeypatch.setattr(bench, "host_memory_bytes", lambda: 200e9)
    assert bench.default_cells() == 160
    monkeypatch.setattr(bench, "host_memory_bytes", lambda: 100e9)
    assert bench.default_cells() == 128
    monkeypatch.setenv("HPDDM_B200_BENCH_M", "96")
    assert bench.default_cells() == 96


def test_full_seam_binary_is_built_from_the_unmodified_driver_and_fails_loudly_without_a_gpu(tmp_path):
    """oracle/_ref/schwarz_b200_full = the reference's examples/schwarz.cpp compiled against HPDDM::Schwarz<HPDDM::B200Sub, ...>
    (hpddm_b200/host/HPDDM_B200_schwarz.hpp).  Without a CUDA device it must stop with the library's "no CPU fallback" error, never
    compute on the host."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "schwarz_b200_full")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/schwarz_b200_full not built (needs /root/reference at build time)")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: the GPU suite runs this binary for real")
    except ImportError:
        pass
    res = subprocess.run([exe, "-hpddm_verbosity=1", "-Nx", "20", "-Ny", "20"], env=dict(os.environ, HPDDM_SHIM_NP="2"), cwd=tmp_path, capture_output=True, text=True, timeout=120)
    out = res.stdout + res.stderr
    assert res.returncode != 0 and "no CPU fallback" in out, out[-1500:]
    assert "converges after" not in out
