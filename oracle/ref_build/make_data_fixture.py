"""Converts two of the reference's own data fixtures (examples/data/40X.tar.gz -> 400.txt, examples/data/mini.tar.gz ->
mini.mtx; SURVEY.md section 8c "Fixtures available in-tree") into tests/golden/refdata_*.npz so that the tests can use
them where /root/reference does not exist (the GPU box).  400.txt format: examples/driver.cpp:84-114 (n, nnz, n+1, a[nnz],
ja[nnz], ia[n+1], rhs[n]; Fortran numbering); mini.mtx: HPDDM's matrix dump format (include/HPDDM_matrix.hpp:121-135,173-244:
"n m nnz" then "row col value", 1-based).

The whole 40X sequence (400.txt ... 409.txt: one sparsity pattern, ten slowly drifting value sets and right-hand sides) goes into
refdata_40X_sequence.npz together with the iteration counts of the UNMODIFIED reference driver on it (oracle/_ref/driver_ref =
examples/driver.cpp, GCRO-DR(40, 20), tol 1e-10, with and without -diagonal_scaling): the reference's own known-answer test of its
recycling Krylov method (Makefile:380; pass window on the total, examples/driver.cpp:152-155).

    python oracle/ref_build/make_data_fixture.py
"""
import io
import os
import re
import subprocess
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
DATA = "/root/reference/examples/data"
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    with tarfile.open(os.path.join(DATA, "40X.tar.gz")) as t:
        tok = t.extractfile("400.txt").read().split()
    n, nnz = int(tok[0]), int(tok[1])
    a = np.array(tok[3:3 + nnz], dtype=np.float64)
    ja = np.array(tok[3 + nnz:3 + 2 * nnz], dtype=np.int32)
    ia = np.array(tok[3 + 2 * nnz:4 + 2 * nnz + n], dtype=np.int32)
    rhs = np.array(tok[4 + 2 * nnz + n:4 + 2 * nnz + 2 * n], dtype=np.float64)
    assert ia[0] == 1 and ia[-1] == nnz + 1 and rhs.size == n
    np.savez_compressed(os.path.join(OUT, "refdata_40X_400.npz"), n=n, ia=ia, ja=ja, a=a, rhs=rhs, numbering="F", source="examples/data/40X.tar.gz:400.txt")
    # the whole sequence + what the reference's own driver does with it
    systems = []
    with tarfile.open(os.path.join(DATA, "40X.tar.gz")) as t, tempfile.TemporaryDirectory() as tmp:
        for i in range(10):
            raw = t.extractfile(f"40{i}.txt").read()
            open(os.path.join(tmp, f"40{i}.txt"), "wb").write(raw)
            tok = raw.split()
            assert int(tok[0]) == n and int(tok[1]) == nnz
            assert np.array_equal(np.array(tok[3 + nnz:3 + 2 * nnz], dtype=np.int32), ja) and np.array_equal(np.array(tok[3 + 2 * nnz:4 + 2 * nnz + n], dtype=np.int32), ia)
            systems.append((np.array(tok[3:3 + nnz], dtype=np.float64), np.array(tok[4 + 2 * nnz + n:4 + 2 * nnz + 2 * n], dtype=np.float64)))
        drv = os.path.join(ROOT, "oracle", "_ref", "driver_ref")
        counts = {}
        for key, extra in (("plain", []), ("diagonal_scaling", ["-diagonal_scaling", "1"]), ("block", [])):
            method = "bgcrodr" if key == "block" else "gcrodr"   # (the pattern below also matches "BGCRODR converges ...")
            out = subprocess.run([drv, f"-path={tmp}", "-hpddm_krylov_method", method, "-hpddm_verbosity", "1"] + extra, env=dict(os.environ, HPDDM_SHIM_NP="1"),
                                 capture_output=True, text=True, timeout=600)
            counts[key] = [int(v) for v in re.findall(r"GCRODR converges after (\d+) iteration", out.stdout)]
            total = int(re.search(r"Total number of iterations: (\d+)", out.stdout).group(1))
            assert out.returncode == 0 and len(counts[key]) == 10 and sum(counts[key]) == total, out.stdout[-500:]   # returncode 0 = inside the reference's window
            print("driver_ref", method, key, counts[key], total)
    np.savez_compressed(os.path.join(OUT, "refdata_40X_sequence.npz"), n=n, ia=ia, ja=ja, a=np.array([s_[0] for s_ in systems]), rhs=np.array([s_[1] for s_ in systems]),
                        numbering="F", gcrodr_40_20_tol1e10=np.array(counts["plain"]), gcrodr_40_20_tol1e10_diagonal_scaling=np.array(counts["diagonal_scaling"]), bgcrodr_40_20_tol1e10=np.array(counts["block"]),
                        source="examples/data/40X.tar.gz:400.txt-409.txt; counts: unmodified examples/driver.cpp (oracle/_ref/driver_ref)")
    with tarfile.open(os.path.join(DATA, "mini.tar.gz")) as t:
        txt = t.extractfile("mini.mtx").read().decode()
    rows = np.loadtxt(io.StringIO(txt), skiprows=1)
    hdr = [int(v) for v in txt.splitlines()[0].split()]
    np.savez_compressed(os.path.join(OUT, "refdata_mini_mtx.npz"), n=hdr[0], nnz=hdr[2], i=rows[:, 0].astype(np.int32), j=rows[:, 1].astype(np.int32), v=rows[:, 2],
                        source="examples/data/mini.tar.gz:mini.mtx")
    print("400.txt:", n, nnz, " mini.mtx:", hdr)


if __name__ == "__main__":
    main()
