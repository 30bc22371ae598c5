"""Converts two of the reference's own data fixtures (examples/data/40X.tar.gz -> 400.txt, examples/data/mini.tar.gz ->
mini.mtx; SURVEY.md section 8c "Fixtures available in-tree") into tests/golden/refdata_*.npz so that the tests can use
them where /root/reference does not exist (the GPU box).  400.txt format: examples/driver.cpp:84-114 (n, nnz, n+1, a[nnz],
ja[nnz], ia[n+1], rhs[n]; Fortran numbering); mini.mtx: HPDDM's matrix dump format (include/HPDDM_matrix.hpp:121-135,173-244:
"n m nnz" then "row col value", 1-based).

    python oracle/ref_build/make_data_fixture.py
"""
import io
import os
import tarfile

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
    with tarfile.open(os.path.join(DATA, "mini.tar.gz")) as t:
        txt = t.extractfile("mini.mtx").read().decode()
    rows = np.loadtxt(io.StringIO(txt), skiprows=1)
    hdr = [int(v) for v in txt.splitlines()[0].split()]
    np.savez_compressed(os.path.join(OUT, "refdata_mini_mtx.npz"), n=hdr[0], nnz=hdr[2], i=rows[:, 0].astype(np.int32), j=rows[:, 1].astype(np.int32), v=rows[:, 2],
                        source="examples/data/mini.tar.gz:mini.mtx")
    print("400.txt:", n, nnz, " mini.mtx:", hdr)


if __name__ == "__main__":
    main()
