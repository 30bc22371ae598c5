"""On-disk formats of the reference, for cross-run reproducibility of the inputs of the hot path:

* the MatrixCSR text dump written by MatrixBase::dump (include/HPDDM_matrix.hpp:121-135; `-hpddm_dump_matrices`,
  `-hpddm_level_2_dump_matrix`) and read back by MatrixCSR(std::ifstream&) (matrix.hpp:173-244): comment lines start with '#' or '%',
  header "n m nnz" or "n m sym nnz [indexing]", then one "i j a_ij" (1-based) -- or "a_ij i j" -- line per coefficient, in row order;
  complex coefficients are written "(re,im)" (matrix.hpp:105-112).
* the linear-system files of examples/data/40X (examples/driver.cpp:84-114): "n nnz n+1", a[nnz], ja[nnz], ia[n+1] (1-based), rhs[n].

Pure host-side I/O: nothing here touches the GPU or the oracle."""
import re

import numpy as np
import scipy.sparse as sp

_CPLX = re.compile(r"\(\s*([^,()\s]+)\s*,\s*([^,()\s]+)\s*\)")


def _val(tok):
    m = _CPLX.fullmatch(tok)
    return complex(float(m.group(1)), float(m.group(2))) if m else float(tok)


def read_matrix(path):
    """MatrixCSR(std::ifstream&) (matrix.hpp:173-244).  Returns (scipy CSR, sym flag); a symmetric dump holds the lower triangle."""
    n = m = nnz = 0
    sym = False
    rows, cols, vals = [], [], []
    with open(path) as f:
        header = False
        for line in f:
            line = line.strip()
            if not line or line[0] in "#%":
                continue
            tok = line.split()
            if not header:
                if len(tok) == 3:
                    n, m, nnz = int(tok[0]), int(tok[1]), int(tok[2])
                elif len(tok) > 3:
                    n, m, sym, nnz = int(tok[0]), int(tok[1]), bool(int(tok[2])), int(tok[3])
                else:
                    raise ValueError(f"{path}: unsupported header {line!r}")
                header = True
                continue
            if re.fullmatch(r"[+-]?\d+", tok[0]):       # "i j a_ij"
                i, j, v = int(tok[0]), int(tok[1]), _val("".join(tok[2:]))
            else:                                        # "a_ij i j"
                v, i, j = _val("".join(tok[:-2])), int(tok[-2]), int(tok[-1])
            rows.append(i - 1)
            cols.append(j - 1)
            vals.append(v)
    if len(vals) != nnz:
        raise ValueError(f"{path}: header announces {nnz} coefficients, found {len(vals)}")
    dtype = np.complex128 if any(isinstance(v, complex) for v in vals) else np.float64
    A = sp.csr_matrix((np.asarray(vals, dtype=dtype), (rows, cols)), shape=(n, m))
    A.sort_indices()
    return A, sym


def write_matrix(path, A, sym=False, numbering="C"):
    """MatrixBase::dump<N> (matrix.hpp:121-135): same header, comments and coefficient order (row by row, CSR order)."""
    A = sp.csr_matrix(A)
    cplx = np.iscomplexobj(A.data)
    with open(path, "w") as f:
        f.write("# First line: n m (is symmetric) nnz indexing\n")
        f.write("# For each nonzero coefficient: i j a_ij such that (i, j) \\in  {1, ..., n} x {1, ..., m}\n")
        f.write(f"{A.shape[0]} {A.shape[1]} {int(bool(sym))}  {A.nnz} {numbering}\n")
        for i in range(A.shape[0]):
            for k in range(A.indptr[i], A.indptr[i + 1]):
                v = A.data[k]
                txt = f"({v.real:.17e},{v.imag:.17e})" if cplx else f"{v:.17e}"
                f.write(f"{i + 1:9d} {A.indices[k] + 1:9d} {txt}\n")


def read_system(path):
    """examples/data/40X/*.txt (examples/driver.cpp:84-114).  Returns (scipy CSR, rhs)."""
    tok = open(path).read().split()
    n, nnz = int(tok[0]), int(tok[1])
    a = np.array(tok[3:3 + nnz], dtype=np.float64)
    ja = np.array(tok[3 + nnz:3 + 2 * nnz], dtype=np.int32)
    ia = np.array(tok[3 + 2 * nnz:4 + 2 * nnz + n], dtype=np.int32)
    rhs = np.array(tok[4 + 2 * nnz + n:4 + 2 * nnz + 2 * n], dtype=np.float64)
    return sp.csr_matrix((a, ja - 1, ia - 1), shape=(n, n)), rhs
