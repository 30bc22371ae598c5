"""CPU study for DESIGN.md section 9 ("Pivoting"): how many digits does a sparse LU WITHOUT row interchanges lose on the config-5
operator as the wavenumber grows?  Proxy for the GPU factorisation (which never pivots): SuperLU with diag_pivot_thresh = 0 (the
diagonal entry is always taken: static pivoting) against SuperLU with threshold pivoting, on one subdomain of the product's own 3-D
Helmholtz generator (-Laplace(u) - k^2 u, first-order absorbing boundary, complex symmetric, indefinite once k^2 exceeds the smallest
Laplace eigenvalue).  Reported: relative residual of one solve and after one step of iterative refinement, pivot growth proxy.
Not a measurement of the GPU code (different ordering, no explicit block inverses) -- a characterisation of the numerical regime.

    python profiles/tools/helmholtz_static_pivoting_study.py [m]
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hpddm_b200.examples.generate import generate_helmholtz3d  # noqa: E402


def main(m):
    print(f"one subdomain of {m}^3 cells on [0,10]^3 (h = {10.0 / m:.3f}); points per wavelength = 2 pi / (k h)")
    print(f"{'k':>6} {'ppw':>6} {'static LU residual':>20} {'+1 refinement':>15} {'pivoted LU residual':>20} {'max|U| / max|A|':>16}")
    rs = np.random.RandomState(0)
    for k in (0.5, 1.0, 2.0, 3.0, 4.0, 6.0, 8.0, 10.0):
        part = generate_helmholtz3d(0, 1, N=(m, m, m), overlap=1, mu=1, k=k)
        A = sp.csc_matrix(part["Mat"])
        n = A.shape[0]
        b = rs.uniform(size=n) + 1j * rs.uniform(size=n)
        out = []
        for thresh in (0.0, 1.0):
            lu = spla.splu(A, permc_spec="COLAMD", diag_pivot_thresh=thresh, options=dict(SymmetricMode=thresh == 0.0))
            x = lu.solve(b)
            r1 = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
            x2 = x + lu.solve(b - A @ x)
            r2 = np.linalg.norm(A @ x2 - b) / np.linalg.norm(b)
            out.append((r1, r2, np.abs(lu.U.data).max() / np.abs(A.data).max(), bool(np.all(lu.perm_r == np.arange(n)) or thresh == 1.0)))
        print(f"{k:6.1f} {2 * np.pi / (k * 10.0 / m):6.1f} {out[0][0]:20.2e} {out[0][1]:15.2e} {out[1][0]:20.2e} {out[0][2]:16.1e}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 20)
