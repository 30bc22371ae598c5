"""Loader for tests/golden/*.npz (outputs of the unmodified reference, see
oracle/ref_build/make_goldens.py)."""
import glob
import os

import numpy as np
import scipy.sparse as sp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cases():
    """real-scalar goldens first, then the complex ones (K = std::complex<double>)"""
    names = [os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))]
    names = [n for n in names if not n.startswith("refdata_")]   # the reference's own data files (load_refdata)
    return sorted(names, key=lambda n: (n.startswith("complex"), n))


def arg_value(args, key, default):
    toks = args.replace("=", " ").split()
    for i, t in enumerate(toks):
        if t.lstrip("-") == key.lstrip("-") and i + 1 < len(toks):
            return toks[i + 1]
    return default


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    P = int(z["np"])
    args = str(z["args"])
    parts, ref = [], []
    for r in range(P):
        g = {k[len(f"r{r}_"):]: z[k] for k in z.files if k.startswith(f"r{r}_")}
        n, nnz, sym, _ = [int(v) for v in g["header"]]
        Mat = sp.csr_matrix((g["a"], g["ja"], g["ia"]), shape=(n, n))
        mapping = [g[f"mapping{i}"] for i in range(len(g["o"]))]
        g["sol"] = np.asfortranarray(g["sol"].reshape(-1, n).T)     # n x mu, column-major like the reference's buffers
        parts.append(dict(o=[int(v) for v in g["o"]], mapping=mapping, ndof=n, Mat=Mat, sym=bool(sym), d=g["d_ramp"].copy(),
                          f=np.asfortranarray(g["f"].reshape(-1, n).T)))
        ref.append(g)
    meta = dict(P=P, args=args, mu=max(1, int(arg_value(args, "generate_random_rhs", 0))), complex=bool(np.iscomplexobj(ref[0]["a"])), Nx=int(arg_value(args, "Nx", 100)), Ny=int(arg_value(args, "Ny", 100)), overlap=int(arg_value(args, "overlap", 1)),
                sym=arg_value(args, "symmetric_csr", "0") == "1", nu=int(arg_value(args, "deflation_vectors", 0)),
                restart=int(arg_value(args, "hpddm_gmres_restart", 40)), max_it=int(arg_value(args, "hpddm_max_it", 100)),
                tol=float(arg_value(args, "hpddm_tol", 1e-6)), method=arg_value(args, "hpddm_schwarz_method", "ras"),
                krylov=arg_value(args, "hpddm_krylov_method", "gmres"), penalised=arg_value(args, "penalise", "0") == "1",
                recycle=int(arg_value(args, "hpddm_recycle", 0)), solves=int(arg_value(args, "solves", 1)),
                recycle_target=arg_value(args, "hpddm_recycle_target", "SM"), same_system=int(arg_value(args, "hpddm_recycle_same_system", 0)))
    # -solves N (recycling drivers): right-hand side / solution / iteration count of solve s >= 2 as f{s}, sol{s}, iterations{s}
    for r in range(P):
        n = parts[r]["ndof"]
        for s in range(2, meta["solves"] + 1):
            for key in (f"f{s}", f"sol{s}"):
                ref[r][key] = np.asfortranarray(ref[r][key].reshape(-1, n).T)
    return parts, ref, meta


def col(v):
    v = np.asarray(v)
    return np.asfortranarray(v.astype(np.complex128 if np.iscomplexobj(v) else np.float64).reshape(-1, 1))


# the complex goldens come from the reference's FORCE_COMPLEX build with the generator's matrix shifted to
# A - (k^2 - i sigma) I and an imaginary part added to f (oracle/ref_build/ref_driver.cpp)
COMPLEX_SHIFT = complex(-3.0, 1.0)


def complexify(part, rank):
    """Apply ref_driver.cpp's documented complex perturbation to a generate2d() part."""
    A = sp.csr_matrix(part["Mat"]).astype(np.complex128)
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    A.data[A.indices == rows] += COMPLEX_SHIFT
    n = A.shape[0]
    f = part["f"][:, :1].astype(np.complex128)
    f[:, 0] = f[:, 0].real + 1j * 0.25 * np.sin(0.05 * np.arange(n) + rank)
    return dict(part, Mat=A, f=np.asfortranarray(f))


PEN_GOLD = 2.0 ** 100   # the power of two next to HPDDM_PEN = 1e30: keeps b / diag and diag * x exact (see ref_driver.cpp)


def penalise(part, rank, P, Nx, Ny, overlap):
    """ref_driver.cpp's `-penalise 1`: Dirichlet data on the side y = 0 by penalisation (diag = 2^100, f = 2^100 * g),
    local layout of examples/generate.cpp:51-61."""
    xg = int(np.sqrt(P))
    while P % xg:
        xg -= 1
    yg = P // xg
    y, x = divmod(rank, xg)
    i0, i1 = max(x * Nx // xg - overlap, 0), min((x + 1) * Nx // xg + overlap, Nx)
    j0 = max(y * Ny // yg - overlap, 0)
    A = sp.csr_matrix(part["Mat"]).copy()
    f = np.array(part["f"], order="F", copy=True)
    if j0 == 0:
        for i in range(i0, i1):
            k = i - i0
            lo, hi = A.indptr[k], A.indptr[k + 1]
            A.data[lo:hi][A.indices[lo:hi] == k] = PEN_GOLD
            for nu in range(f.shape[1]):
                f[k, nu] = PEN_GOLD * ((1.0 + 0.5 * np.sin(0.3 * i + nu)) + (1j * 0.25 * np.cos(0.2 * i) if np.iscomplexobj(f) else 0.0))
    return dict(part, Mat=A, f=f)


def load_refdata():
    """The reference's in-tree data fixtures, converted by oracle/ref_build/make_data_fixture.py: examples/data/40X (400.txt:
    SPD finite-element system, kappa ~ 1e4, Fortran-numbered CSR + right-hand side; examples/driver.cpp:84-114) and mini.mtx
    (HPDDM matrix dump format, include/HPDDM_matrix.hpp:121-135).  Returns {name: (ia, ja, a, numbering, rhs, scipy matrix)}."""
    out = {}
    z = np.load(os.path.join(GOLDEN_DIR, "refdata_40X_400.npz"))
    n = int(z["n"])
    A = sp.csr_matrix((z["a"], z["ja"] - 1, z["ia"] - 1), shape=(n, n))
    out["40X_400"] = (z["ia"].astype(np.int32), z["ja"].astype(np.int32), z["a"].copy(), "F", z["rhs"].copy(), A)
    m = np.load(os.path.join(GOLDEN_DIR, "refdata_mini_mtx.npz"))
    n = int(m["n"])
    M = sp.csr_matrix((m["v"], (m["i"] - 1, m["j"] - 1)), shape=(n, n))
    M.sort_indices()
    rhs = np.sin(0.1 * np.arange(n)) + 1.0
    out["mini_mtx"] = (M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.copy(), "C", rhs, M)
    return out
