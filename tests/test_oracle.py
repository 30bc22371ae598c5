"""Oracle self-consistency (CPU): the restated algorithm satisfies the properties the
reference's own drivers test (examples/schwarz.cpp:140-144,178) and basic identities."""
import numpy as np

from oracle.generate import generate2d, generate_world
from oracle.krylov import OracleOperator, gmres
from oracle.schwarz import ADDITIVE, BALANCED, DEFLATED, SchwarzWorld


def test_generate2d_matches_reference_sizes():
    # n_loc = 51^2 = 2601, nnz = 12801, maps (102, 102, 4)
    p = generate2d(0, 4, Nx=100, Ny=100, overlap=1)
    assert p["ndof"] == 2601 and p["Mat"].nnz == 12801
    assert sorted(len(m) for m in p["mapping"]) == [4, 102, 102]
    assert p["o"] == [1, 2, 3]
    ps = generate2d(0, 4, Nx=100, Ny=100, overlap=1, sym=True)
    assert ps["Mat"].nnz == 2601 * 3 - 51 - 51


def test_partition_of_unity_and_consistency():
    for dim, kw in ((2, dict(Nx=40, Ny=30, overlap=2)), (3, dict(N=(12, 10, 8), overlap=1))):
        parts = generate_world(8 if dim == 3 else 6, dim=dim, mu=1, **kw)
        w = SchwarzWorld(parts)
        w.multiplicity_scaling()
        ones = [np.ones((n, 1)) for n in w.n]
        w.exchange(ones)
        assert max(np.abs(o - 1).max() for o in ones) < 1e-14


def test_config1_one_level_ras_converges_like_the_reference_test():
    parts = generate_world(4, dim=2, Nx=100, Ny=100, overlap=1, mu=0)
    w = SchwarzWorld(parts)
    w.multiplicity_scaling()
    w.numfact()
    b = [p["f"].copy() for p in parts]
    it, x, applies = gmres(OracleOperator(w), b, restart=25, max_it=80)
    res = w.compute_residual(x, b)
    assert it <= 45 and res[0, 1] / res[0, 0] <= 1e-2   # examples/schwarz.cpp:140-144
    assert applies == it + 2


def test_two_level_reduces_iterations_3d():
    parts = generate_world(8, dim=3, N=(16, 16, 16), overlap=1, mu=1, neumann=True)
    w = SchwarzWorld(parts)
    w.multiplicity_scaling()
    w.numfact()
    b = w.exchange([p["f"].copy() for p in parts])
    it1, _, _ = gmres(OracleOperator(w), b)
    w.solve_gevp([p["MatNeumann"] for p in parts], nu=4)
    E = w.build_coarse()
    assert np.abs(E - E.T).max() < 1e-10 * np.abs(E).max()
    for c in (DEFLATED, BALANCED, ADDITIVE):
        it2, x, _ = gmres(OracleOperator(w, c), b)
        assert it2 <= it1
        res = w.compute_residual(x, b)
        assert res[0, 1] / res[0, 0] < 1e-5
    # deflated apply: Q is a projection-like operator: Z^T D A (Q v) = Z^T D v on consistent vectors
    v = w.exchange([np.random.RandomState(r).standard_normal((w.n[r], 1)) for r in range(w.P)])
    q = w.deflation(v)
    Aq = w.GMV(q)
    lhs = np.concatenate([w.Z[r].T @ (w.d[r][:, None] * Aq[r]) for r in range(w.P)])
    rhs = np.concatenate([w.Z[r].T @ (w.d[r][:, None] * v[r]) for r in range(w.P)])
    assert np.abs(lhs - rhs).max() < 1e-9 * np.abs(rhs).max()


def test_cpu_benchmark_arm_matches_the_oracle():
    """oracle/cpu_ras.cpp (the CPU arm timed by bench.py) computes the same deflated apply."""
    import bench
    from hpddm_b200.examples.generate import generate3d
    L = bench._cpu_lib()
    m, nu = 14, 6
    part = generate3d(0, 1, N=(m, m, m), overlap=1, mu=1, grid=(1, 1, 1))
    Z = bench.cosine_modes(part["dims"], nu)
    h = L.cpu_ras_create(m, nu, Z.ctypes.data, 2)
    x = part["f"][:, 0].copy()
    y = np.empty_like(x)
    L.cpu_ras_apply(h, x.ctypes.data, y.ctypes.data)
    L.cpu_ras_destroy(h)
    w = SchwarzWorld([part])
    w.multiplicity_scaling()
    w.numfact()
    w.set_vectors([Z])
    w.build_coarse()
    ref = w.apply([x.reshape(-1, 1)], DEFLATED)[0][:, 0]
    assert np.abs(ref - y).max() / np.abs(ref).max() < 1e-12


def test_reference_data_fixtures_load_and_solve():
    """tests/golden/refdata_*.npz (the reference's examples/data files, converted by oracle/ref_build/make_data_fixture.py)"""
    import scipy.sparse as sp
    from oracle.schwarz import LocalSolver
    from tests.golden_util import load_refdata
    data = load_refdata()
    assert data["40X_400"][5].shape == (3988, 3988) and data["40X_400"][5].nnz == 53608      # "3988 53608 3989" header of 400.txt
    assert data["mini_mtx"][5].shape == (976, 976) and data["mini_mtx"][5].nnz == 6526
    for name, (ia, ja, a, numbering, b, A) in data.items():
        x = LocalSolver(A).solve(b)
        assert np.abs(A @ x - b).max() / np.abs(b).max() < 1e-10


def test_reference_on_disk_formats_round_trip(tmp_path):
    """hpddm_b200/io.py: the MatrixCSR dump format (include/HPDDM_matrix.hpp:121-135 writer, 173-244 reader) -- both coefficient
    orders, comments, 3- and 5-field headers, complex "(re,im)" values -- and the examples/data/40X system format."""
    import scipy.sparse as sp
    from hpddm_b200.io import read_matrix, read_system, write_matrix
    A = sp.random(30, 30, density=0.2, random_state=3, format="csr") + sp.eye(30)
    write_matrix(tmp_path / "a.txt", A)
    B, sym = read_matrix(tmp_path / "a.txt")
    assert not sym and abs(A - B).max() == 0.0                      # 17 significant digits: exact round trip
    C = sp.csr_matrix(A + 1j * A.T)
    write_matrix(tmp_path / "c.txt", C, numbering="F")
    D, _ = read_matrix(tmp_path / "c.txt")
    assert D.dtype == np.complex128 and abs(C - D).max() == 0.0
    # what MatrixBase::dump prints (std::scientific, setw(9)) and the alternative "a_ij i j" order with a 3-field header
    (tmp_path / "d.txt").write_text("# First line: n m (is symmetric) nnz indexing\n2 2 1  2 C\n        1         1 4.000000e+00\n        2         1 -1.500000e+00\n")
    E, sym = read_matrix(tmp_path / "d.txt")
    assert sym and E.toarray().tolist() == [[4.0, 0.0], [-1.5, 0.0]]
    (tmp_path / "e.txt").write_text("% comment\n2 2 2\n4.0e+00 1 1\n-1.5e+00 2 1\n")
    F, _ = read_matrix(tmp_path / "e.txt")
    assert abs(E - F).max() == 0.0
    # examples/driver.cpp:84-114
    n, ia, ja, a, rhs = 3, [1, 3, 5, 6], [1, 2, 1, 2, 3], [2.0, -1.0, -1.0, 2.0, 1.0], [1.0, 0.0, 3.0]
    (tmp_path / "s.txt").write_text(f"{n} {len(a)} {n + 1}\n" + " ".join(f"{v:.17E}" for v in a) + "\n" + " ".join(map(str, ja)) + "\n" + " ".join(map(str, ia)) + "\n" + " ".join(f"{v:.17E}" for v in rhs) + "\n")
    S, r = read_system(tmp_path / "s.txt")
    assert S.toarray().tolist() == [[2.0, -1.0, 0.0], [-1.0, 2.0, 0.0], [0.0, 0.0, 1.0]] and r.tolist() == rhs


def test_elasticity_generator_local_assembly_equals_restriction_of_the_global_matrix():
    """generate_elasticity3d(assembly="local") (scalable: only the elements touching the subdomain) against the restriction of the
    globally assembled matrix, every rank of two decompositions, including the penalised clamped face"""
    from hpddm_b200.examples.generate import generate_elasticity3d
    for size, grid, Nn, ov in ((4, (2, 2, 1), (9, 9, 6), 1), (8, (2, 2, 2), (8, 7, 9), 2)):
        for r in range(size):
            g = generate_elasticity3d(r, size, Nn=Nn, overlap=ov, mu=2, grid=grid, assembly="global")
            l = generate_elasticity3d(r, size, Nn=Nn, overlap=ov, mu=2, grid=grid, assembly="local")
            assert g["ndof"] == l["ndof"] and np.array_equal(g["Mat"].indptr, l["Mat"].indptr) and np.array_equal(g["Mat"].indices, l["Mat"].indices)
            scale = np.abs(g["Mat"].data[np.abs(g["Mat"].data) < 1e29]).max()
            assert np.abs(g["Mat"].data - l["Mat"].data).max() <= 1e-13 * scale or np.allclose(g["Mat"].data, l["Mat"].data, rtol=1e-13, atol=1e-13 * scale)
            assert all(np.array_equal(a, b) for a, b in zip(g["mapping"], l["mapping"])) and g["o"] == l["o"]


def test_elasticity_neumann_matrix_annihilates_the_rigid_body_modes():
    """consistency of the Q1 elasticity generator: on a floating subdomain (no clamped face) the Neumann matrix has exactly the six
    rigid-body modes in its kernel; on the clamped one the penalised matrix is positive definite"""
    import scipy.sparse.linalg as spla
    from hpddm_b200.examples.generate import generate_elasticity3d, rigid_body_modes
    Nn = (9, 8, 7)
    parts = [generate_elasticity3d(r, 4, Nn=Nn, overlap=1, mu=1, grid=(2, 2, 1), neumann=True, assembly="local") for r in range(4)]
    floating = [p for p in parts if p["box"][0][0] > 0]
    assert len(floating) == 2
    for p in floating:
        A = p["MatNeumann"]
        Z = rigid_body_modes(p, Nn)
        assert Z.shape == (p["ndof"], 6) and np.linalg.matrix_rank(Z) == 6
        scale = abs(A).max()
        assert np.abs(A @ Z).max() < 1e-12 * scale                        # A_Neu * RBM = 0
        w = np.linalg.eigvalsh(A.toarray())
        assert np.sum(np.abs(w) < 1e-10 * scale) == 6 and w.min() > -1e-10 * scale
    clamped = [p for p in parts if p["box"][0][0] == 0][0]
    lu = spla.splu(clamped["MatNeumann"].tocsc())                        # penalised rows: non-singular
    x = lu.solve(np.ones(clamped["ndof"]))
    assert np.all(np.isfinite(x))
