"""GPU parity tests of the complex instantiation (hpddm_b200z_*, include/hpddm_b200z.h; K = std::complex<double>,
BASELINE.json config 5: Helmholtz, ORAS, user-supplied coarse vectors), through the C ABI against the CPU oracle on
the same seeded inputs.  Tolerance: FP64 relative 1e-10, identical Krylov iteration counts.  (The complex outputs of
the unmodified reference's FORCE_COMPLEX build are checked in tests/test_gpu_golden.py.)"""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from hpddm_b200 import Decomposition, KrylovOperator
from hpddm_b200.examples.generate import generate_helmholtz3d
from oracle.krylov import OracleOperator, gmres
from oracle.schwarz import ADDITIVE, BALANCED, DEFLATED, OG, SchwarzWorld
from tests.helpers import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def crand(rs, *shape):
    return rs.uniform(-1, 1, size=shape) + 1j * rs.uniform(-1, 1, size=shape)


@pytest.fixture(scope="module")
def helmholtz():
    """2x2x2 subdomains of a 16^3 Helmholtz grid (k = 2), ORAS with impedance transmission conditions, nu = 4 plane waves."""
    P = 8
    parts = [generate_helmholtz3d(r, P, N=(16, 16, 16), overlap=1, mu=4, k=2.0, nu=4) for r in range(P)]
    w = SchwarzWorld(parts, method=OG)
    w.multiplicity_scaling()
    w.numfact([p["MatRobin"] for p in parts])
    w.set_vectors([p["Z"] for p in parts])
    w.build_coarse()
    deco = Decomposition(0, dtype=np.complex128)
    for r, p in enumerate(parts):
        s = deco.add(r)
        s.initialize(p["Mat"], p["o"], p["mapping"])
        s.setGridHint(*p["dims"])
    ds = deco.multiplicityScaling([p["d"] for p in parts])
    for r in range(P):
        assert np.abs(ds[r] - w.d[r]).max() < 1e-15
    for r, s in enumerate(deco.subs):
        s.callNumfact(A=parts[r]["MatRobin"], method="oras")
        s.setVectors(parts[r]["Z"])
    deco.buildTwo()
    yield parts, w, deco
    deco.close()


def rhs(parts, w, seed, mu=2):
    rs = np.random.RandomState(seed)
    x = [np.asfortranarray(crand(rs, p["ndof"], mu)) for p in parts]
    return w.exchange(x)   # consistent vectors, like examples/schwarz.cpp:98


@pytest.mark.parametrize("n_side,mu", [(6, 1), (14, 1), (14, 3), (14, 4), (14, 7)])
def test_complex_local_solve(n_side, mu):
    """SUBDOMAIN::solve on one Helmholtz subdomain: no-pivot complex LU, small fronts (batched kernel) and, at 14^3,
    fronts > 160 rows (cuSOLVER Zgetrf / cuBLAS Ztrsm, Zgemm, Ztrmm path), 1 / 2 / 4 right-hand sides per sweep."""
    p = generate_helmholtz3d(0, 1, N=(n_side,) * 3, overlap=1, mu=mu, k=1.5)
    deco = Decomposition(0, dtype=np.complex128)
    s = deco.add(0)
    s.initialize(p["Mat"], p["o"], p["mapping"])
    s.setGridHint(*p["dims"])
    s.setScaling(np.ones(p["ndof"]))
    s.callNumfact()
    st = s.statistics()
    assert st["symmetric"] == 0 and st["factor_bytes"] % 32 == 0   # LU: two panel sets of 16-byte scalars
    b = np.asfortranarray(crand(np.random.RandomState(n_side + mu), p["ndof"], mu))
    x = s.solve(b)
    ref = spla.splu(sp.csc_matrix(p["Mat"])).solve(b)
    assert np.abs(x - ref).max() / np.abs(ref).max() < TOL
    assert np.abs(p["Mat"] @ x - b).max() / np.abs(b).max() < 1e-11
    deco.close()


def test_complex_local_solve_algebraic_ordering_2d():
    """no grid hint (algebraic nested dissection) + a matrix that is complex symmetric with strongly complex pivots"""
    n = 40
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n, n))
    A = sp.csr_matrix(sp.kron(sp.eye(n), T) + sp.kron(T, sp.eye(n))).astype(np.complex128) * 16.0
    A = sp.csr_matrix(A + sp.diags(np.full(n * n, -7.0 + 3.0j)))
    deco = Decomposition(0, dtype=np.complex128)
    s = deco.add(0)
    s.initialize(A, [], [])
    s.setScaling(np.ones(n * n))
    s.callNumfact()
    b = np.asfortranarray(crand(np.random.RandomState(3), n * n, 2))
    x = s.solve(b)
    assert np.abs(A @ x - b).max() / np.abs(b).max() < 1e-11
    deco.close()


def test_complex_exchange_gmv_dot(helmholtz):
    parts, w, deco = helmholtz
    rs = np.random.RandomState(11)
    x = [np.asfortranarray(crand(rs, p["ndof"], 3)) for p in parts]
    assert relerr(deco.exchange(x, scaled=True), w.exchange([v.copy() for v in x])) < 1e-14
    assert relerr(deco.exchange(x, scaled=False), w.subdomain_exchange([v.copy() for v in x])) < 1e-14
    assert relerr(deco.GMV(x), w.GMV(x)) < 1e-13
    y = [np.asfortranarray(crand(rs, p["ndof"], 3)) for p in parts]
    got, ref = deco.dot(x, y), w.dot(x, y)     # sum_i d_i conj(x_i) y_i : conjugated on the first argument
    assert got.dtype == np.complex128 and np.abs(got - ref).max() / np.abs(ref).max() < 1e-13
    assert np.abs(deco.dot(x, x).imag).max() < 1e-12 * np.abs(deco.dot(x, x).real).max()


def test_complex_coarse_operator_is_z_hermitian_a_z(helmholtz):
    parts, w, deco = helmholtz
    E = deco.getCoarse()
    assert E.dtype == np.complex128 and E.shape == w.E.shape
    assert np.abs(E - w.E).max() / np.abs(w.E).max() < 1e-12
    rs = np.random.RandomState(5)
    uc = [np.asfortranarray(crand(rs, nu, 2)) for nu in w.nu]
    assert relerr(deco.callSolver(uc), w.call_solver(uc)) < TOL


def test_complex_deflation(helmholtz):
    parts, w, deco = helmholtz
    x = rhs(parts, w, 21)
    assert relerr(deco.deflation(x), w.deflation([v.copy() for v in x])) < TOL


@pytest.mark.parametrize("correction", [None, DEFLATED, ADDITIVE, BALANCED])
@pytest.mark.parametrize("mu", [1, 4])
def test_complex_apply_oras(helmholtz, correction, mu):
    """Schwarz::apply, Prcndtnr::OG branch (schwarz.hpp:531-547) and the three coarse corrections (552-608), K complex"""
    parts, w, deco = helmholtz
    x = rhs(parts, w, 33 + mu, mu=mu)
    assert relerr(deco.apply(x, correction), w.apply(x, correction)) < TOL


@pytest.mark.parametrize("correction", [None, DEFLATED])
def test_complex_gmres_iteration_count_identical(helmholtz, correction):
    """the reference's GMRES (complex Givens rotations, iterative.hpp:690-710) driven by the GPU operator, by the oracle
    operator, and the device-resident driver hpddm_b200z_solve: same iteration counts, same solution"""
    parts, w, deco = helmholtz
    b = w.exchange([p["f"][:, :1].copy() for p in parts])
    it_ref, x_ref, _ = gmres(OracleOperator(w, correction), b)
    it_gpu, x_gpu, _ = gmres(KrylovOperator(deco, correction), b)
    assert 5 < it_ref < 60 and it_gpu == it_ref
    assert relerr(x_gpu, x_ref) < 1e-8
    it_dev, x_dev, res = deco.solve(b, correction=correction)
    assert it_dev == it_ref
    assert relerr(x_dev, x_ref) < 1e-8
    res = w.compute_residual(x_dev, b)
    assert res[0, 1] / res[0, 0] < 1e-5


def test_complex_solve_gevp_is_refused_cleanly(helmholtz):
    parts, w, deco = helmholtz
    from hpddm_b200 import capi
    with pytest.raises(capi.HpddmB200Error, match="real scalars only"):
        deco.subs[0].solveGEVP(parts[0]["Mat"], nu=2)


def test_real_and_complex_contexts_coexist():
    """both scalar instantiations live in one process / one library"""
    from oracle.generate import generate_world
    parts = generate_world(1, dim=3, N=(8, 8, 8), overlap=1, mu=1)
    dr = Decomposition(0)
    dz = Decomposition(0, dtype=np.complex128)
    sr, sz = dr.add(0), dz.add(0)
    A = sp.csr_matrix(parts[0]["Mat"])
    sr.initialize(A, [], [])
    sz.initialize(A.astype(np.complex128) * (1.0 + 0.5j), [], [])
    for s in (sr, sz):
        s.setScaling(np.ones(A.shape[0]))
        s.callNumfact()
    b = np.random.RandomState(0).uniform(size=(A.shape[0], 1))
    xr, xz = sr.solve(b), sz.solve(b.astype(np.complex128))
    assert np.abs(xz * (1.0 + 0.5j) - xr).max() / np.abs(xr).max() < 1e-12
    dr.close()
    dz.close()


# ---------------------------------------------------------------- size-independent properties at a size the oracle cannot reach quickly
@pytest.fixture(scope="module")
def big_helmholtz():
    m = 56
    part = generate_helmholtz3d(0, 1, N=(m, m, m), overlap=1, mu=4, k=2.0, nu=12)
    deco = Decomposition(0, dtype=np.complex128)
    s = deco.add(0)
    s.initialize(part["Mat"], part["o"], part["mapping"])
    s.setGridHint(*part["dims"])
    deco.multiplicityScaling([part["d"]])
    s.callNumfact()
    s.setVectors(part["Z"])
    deco.buildTwo()
    yield part, deco, s
    deco.close()


def test_complex_solve_round_trip_at_scale(big_helmholtz):
    part, deco, s = big_helmholtz
    b = part["f"]                                      # 4 complex right-hand sides: one pass over the panels
    x = s.solve(b)
    r = np.linalg.norm(part["Mat"] @ x - b, axis=0) / np.linalg.norm(b, axis=0)
    assert r.max() < 1e-11
    st = s.statistics()
    assert st["symmetric"] == 0 and st["nnz_factor"] > 5e7


def test_complex_apply_is_complex_linear_and_blocks_equal_columns(big_helmholtz):
    part, deco, s = big_helmholtz
    rs = np.random.RandomState(1)
    x, y = (np.asfortranarray(crand(rs, part["ndof"], 1)) for _ in range(2))
    a, b = 2.5 - 1.5j, -0.75 + 0.25j
    for corr in (None, DEFLATED, ADDITIVE, BALANCED):
        mx, my = deco.apply([x], corr)[0], deco.apply([y], corr)[0]
        mz = deco.apply([a * x + b * y], corr)[0]
        assert np.abs(mz - (a * mx + b * my)).max() / np.abs(mz).max() < 1e-11      # linear over C (no stray conjugation)
        blk = deco.apply([np.asfortranarray(np.hstack([x, y, x + y, 1j * x]))], corr)[0]
        assert np.abs(blk[:, :1] - mx).max() / np.abs(mx).max() < 1e-11
        assert np.abs(blk[:, 2:3] - (mx + my)).max() / np.abs(mx).max() < 1e-11
        assert np.abs(blk[:, 3:] - 1j * mx).max() / np.abs(mx).max() < 1e-11


def test_complex_two_level_apply_of_A_times_coarse_vector_returns_it(big_helmholtz):
    """M^-1 A z = z for z in the coarse space with the deflated correction: involves Z^H (conjugated), E^-1, Z, SpMV and the solve"""
    part, deco, s = big_helmholtz
    z = np.asfortranarray(part["Z"] @ crand(np.random.RandomState(2), part["Z"].shape[1], 1))
    Az = deco.GMV([z])[0]
    back = deco.apply([Az], DEFLATED)[0]
    assert np.abs(back - z).max() / np.abs(z).max() < 1e-9
