"""GPU parity tests: every hot-path entry point of libhpddm_b200.so, called through
the C ABI (ctypes), against the CPU oracle on the same seeded inputs.

Tolerance: FP64 relative 1e-10 (BASELINE.json north_star: <= 1e-10 on the
preconditioned residual, identical Krylov iteration count).
"""
import numpy as np
import pytest

from oracle.generate import generate_world
from oracle.krylov import OracleOperator, gmres
from oracle.schwarz import ADDITIVE, BALANCED, DEFLATED, SchwarzWorld
from hpddm_b200 import KrylovOperator
from tests.helpers import build_gpu_decomposition, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def make_world(dim, size, nu=0, mu=2, **kw):
    parts = generate_world(size, dim=dim, mu=mu, neumann=nu > 0, **kw)
    w = SchwarzWorld(parts)
    w.multiplicity_scaling()
    w.numfact()
    if nu > 0:
        w.solve_gevp([p["MatNeumann"] for p in parts], nu=nu)
        w.build_coarse()
    return parts, w


@pytest.fixture(scope="module")
def poisson3d():
    parts, w = make_world(3, 8, nu=4, mu=3, N=(14, 14, 14), overlap=1)
    deco = build_gpu_decomposition(parts, w, two_level=True)
    yield parts, w, deco
    deco.close()


def rhs(parts, w, seed=0):
    rs = np.random.RandomState(seed)
    x = [np.asfortranarray(rs.standard_normal(p["f"].shape)) for p in parts]
    return x


def test_local_solve(poisson3d):
    parts, w, deco = poisson3d
    for r, s in enumerate(deco.subs):
        b = parts[r]["f"]
        got = s.solve(b)
        ref = w.solver[r].solve(b)
        assert relerr([got], [ref]) < TOL
        st = s.statistics()
        assert st["symmetric"] == 1 and st["nnz_factor"] > 0


def test_exchange_scaled_and_unscaled(poisson3d):
    parts, w, deco = poisson3d
    x = rhs(parts, w, 1)
    ref = w.exchange([v.copy() for v in x])
    assert relerr(deco.exchange(x, scaled=True), ref) < 1e-14
    ref = w.subdomain_exchange([v.copy() for v in x])
    assert relerr(deco.exchange(x, scaled=False), ref) < 1e-14


def test_multiplicity_scaling_matches_oracle():
    parts, w = make_world(3, 8, N=(10, 10, 10), overlap=2)
    deco = build_gpu_decomposition(parts, None, own_scaling=True)
    ds = deco.multiplicityScaling([p["d"] for p in parts])
    for r in range(8):
        assert np.abs(ds[r] - w.d[r]).max() < 1e-15
    deco.close()


def test_gmv(poisson3d):
    parts, w, deco = poisson3d
    x = rhs(parts, w, 2)
    assert relerr(deco.GMV(x), w.GMV(x)) < 1e-13


def test_coarse_operator_assembly(poisson3d):
    parts, w, deco = poisson3d
    E = deco.getCoarse()
    assert E.shape == w.E.shape
    assert np.abs(E - w.E).max() / np.abs(w.E).max() < 1e-12


def test_coarse_solve(poisson3d):
    parts, w, deco = poisson3d
    rs = np.random.RandomState(3)
    uc = [np.asfortranarray(rs.standard_normal((w.nu[r], 2))) for r in range(w.P)]
    ref = w.call_solver([u.copy() for u in uc])
    got = deco.callSolver(uc)
    assert relerr(got, ref) < 1e-11


def test_deflation(poisson3d):
    parts, w, deco = poisson3d
    x = rhs(parts, w, 4)
    assert relerr(deco.deflation(x), w.deflation(x)) < TOL


@pytest.mark.parametrize("correction", [None, DEFLATED, ADDITIVE, BALANCED])
def test_apply(poisson3d, correction):
    parts, w, deco = poisson3d
    x = rhs(parts, w, 5)
    ref = w.apply(x, correction)
    got = deco.apply(x, correction)
    assert relerr(got, ref) < TOL


def test_start_and_dot(poisson3d):
    parts, w, deco = poisson3d
    b = rhs(parts, w, 6)
    x0 = rhs(parts, w, 7)
    ref = w.start(b, [v.copy() for v in x0])
    got = deco.start(b, x0)
    deco.end()
    assert relerr(got, ref) < 1e-14
    assert np.abs(deco.dot(b, x0) - w.dot(b, x0)).max() / np.abs(w.dot(b, x0)).max() < 1e-12


@pytest.mark.parametrize("correction", [None, DEFLATED])
def test_gmres_iteration_count_identical(poisson3d, correction):
    """north_star: identical Krylov iteration count, preconditioned residual parity."""
    parts, w, deco = poisson3d
    b = w.exchange([p["f"].copy() for p in parts])
    it_ref, x_ref, ap_ref = gmres(OracleOperator(w, correction), b)
    it_gpu, x_gpu, ap_gpu = gmres(KrylovOperator(deco, correction), b)
    assert it_gpu == it_ref and ap_gpu == ap_ref
    assert relerr(x_gpu, x_ref) < 1e-8
    res = w.compute_residual(x_gpu, b)
    assert np.all(res[:, 1] / res[:, 0] < 1e-5)


def test_config1_2d_quirk_matrix_lu_path():
    """BASELINE config 1: examples/generate.cpp 2-D Poisson 100x100, 4 ranks, overlap 1, one-level RAS.
    The generator's skewed stencil makes the local matrices non-symmetric -> LU panels."""
    parts, w = make_world(2, 4, mu=0, Nx=100, Ny=100, overlap=1)
    deco = build_gpu_decomposition(parts, w)
    assert deco.subs[0].statistics()["symmetric"] == 0
    x = [p["f"].copy() for p in parts]
    assert relerr(deco.apply(x, None), w.apply(x, None)) < TOL
    b = [p["f"].copy() for p in parts]
    it_ref, _, _ = gmres(OracleOperator(w), b, restart=25, max_it=80)
    it_gpu, xg, _ = gmres(KrylovOperator(deco), b, restart=25, max_it=80)
    assert it_gpu == it_ref and it_gpu <= 45           # examples/schwarz.cpp:140
    res = w.compute_residual(xg, b)
    assert res[0, 1] / res[0, 0] < 1e-2                # examples/schwarz.cpp:143
    deco.close()


def test_algebraic_ordering_without_grid_hint():
    parts, w = make_world(3, 2, mu=1, N=(12, 10, 9), overlap=1)
    deco = build_gpu_decomposition(parts, w, grid_hint=False)
    x = [p["f"].copy() for p in parts]
    assert relerr(deco.apply(x, None), w.apply(x, None)) < TOL
    deco.close()


def test_symmetric_csr_input():
    parts, w = make_world(3, 2, mu=1, N=(10, 10, 10), overlap=1, sym=True)
    deco = build_gpu_decomposition(parts, w)
    x = [p["f"].copy() for p in parts]
    assert relerr(deco.apply(x, None), w.apply(x, None)) < TOL
    assert relerr(deco.GMV(x), w.GMV(x)) < 1e-13
    deco.close()


def test_single_subdomain_direct_solve_residual():
    """examples/schwarz.cpp:149-178 (1 rank): numfact + solve, ||Ax-b||/||b|| <= 1e-6."""
    parts, w = make_world(3, 1, mu=2, N=(24, 24, 24), overlap=1)
    deco = build_gpu_decomposition(parts, w)
    b = parts[0]["f"]
    x = deco.subs[0].solve(b)
    A = w.A[0]
    r = np.linalg.norm(A @ x - b, axis=0) / np.linalg.norm(b, axis=0)
    assert r.max() < 1e-12
    deco.close()


def test_config4_elasticity_block_rhs_penalised_dirichlet():
    """BASELINE config 4 in small: Q1 linear elasticity, 3 dofs/node, 2x2x1 subdomains, GenEO nu=6,
    block of 4 right-hand sides (one pass over the panels), clamped face by 1e30 penalisation ->
    Subdomain::boundaryConditions path of Schwarz::start."""
    from hpddm_b200.examples.generate import generate_elasticity3d
    parts = [generate_elasticity3d(r, 4, Nn=(9, 9, 6), overlap=1, mu=4, grid=(2, 2, 1), neumann=True) for r in range(4)]
    w = SchwarzWorld(parts)
    w.multiplicity_scaling()
    w.numfact()
    w.solve_gevp([p["MatNeumann"] for p in parts], nu=6)
    w.build_coarse()
    deco = build_gpu_decomposition(parts, w, two_level=True)
    assert deco.subs[0].statistics()["symmetric"] == 1
    x = rhs(parts, w, 11)
    for corr in (None, DEFLATED, BALANCED):
        assert relerr(deco.apply(x, corr), w.apply(x, corr)) < TOL
    b = [p["f"].copy() for p in parts]
    x0 = rhs(parts, w, 12)
    assert relerr(deco.start(b, x0), w.start(b, [v.copy() for v in x0])) < 1e-13
    deco.end()
    b = w.exchange([p["f"].copy() for p in parts])
    it_ref, x_ref, _ = gmres(OracleOperator(w, DEFLATED), b)
    it_gpu, x_gpu, _ = gmres(KrylovOperator(deco, DEFLATED), b)
    assert it_gpu == it_ref
    assert relerr(x_gpu, x_ref) < 1e-7
    deco.close()


@pytest.mark.parametrize("mu", [1, 2, 4, 7])
def test_block_solve_matches_column_by_column(poisson3d, mu):
    parts, w, deco = poisson3d
    rs = np.random.RandomState(20 + mu)
    for r, s in enumerate(deco.subs[:2]):
        b = np.asfortranarray(rs.standard_normal((w.n[r], mu)))
        got = s.solve(b)
        ref = w.solver[r].solve(b)
        assert relerr([got], [ref]) < TOL


@pytest.mark.parametrize("correction", [None, DEFLATED])
def test_device_resident_gmres_matches_reference_algorithm(poisson3d, correction):
    """hpddm_b200_solve (Krylov basis in HBM) = same iteration count / solution as the restated
    IterativeMethod::GMRES driving the oracle and as the host-driven loop over the C ABI."""
    parts, w, deco = poisson3d
    b = w.exchange([p["f"].copy() for p in parts])
    it_ref, x_ref, _ = gmres(OracleOperator(w, correction), b)
    it_dev, x_dev, res = deco.solve(b, correction=correction)
    assert it_dev == it_ref
    assert relerr(x_dev, x_ref) < 1e-8
    assert np.all(res <= 1e-6)
    r = w.compute_residual(x_dev, b)
    assert np.all(r[:, 1] / r[:, 0] < 1e-5)
    # restart path
    it_ref2, x_ref2, _ = gmres(OracleOperator(w, correction), b, restart=3, max_it=60, tol=1e-8)
    it_dev2, x_dev2, _ = deco.solve(b, correction=correction, restart=3, max_it=60, tol=1e-8)
    assert it_dev2 == it_ref2
    assert relerr(x_dev2, x_ref2) < 1e-8


@pytest.mark.parametrize("method", ["asm", "oras", "soras"])
def test_one_level_variants_asm_oras_soras(method):
    """The other branches of Schwarz::apply (schwarz.hpp:538-546): ASM (SY), ORAS (OG, factorises a
    user-supplied Robin-like matrix), SORAS (OS: D-scaled in and out, unscaled exchange)."""
    import scipy.sparse as sp
    from oracle.schwarz import OG, OS, SY
    parts, w = make_world(3, 4, mu=2, N=(12, 12, 6), overlap=2)
    robin = None
    if method in ("oras", "soras"):
        # optimised transmission conditions modelled by an extra diagonal term on the overlap
        robin = [sp.csr_matrix(w.A[r] + sp.diags(0.5 * abs(w.A[r].diagonal()).max() * (w.d[r] < 1.0))) for r in range(w.P)]
    w.type = {"asm": SY, "oras": OG, "soras": OS}[method]
    w.numfact(robin)
    from hpddm_b200 import Decomposition
    deco = Decomposition(0)
    for r, p in enumerate(parts):
        s = deco.add(r)
        s.initialize(p["Mat"], p["o"], p["mapping"])
        s.setGridHint(*p["dims"])
        s.setScaling(w.d[r])
    for r, s in enumerate(deco.subs):
        s.callNumfact(A=None if robin is None else robin[r], method=method)
    x = rhs(parts, w, 31)
    assert relerr(deco.apply(x, None), w.apply(x, None)) < TOL
    deco.close()


def test_nonuniform_number_of_deflation_vectors():
    """-nonuniform of the reference's test matrix (Makefile:312-347): a different nu on every subdomain."""
    parts, w = make_world(3, 4, nu=4, mu=2, N=(12, 12, 6), overlap=1)
    w.set_vectors([z[:, :1 + r] for r, z in enumerate(w.Z)])   # nu = 1, 2, 3, 4
    w.build_coarse()
    deco = build_gpu_decomposition(parts, w, two_level=True)
    assert deco.getCoarse().shape == (10, 10)
    assert np.abs(deco.getCoarse() - w.E).max() / np.abs(w.E).max() < 1e-12
    x = rhs(parts, w, 41)
    for corr in (DEFLATED, ADDITIVE, BALANCED):
        assert relerr(deco.apply(x, corr), w.apply(x, corr)) < TOL
    deco.close()


def test_error_paths_return_negative_codes_not_crashes():
    """HPDDM_CALL only propagates negative returns (include/HPDDM_iterative.hpp:30-34): every misuse
    must come back as an error code + message, never as a crash or a silent fallback."""
    import scipy.sparse as sp
    from hpddm_b200 import Decomposition, capi
    deco = Decomposition(0)
    s = deco.add(0)
    A = sp.diags([[-1.0] * 9, [2.0] * 10, [-1.0] * 9], [-1, 0, 1], format="csr")
    s.initialize(A, [], [])
    s.setScaling(np.ones(10))
    x = [np.ones((10, 1), order="F")]
    with pytest.raises(capi.HpddmB200Error, match="no factorisation"):
        deco.apply(x, None)                                   # apply before callNumfact
    s.callNumfact()
    assert relerr(deco.apply(x, None), [np.linalg.solve(A.toarray(), x[0])]) < 1e-13
    with pytest.raises(capi.HpddmB200Error, match="no coarse operator"):
        deco.deflation(x)                                     # deflation without buildTwo
    assert relerr(deco.apply(x, "deflated"), deco.apply(x, None)) == 0.0   # correction set but no coarse space: one-level branch (schwarz.hpp:531)
    # singular matrix: numfact reports a pivot breakdown
    s2 = deco.add(1)
    Z = sp.csr_matrix(np.zeros((4, 4)) + np.eye(4) * 0.0 + np.diag([1.0, 0.0, 1.0, 1.0]))
    s2.initialize(Z + sp.csr_matrix(([0.0], ([1], [1])), shape=(4, 4)), [], [])
    with pytest.raises(capi.HpddmB200Error, match="pivot"):
        s2.callNumfact()
    # neighbour that does not exist in a single-process context
    deco2 = Decomposition(0)
    t = deco2.add(0)
    t.initialize(A, [3], [np.array([0, 1], dtype=np.int32)])
    t.setScaling(np.ones(10))
    with pytest.raises(capi.HpddmB200Error, match="does not exist"):
        deco2.exchange(x)
    deco.close()
    deco2.close()


def test_geneo_eigensolve_on_gpu_spans_the_reference_subspace():
    """Schwarz::solveGEVP on the GPU: same eigenvalues and deflation subspace as the dense generalised
    eigensolve of the oracle (ARPACK's tolerance in the reference is 1e-6), same GMRES iteration count."""
    import scipy.linalg as sla
    parts, w = make_world(3, 8, nu=4, mu=1, N=(12, 12, 12), overlap=1)
    deco = build_gpu_decomposition(parts, w, two_level=False)
    for r, s in enumerate(deco.subs):
        lam, it = s.solveGEVP(parts[r]["MatNeumann"], nu=4)
        assert 0 < it < 100
        Zg = s.getVectors()
        A = parts[r]["MatNeumann"].toarray()
        B = w.scale_into_overlap(parts[r]["MatNeumann"], r).toarray()
        th = sla.eigh(B, A, eigvals_only=True)[::-1][:4]
        assert np.abs(lam - 1.0 / th).max() / np.abs(1.0 / th).max() < 1e-6
        # principal angles between the two subspaces (nu = 4 = 1 + a 3-fold degenerate cluster of the cubic subdomain: a closed cluster)
        Qg, _ = np.linalg.qr(Zg)
        Qo, _ = np.linalg.qr(w.Z[r])
        sv = np.linalg.svd(Qg.T @ Qo, compute_uv=False)
        assert np.sqrt(max(0.0, 1 - sv.min() ** 2)) < 1e-3
    deco.buildTwo()
    b = w.exchange([p["f"].copy() for p in parts])
    it_ref, _, _ = gmres(OracleOperator(w, DEFLATED), b)
    it_gpu, x, _ = deco.solve(b, correction=DEFLATED)
    assert it_gpu == it_ref
    res = w.compute_residual(x, b)
    assert np.all(res[:, 1] / res[:, 0] < 1e-5)
    deco.close()


@pytest.mark.parametrize("name", ["40X_400", "mini_mtx"])
def test_local_solve_on_the_reference_data_fixtures(name):
    """SUBDOMAIN::numfact / solve on the reference's own data files (examples/data/40X/400.txt, Fortran numbering, SPD with
    kappa ~ 1e4; examples/data/mini.mtx) -- unstructured matrices: algebraic nested dissection, no grid hint.  The
    reference's acceptance bound for these systems is a relative residual <= 1e-7 (examples/driver.cpp:140)."""
    import scipy.sparse.linalg as spla
    from hpddm_b200 import Decomposition, capi
    from tests.golden_util import load_refdata
    ia, ja, a, numbering, b, A = load_refdata()[name]
    n = A.shape[0]
    deco = Decomposition(0)
    s = deco.add(0)
    L = capi.lib()
    capi.check(L.hpddm_b200_sub_set_matrix(s.h, n, int(a.size), capi.ptr(ia), capi.ptr(ja), capi.ptr(a), 0, numbering.encode()))
    s.n = n
    capi.check(L.hpddm_b200_sub_set_neighbors(s.h, 0, None, None, None))
    s.setScaling(np.ones(n))
    s.callNumfact()
    x = s.solve(b)[:, 0]
    ref = spla.splu(A.tocsc()).solve(b)
    assert np.abs(A @ x - b).max() / np.abs(b).max() < 1e-10
    assert np.abs(x - ref).max() / np.abs(ref).max() < 1e-9
    deco.close()


def test_device_resident_cg_matches_reference_algorithm():
    """hpddm_b200_solve_cg (IterativeMethod::CG, HPDDM_CG.hpp:31-168, r/p/z in HBM) on the symmetric one-level method (ASM),
    3 right-hand sides advancing together: same iteration count and solution as the restated CG driving the oracle; with RAS
    the entry point forwards to GMRES like the reference (CG.hpp:41-44)."""
    from oracle.krylov import cg
    from oracle.schwarz import SY
    parts, w = make_world(3, 8, mu=3, N=(14, 14, 14), overlap=1)
    w.type = SY
    deco = build_gpu_decomposition(parts, w, method="asm")
    b = w.exchange([p["f"].copy() for p in parts])
    it_ref, x_ref = cg(OracleOperator(w, None), b, tol=1e-8)
    it_dev, x_dev, res = deco.solve_cg(b, tol=1e-8)
    assert 5 < it_ref < 100 and it_dev == it_ref
    assert relerr(x_dev, x_ref) < 1e-6          # two CG runs stopped at tol = 1e-8 agree to O(tol * kappa), not to round-off
    assert np.all(res <= 1e-8)
    r = w.compute_residual(x_dev, b)
    assert np.all(r[:, 1] / r[:, 0] < 1e-6)
    deco.close()
    # non-symmetric preconditioner (RAS): GMRES fallback
    w2 = SchwarzWorld(parts)
    w2.multiplicity_scaling()
    w2.numfact()
    deco = build_gpu_decomposition(parts, w2)
    it_g, x_g, _ = gmres(OracleOperator(w2, None), b)
    it_c, x_c, _ = deco.solve_cg(b)
    assert it_c == it_g and relerr(x_c, x_g) < 1e-8
    deco.close()


@pytest.mark.parametrize("correction", [None, DEFLATED])
def test_device_resident_bgmres_matches_reference_algorithm(poisson3d, correction):
    """hpddm_b200_solve_bgmres (IterativeMethod::BGMRES, GMRES.hpp:160-313, block basis in HBM) against the restated reference
    algorithm driving the oracle: 3 right-hand sides in one block Krylov space, same iteration count and solution; and fewer
    iterations than the non-block driver on the same right-hand sides."""
    from oracle.krylov import bgmres
    parts, w, deco = poisson3d
    b = w.exchange([p["f"].copy() for p in parts])
    it_ref, x_ref = bgmres(OracleOperator(w, correction), b)
    it_dev, x_dev, res = deco.solve_bgmres(b, correction=correction)
    assert it_dev == it_ref
    assert relerr(x_dev, x_ref) < 1e-7
    r = w.compute_residual(x_dev, b)
    assert np.all(r[:, 1] / r[:, 0] < 1e-5)
    it_gmres, _, _ = deco.solve(b, correction=correction)
    assert it_dev <= it_gmres


def test_geneo_threshold_selects_a_different_nu_per_subdomain():
    """-hpddm_geneo_threshold semantics on top of the GPU eigensolver: keep the eigenpairs below the threshold; the coarse space
    (non-uniform nu) still builds and the two-level apply matches the oracle run on the same vectors."""
    parts, w = make_world(3, 4, nu=4, mu=2, N=(12, 12, 6), overlap=1)
    deco = build_gpu_decomposition(parts, w)
    kept = []
    for r, s in enumerate(deco.subs):
        lam_all, _ = s.solveGEVP(parts[r]["MatNeumann"], nu=4, tol=1e-8)
        # keep 1 vector (the near-kernel mode of the floating subdomain) on even ranks, all 4 on odd ones
        thr = 0.5 * (lam_all[0] + lam_all[1]) if r % 2 == 0 else 2.0 * abs(lam_all[-1])
        lam, _ = s.solveGEVP(parts[r]["MatNeumann"], nu=4, tol=1e-8, threshold=thr)
        kept.append(len(lam))
        assert np.all(lam < thr)
    assert kept == [1, 4, 1, 4]
    deco.buildTwo()
    w.set_vectors([s.getVectors() for s in deco.subs])
    w.build_coarse()
    x = rhs(parts, w, 5)
    assert relerr(deco.apply(x, DEFLATED), w.apply(x, DEFLATED)) < TOL
    deco.close()


def test_pageable_host_vectors_are_pinned_in_place_inside_a_start_end_bracket():
    """Host boundary (include/HPDDM_GMRES.hpp:45-50,116: the Krylov driver passes ordinary new K[] memory): a range passed for the
    fourth time between start() and end() is registered in place, results are unchanged, and end() drops every registration."""
    parts, w = make_world(3, 1, mu=1, N=(40, 40, 40), overlap=1)   # 64 000 dofs = 512 KB per vector (>= the 64 KB registration floor)
    deco = build_gpu_decomposition(parts, w)
    rs = np.random.RandomState(3)
    x = [np.asfortranarray(rs.standard_normal(p["f"].shape)) for p in parts]
    ref = deco.apply(x)                                            # outside a bracket: plain pageable copies
    assert deco.api.ctx_hostreg_count(deco.ctx) == 0
    y = [np.empty_like(v, order="F") for v in x]
    deco.start([p["f"] for p in parts], [np.zeros_like(p["f"]) for p in parts])
    for k in range(6):
        y[0][:] = 0.0
        deco.apply_host_inplace(x, y, 1, None)
        assert relerr(y, ref) < 1e-13, k                           # (FP64 atomics in the sweeps: equal to round-off, not bit-wise)
        # an unaligned view straddling the registered range: the copy is split at the registration boundaries
        z = np.empty_like(y[0])
        deco.apply_host_inplace(x, [z], 1, None)
        assert relerr([z], ref) < 1e-13
    assert deco.api.ctx_hostreg_count(deco.ctx) >= 2               # x and y (z is a fresh range every time: never registered)
    deco.end()
    n0 = deco.api.ctx_hostreg_count(deco.ctx)
    deco.apply_host_inplace(x, y, 1, None)                         # after end(): pageable again, still correct
    assert relerr(y, ref) < 1e-13 and deco.api.ctx_hostreg_count(deco.ctx) == n0
    deco.close()


@pytest.mark.parametrize("xnode", [5, 18])
def test_cholesky_breakdown_in_any_large_front_of_a_level_falls_back_to_lu(xnode):
    """A symmetric INDEFINITE matrix whose negative pivot sits in one of the two second-level separators (288 unknowns each: both go
    through cuSOLVER, same level).  Every large front has its own status slot, so the failed potrf is seen whichever front of the
    level it hits (a shared slot is overwritten with 0 by the next successful front) and numfact retries with LU."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    from hpddm_b200 import Decomposition
    m = 24
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m))
    I = sp.identity(m)
    A = (sp.kron(sp.kron(I, I), T) + sp.kron(sp.kron(I, T), I) + sp.kron(sp.kron(T, I), I)).tolil()
    i = (7 * m + 12) * m + xnode                # node (x, y, z) = (xnode, 12, 7): inside the y = 12 separator of one x-half
    A[i, i] = -5.0
    A = sp.csr_matrix(A)
    deco = Decomposition(0)
    s = deco.add(0)
    s.initialize(A, [], [])
    s.setGridHint(m, m, m)
    s.callNumfact()
    st = s.statistics()
    assert st["symmetric"] == 0                 # Cholesky was refused, LU took over
    b = np.random.RandomState(1).standard_normal((m ** 3, 1))
    x = s.solve(b)
    ref = spl.spsolve(sp.csc_matrix(A), b[:, 0])
    assert np.abs(x[:, 0] - ref).max() < 1e-9 * np.abs(ref).max()
    deco.close()


def test_large_coarse_space_paths_device_inverse_and_grid_wide_solve(poisson3d, monkeypatch):
    """The code paths taken by coarse spaces too large for one CTA / a host inversion (E^-1 by cuSOLVER LU on the device, the
    replicated solve as three grid-wide passes), forced on the small fixture: same deflation as the oracle."""
    parts, w, deco = poisson3d
    monkeypatch.setenv("HPDDM_B200_COARSE_HOST_LIMIT", "0")
    monkeypatch.setenv("HPDDM_B200_COARSE_ONE_CTA_LIMIT", "0")
    deco.setCoarse(w.E)
    x = rhs(parts, w, 11)
    assert relerr(deco.deflation(x), w.deflation(x)) < TOL
    assert relerr(deco.apply(x, "balanced"), w.apply(x, BALANCED)) < TOL
    monkeypatch.delenv("HPDDM_B200_COARSE_HOST_LIMIT")
    monkeypatch.delenv("HPDDM_B200_COARSE_ONE_CTA_LIMIT")
    deco.setCoarse(w.E)      # back to the default paths for the remaining tests of the module
    assert relerr(deco.deflation(x), w.deflation(x)) < TOL


def test_repeated_applies_replay_the_whole_apply_graph(poisson3d):
    """From its third call on, an apply of a given (mu, correction) is ONE graph launch (single-process context: deflation, SpMV,
    copies between co-hosted subdomains, nested sweep graphs).  Every replay must reproduce the oracle, different inputs included,
    and a setter (here: a new coarse operator) must invalidate the captured graphs."""
    parts, w, deco = poisson3d
    for corr, oc in ((None, None), ("deflated", DEFLATED), ("additive", ADDITIVE), ("balanced", BALANCED)):
        for k in range(5):
            x = rhs(parts, w, 100 + k)
            assert relerr(deco.apply(x, corr), w.apply(x, oc)) < TOL, (corr, k)
    l0 = deco.launches
    x = rhs(parts, w, 200)
    deco.apply(x, "deflated")
    assert deco.launches - l0 > 20            # the replay still accounts for the kernels it launches
    deco.setCoarse(2.0 * w.E)                 # E -> 2 E: the deflated correction halves; a stale graph would keep the old E^-1
    Q = w.deflation(x)
    got = deco.deflation(x)
    assert relerr(got, [0.5 * q for q in Q]) < TOL
    for k in range(4):
        assert relerr(deco.deflation(x), [0.5 * q for q in Q]) < TOL
    deco.setCoarse(w.E)
    for k in range(4):
        assert relerr(deco.apply(x, "deflated"), w.apply(x, DEFLATED)) < TOL


def test_python_mirror_refuses_vector_lists_of_the_wrong_shape(poisson3d):
    """the C ABI takes bare pointer arrays: a list with too few blocks, or a block with the wrong number of rows, would be read out of
    bounds -- the Python mirror refuses them before the call"""
    from hpddm_b200.capi import HpddmB200Error
    parts, w, deco = poisson3d
    x = [p["f"].copy() for p in parts]
    with pytest.raises(HpddmB200Error):
        deco.apply(x[:-1], None)
    with pytest.raises(HpddmB200Error):
        deco.GMV([v[:-1] for v in x])
    with pytest.raises(HpddmB200Error):
        deco.dot(x, [x[0]] + [np.hstack([v, v]) for v in x[1:]])      # blocks of one vector with different numbers of columns
    assert relerr(deco.apply(x, None), w.apply([v.copy() for v in x], None)) < TOL
