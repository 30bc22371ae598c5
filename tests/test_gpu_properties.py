"""GPU: size-independent properties at sizes the CPU oracle cannot reach in seconds
(round trip, linearity, projector identities, block = columns), through the C ABI."""
import numpy as np
import pytest

from bench import cosine_modes
from hpddm_b200 import Decomposition
from hpddm_b200.examples.generate import generate3d, generate_world

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_subdomain():
    m = 72
    part = generate3d(0, 1, N=(m, m, m), overlap=1, mu=4, grid=(1, 1, 1))
    deco = Decomposition(0)
    s = deco.add(0)
    s.initialize(part["Mat"], part["o"], part["mapping"])
    s.setGridHint(*part["dims"])
    deco.multiplicityScaling([part["d"]])
    s.callNumfact()
    s.setVectors(cosine_modes(part["dims"], 20))
    deco.buildTwo()
    yield part, deco, s
    deco.close()


def test_solve_round_trip_at_scale(big_subdomain):
    part, deco, s = big_subdomain
    b = part["f"]
    x = s.solve(b)                                     # 4 right-hand sides: one pass over the panels
    r = np.linalg.norm(part["Mat"] @ x - b, axis=0) / np.linalg.norm(b, axis=0)
    assert r.max() < 1e-11
    st = s.statistics()
    assert st["symmetric"] == 1 and st["nnz_factor"] > 2e8


def test_apply_is_linear_and_blocks_equal_columns(big_subdomain):
    part, deco, s = big_subdomain
    rs = np.random.RandomState(1)
    x, y = (np.asfortranarray(rs.standard_normal((part["ndof"], 1))) for _ in range(2))
    for corr in (None, "deflated", "additive", "balanced"):
        mx, my = deco.apply([x], corr)[0], deco.apply([y], corr)[0]
        mz = deco.apply([2.5 * x - 0.75 * y], corr)[0]
        assert np.abs(mz - (2.5 * mx - 0.75 * my)).max() / np.abs(mz).max() < 1e-11
        blk = deco.apply([np.asfortranarray(np.hstack([x, y, x + y]))], corr)[0]
        assert np.abs(blk[:, :1] - mx).max() / np.abs(mx).max() < 1e-11
        assert np.abs(blk[:, 2:] - (mx + my)).max() / np.abs(mx).max() < 1e-11


def test_two_level_apply_of_A_times_coarse_vector_returns_it(big_subdomain):
    """With the deflated correction M^-1 A z = z for every z in the coarse space (Q A z = z and the
    fine part sees (I - A Q) A z = 0): a projector identity that involves every kernel of the apply."""
    part, deco, s = big_subdomain
    Z = cosine_modes(part["dims"], 20)
    z = np.asfortranarray(Z @ np.random.RandomState(2).standard_normal((20, 1)))
    Az = deco.GMV([z])[0]
    back = deco.apply([Az], "deflated")[0]
    assert np.abs(back - z).max() / np.abs(z).max() < 1e-9
    q = deco.deflation([Az])[0]
    assert np.abs(q - z).max() / np.abs(z).max() < 1e-9


def test_eight_subdomains_on_one_gpu_projector_and_consistency():
    """2x2x2 subdomains of 20^3 cells hosted by one GPU, GenEO computed on the GPU (nu = 4): partition of
    unity, consistency of the halo sum, Q A Q = Q, and a converged device-resident GMRES."""
    parts = generate_world(8, dim=3, N=(40, 40, 40), overlap=1, mu=1, neumann=True)
    deco = Decomposition(0)
    for r, p in enumerate(parts):
        s = deco.add(r)
        s.initialize(p["Mat"], p["o"], p["mapping"])
        s.setGridHint(*p["dims"])
    deco.multiplicityScaling([p["d"] for p in parts])
    ones = deco.exchange([np.ones((p["ndof"], 1)) for p in parts], scaled=True)
    assert max(np.abs(o - 1).max() for o in ones) < 1e-14            # sum_j R_j^T D_j R_j = I
    for s, p in zip(deco.subs, parts):
        s.callNumfact()
        lam, it = s.solveGEVP(p["MatNeumann"], nu=4)
        assert np.all(lam > 0) and np.all(np.diff(lam) >= -1e-9 * lam.max())
    deco.buildTwo()
    E = deco.getCoarse()
    assert np.abs(E - E.T).max() < 1e-9 * np.abs(E).max()
    rs = np.random.RandomState(3)
    v = deco.exchange([rs.standard_normal((p["ndof"], 1)) for p in parts], scaled=True)   # consistent vector
    q = deco.deflation(v)
    q2 = deco.deflation(deco.GMV(q))
    assert max(np.abs(a - b).max() for a, b in zip(q, q2)) / max(np.abs(a).max() for a in q) < 1e-9   # Q A Q = Q
    b = deco.exchange([p["f"].copy() for p in parts], scaled=True)
    it1, x1, _ = deco.solve(b, correction=None)
    it2, x2, res = deco.solve(b, correction="deflated")
    assert 0 < it2 <= it1 and np.all(res <= 1e-6)
    r = deco.GMV(x2)
    num = np.sqrt(deco.dot([a - c for a, c in zip(r, b)], [a - c for a, c in zip(r, b)]))
    den = np.sqrt(deco.dot(b, b))
    assert (num / den).max() < 1e-5
    deco.close()
