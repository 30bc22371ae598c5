"""CPU: the driver-side generator and the oracle restatement against outputs of the
UNMODIFIED reference (HPDDM headers + examples/generate.cpp compiled against the in-box MPI
shim with the dense LAPACK plugins: oracle/ref_build/).  This is what pins the oracle."""
import numpy as np
import pytest

from hpddm_b200.examples.generate import generate2d
from oracle.gcrodr import bgcrodr, gcrodr
from oracle.krylov import OracleOperator, bgmres, cg, gmres
from oracle.schwarz import ADDITIVE, BALANCED, DEFLATED, SchwarzWorld
from tests.golden_util import cases, col, complexify, load, penalise

TOL = 1e-10


def rel(a, b):
    return np.abs(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name", cases())
def test_generator_is_bit_identical_to_examples_generate_cpp(name):
    parts, ref, meta = load(name)
    for r in range(meta["P"]):
        mine = generate2d(r, meta["P"], Nx=meta["Nx"], Ny=meta["Ny"], overlap=meta["overlap"], mu=0, sym=meta["sym"])
        if meta["complex"]:
            mine = complexify(mine, r)
        if meta["penalised"]:
            mine = penalise(mine, r, meta["P"], meta["Nx"], meta["Ny"], meta["overlap"])
        g = ref[r]
        assert mine["ndof"] == int(g["header"][0])
        assert np.array_equal(mine["Mat"].indptr, g["ia"]) and np.array_equal(mine["Mat"].indices, g["ja"])
        assert np.array_equal(mine["Mat"].data, g["a"])          # bit-exact values, same entry order
        assert np.array_equal(mine["d"], g["d_ramp"])
        assert mine["o"] == [int(v) for v in g["o"]]
        for a, b in zip(mine["mapping"], parts[r]["mapping"]):
            assert np.array_equal(a, b)
        if meta["mu"] == 1:   # (random right-hand sides of -generate_random_rhs come from std::random_device: inputs, not reproducible)
            assert rel(mine["f"], g["f_local" if meta["complex"] else "f"]) < 1e-15


@pytest.mark.parametrize("name", cases())
def test_oracle_reproduces_the_reference(name):
    parts, ref, meta = load(name)
    P = meta["P"]
    w = SchwarzWorld(parts, method=meta["method"])
    w.multiplicity_scaling()
    for r in range(P):
        assert np.abs(w.d[r] - ref[r]["d"]).max() < 1e-15
    w.numfact()
    v = [col(ref[r]["v"]) for r in range(P)]
    got = w.subdomain_exchange([x.copy() for x in v])
    assert max(rel(got[r], ref[r]["subdomain_exchange_v"]) for r in range(P)) < 1e-14
    got = w.GMV(v)
    assert max(rel(got[r], ref[r]["gmv_v"]) for r in range(P)) < 1e-13
    got = w.apply(v, None)
    assert max(rel(got[r], ref[r]["apply_onelevel_v"]) for r in range(P)) < TOL
    corr = None
    if meta["nu"] > 0:
        w.set_vectors([ref[r]["Z"].reshape(meta["nu"], -1).T for r in range(P)])
        # lapack_tr_quirk: see SchwarzWorld.build_coarse -- the goldens were produced with the dense LAPACK coarse plugin
        w.build_coarse(lapack_tr_quirk=True)
        got = w.deflation(v)
        assert max(rel(got[r], ref[r]["deflation_v"]) for r in range(P)) < TOL
        for c, key in ((DEFLATED, "apply_deflated_v"), (ADDITIVE, "apply_additive_v"), (BALANCED, "apply_balanced_v")):
            got = w.apply(v, c)
            assert max(rel(got[r], ref[r][key]) for r in range(P)) < TOL, key
        corr = DEFLATED
    b = [parts[r]["f"].copy() for r in range(P)]
    if meta["krylov"] in ("gcrodr", "bgcrodr"):
        # IterativeMethod::GCRODR over successive solves that share the recycled pair (U, C): identical iteration counts and
        # solutions for every solve of the sequence (the first one builds the pair, the later ones start from it)
        state = None
        for s in range(1, meta["solves"] + 1):
            tag = "" if s == 1 else str(s)
            bs = b if s == 1 else [ref[r]["f" + tag].copy() for r in range(P)]
            it, x, state = (gcrodr if meta["krylov"] == "gcrodr" else bgcrodr)(OracleOperator(w, corr), bs, restart=meta["restart"], max_it=meta["max_it"], tol=meta["tol"], recycle=meta["recycle"], state=state,
                                  target=meta["recycle_target"], same_system=min(s, 2) if meta["same_system"] else 0)
            assert it == int(ref[0]["iterations" + tag][0]), (s, it)
            assert max(rel(x[r], ref[r]["sol" + tag]) for r in range(P)) < 1e-9, s
            res = w.compute_residual(x, bs)
            gold = ref[0]["residual" + tag].reshape(-1, 2)
            assert np.abs(res[:, 0] - gold[:, 0]).max() < 1e-10 * gold[:, 0].max()
            assert np.all(np.abs(res[:, 1] - gold[:, 1]) < 1e-5 * gold[:, 1])
        return
    if meta["krylov"] == "cg":
        it, x = cg(OracleOperator(w, corr), b, max_it=meta["max_it"], tol=meta["tol"])
    elif meta["krylov"] == "bgmres":
        it, x = bgmres(OracleOperator(w, corr), b, restart=meta["restart"], max_it=meta["max_it"], tol=meta["tol"])
    else:
        it, x, _ = gmres(OracleOperator(w, corr), b, restart=meta["restart"], max_it=meta["max_it"], tol=meta["tol"])
    # BGMRES restarted from an ill-conditioned block residual is numerically sensitive by construction: CholQR of a block whose
    # diag(R) spans 5e-7 .. 1e-12 (reference log stored in the golden) loses kappa^2 * eps ~ 1e-5 of orthogonality, so two
    # correct implementations drift apart at the per-cent level after a few restarts.  Exact counts are required everywhere
    # else (GMRES, CG, BGMRES without restart); here +-1 iteration and a residual of the same size.
    sensitive = meta["krylov"] == "bgmres" and int(ref[0]["iterations"][0]) > meta["restart"]
    assert abs(it - int(ref[0]["iterations"][0])) <= (1 if sensitive else 0)     # identical Krylov iteration count
    assert max(rel(x[r], ref[r]["sol"]) for r in range(P)) < (1e-5 if sensitive else 1e-7)
    res = w.compute_residual(x, b)
    gold = ref[0]["residual"].reshape(-1, 2)                       # per right-hand side: ||f||_D, ||A x - f||_D
    assert np.abs(res[:, 0] - gold[:, 0]).max() < 1e-10 * gold[:, 0].max()
    assert np.all(np.abs(res[:, 1] - gold[:, 1]) < (0.5 if sensitive else 1e-3) * gold[:, 1])
