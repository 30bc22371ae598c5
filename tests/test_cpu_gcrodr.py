"""CPU: the product's GCRO-DR driver (hpddm_b200/csrc/hb_gcrodr.cpp -- the code libhpddm_b200.so runs on top of its device
kernels) compiled with a host vector backend (tests/native/gcrodr_host.cpp) and the oracle's operator, against goldens of the
UNMODIFIED reference's IterativeMethod::GCRODR (include/HPDDM_GCRODR.hpp:35-444): identical iteration counts and solutions
for sequences of solves that share the recycled pair (U, C), real and complex scalars, one- and two-level preconditioners,
several right-hand sides.  Also pins the small dense eigen-solver the driver uses instead of LAPACK's hseqr / hsein / ggev."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.gcrodr import bgcrodr as oracle_bgcrodr
from oracle.gcrodr import gcrodr as oracle_gcrodr
from oracle.krylov import OracleOperator
from oracle.schwarz import DEFLATED, SchwarzWorld
from tests.golden_util import cases, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OP_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p))
NORM_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_double))
TARGETS = {"SM": 0, "LM": 1, "SR": 2, "LR": 3, "SI": 4, "LI": 5}


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = {}
    tmp = tmp_path_factory.mktemp("gcrodr")
    for name, flags in (("real", []), ("complex", ["-DHB_COMPLEX"])):
        so = str(tmp / f"libgcrodr_host_{name}.so")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-Wall", "-I/usr/local/cuda/include"] + os.environ.get("HB_TEST_CXXFLAGS", "").split() + flags +
                              ["-o", so, os.path.join(ROOT, "tests", "native", "gcrodr_host.cpp"), os.path.join(ROOT, "hpddm_b200", "csrc", "hb_gcrodr.cpp")])
        lib = C.CDLL(so)
        lib.gcrodr_host_run.restype = C.c_int
        lib.gcrodr_host_run.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), OP_CB, OP_CB, OP_CB, NORM_CB, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_long), C.c_int]
        lib.gcrodr_host_free.argtypes = [C.c_void_p]
        lib.gcrodr_host_recycled_dim.argtypes = [C.c_void_p]
        lib.gcrodr_host_eig.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        out[name] = lib
    return out


class HostGcrodr:
    """Drives gcrodr_host_run with an operator of oracle/krylov.py's concept (start / apply / GMV / rhs_norm)."""

    def __init__(self, lib, op, n, d, mu, dtype):
        self.lib, self.op, self.n, self.mu, self.dtype = lib, op, list(n), mu, dtype
        self.P = len(n)
        self.d = [np.ascontiguousarray(v, dtype=np.float64) for v in d]
        self.state = C.c_void_p(None)
        self.error = None

        def views(ptrs, count):
            return [np.ctypeslib.as_array(C.cast(ptrs[q], C.POINTER(C.c_double)), shape=(self.mu * self.n[q % self.P] * (2 if dtype == np.complex128 else 1),))
                    .view(dtype).reshape(self.mu, self.n[q % self.P]).T for q in range(count)]

        def wrap(fn, nin):
            def cb(_user, pin, pout):
                try:
                    ins = views(pin, nin)
                    outs = views(pout, self.P)
                    res = fn(ins)
                    for q in range(self.P):
                        outs[q][...] = res[q]
                    return 0
                except Exception as e:      # never raise through the C frames
                    self.error = e
                    return -1
            return OP_CB(cb)

        self.cb_apply = wrap(lambda v: self.op.apply([np.asfortranarray(a) for a in v]), self.P)
        self.cb_gmv = wrap(lambda v: self.op.GMV([np.asfortranarray(a) for a in v]), self.P)
        self.cb_start = wrap(lambda v: self.op.start([np.asfortranarray(a) for a in v[:self.P]], [np.array(a, order="F", copy=True) for a in v[self.P:]]), 2 * self.P)

        def norm_cb(_user, pb, pout):
            try:
                nb = self.op.rhs_norm([np.asfortranarray(a) for a in views(pb, self.P)])
                for nu in range(self.mu):
                    pout[nu] = nb[nu]
                return 0
            except Exception as e:
                self.error = e
                return -1
        self.cb_norm = NORM_CB(norm_cb)

    def solve(self, b, restart, recycle, max_it=100, tol=1e-6, target="SM", strategy="A", same_system=0, block=False):
        bs = [np.array(v, dtype=self.dtype, order="F", copy=True) for v in b]
        xs = [np.zeros_like(v, order="F") for v in bs]
        arr = lambda vs: (C.c_void_p * self.P)(*[v.ctypes.data for v in vs])
        it = C.c_int(0)
        rel = np.zeros(self.mu)
        counts = (C.c_long * 2)()
        nn = (C.c_int * self.P)(*self.n)
        rc = self.lib.gcrodr_host_run(self.P, nn, arr(self.d), self.cb_apply, self.cb_gmv, self.cb_start, self.cb_norm, None, arr(bs), arr(xs), self.mu, restart, recycle,
                                      max_it, tol, TARGETS[target], 0 if strategy == "A" else 1, same_system, C.byref(it), rel.ctypes.data_as(C.POINTER(C.c_double)), C.byref(self.state), counts, int(block))
        if self.error is not None:
            raise self.error
        assert rc == 0, rc
        return it.value, xs, rel, (counts[0], counts[1])

    def close(self):
        self.lib.gcrodr_host_free(self.state)
        self.state = C.c_void_p(None)


def rel(a, b):
    return np.abs(np.asarray(a).reshape(-1) - np.asarray(b).reshape(-1)).max() / max(np.abs(b).max(), 1e-300)


def _world(name):
    parts, ref, meta = load(name)
    P = meta["P"]
    w = SchwarzWorld(parts, method=meta["method"])
    w.multiplicity_scaling()
    w.numfact()
    corr = None
    if meta["nu"] > 0:
        w.set_vectors([ref[r]["Z"].reshape(meta["nu"], -1).T for r in range(P)])
        w.build_coarse(lapack_tr_quirk=True)
        corr = DEFLATED
    return parts, ref, meta, w, OracleOperator(w, corr)


@pytest.mark.parametrize("name", [n for n in cases() if "_gcrodr_" in n])
def test_product_driver_reproduces_the_reference_gcrodr(harness, name):
    parts, ref, meta, w, op = _world(name)
    P = meta["P"]
    dtype = np.complex128 if meta["complex"] else np.float64
    h = HostGcrodr(harness["complex" if meta["complex"] else "real"], op, [p["ndof"] for p in parts], w.d, meta["mu"], dtype)
    for s in range(1, meta["solves"] + 1):
        tag = "" if s == 1 else str(s)
        b = [parts[r]["f"] if s == 1 else ref[r]["f" + tag] for r in range(P)]
        it, x, res, _ = h.solve(b, meta["restart"], meta["recycle"], max_it=meta["max_it"], tol=meta["tol"], target=meta["recycle_target"],
                                same_system=min(s, 2) if meta["same_system"] else 0)
        assert it == int(ref[0]["iterations" + tag][0]), (s, it)                       # the reference's iteration count
        assert max(rel(x[r], ref[r]["sol" + tag]) for r in range(P)) < 1e-8, s          # and its solution
        assert np.all(res <= meta["tol"])
    assert h.lib.gcrodr_host_recycled_dim(h.state) == min(meta["recycle"], meta["restart"] - 1)
    h.close()


@pytest.mark.parametrize("target,strategy", [("SM", "B"), ("LM", "A"), ("LR", "A"), ("LR", "B")])
def test_product_driver_matches_the_oracle_for_other_targets_and_strategy_b(harness, target, strategy):
    """recycle_target / recycle_strategy other than the golden ones, against the oracle restatement (which reproduces the reference's
    counts for LM / SR / LR with strategy A: 30 + 34, 22 + 15 iterations on this problem).  Strategy B cannot be pinned to the
    iteration: its pencil [I B; B^H ...] z = theta [I 0; B^H ...] z has the eigenvalue 1 with multiplicity k by construction, so the
    choice among equal harmonic Ritz values -- and the basis of that eigenspace -- is rounding noise in the reference itself (observed
    with a debug build: 1.000000 three times for k = 3).  Strategy A: exact counts; strategy B: same behaviour, a few iterations."""
    parts, ref, meta, w, op = _world("small_40x40_p4_gcrodr_m8_k4_solves3")
    P = meta["P"]
    h = HostGcrodr(harness["real"], op, [p["ndof"] for p in parts], w.d, 1, np.float64)
    state = None
    for s in (1, 2):
        b = [parts[r]["f"] if s == 1 else ref[r]["f2"] for r in range(P)]
        it0, x0, state = oracle_gcrodr(op, b, restart=8, recycle=3, tol=1e-7, state=state, target=target, strategy=strategy)
        it, x, _, _ = h.solve(b, 8, 3, tol=1e-7, target=target, strategy=strategy)
        assert abs(it - it0) <= (0 if strategy == "A" else 4), (s, it, it0)
        assert max(rel(x[r], x0[r]) for r in range(P)) < (1e-9 if strategy == "A" else 1e-5)
    h.close()


def test_driver_edge_cases(harness):
    """zero right-hand side (iterations = 0 like the reference), iteration limit inside a cycle, recycled dimension clipped to restart - 1,
    a pair stored for another number of right-hand sides is dropped"""
    parts, ref, meta, w, op = _world("small_40x40_p4_gcrodr_m8_k4_solves3")
    P = meta["P"]
    n = [p["ndof"] for p in parts]
    h = HostGcrodr(harness["real"], op, n, w.d, 1, np.float64)
    it, x, _, _ = h.solve([np.zeros((k, 1)) for k in n], 8, 4)
    assert it == 0 and all(np.all(v == 0) for v in x)
    b = [parts[r]["f"] for r in range(P)]
    for max_it in (5, 8, 11, 13):
        h.close()
        it0, x0, _ = oracle_gcrodr(op, b, restart=8, recycle=4, max_it=max_it)
        it, x, _, _ = h.solve(b, 8, 4, max_it=max_it)
        assert it == it0 == max_it
        assert max(rel(x[r], x0[r]) for r in range(P)) < 1e-9, max_it
    h.close()
    it, x, _, _ = h.solve(b, 6, 50)
    assert harness["real"].gcrodr_host_recycled_dim(h.state) == 5
    h2 = HostGcrodr(harness["real"], op, n, w.d, 2, np.float64)
    h2.state = h.state
    b2 = [np.asfortranarray(np.hstack([v, 2.0 * v[::-1]])) for v in b]
    it2, x2, _, _ = h2.solve(b2, 8, 4)
    res = w.compute_residual(x2, b2)
    assert np.all(res[:, 1] <= 1.5e-6 * res[:, 0])
    h2.close()


@pytest.mark.parametrize("kind", ["real", "complex", "nonnormal", "companion"])
def test_dense_eigen_solver(harness, kind):
    rng = np.random.default_rng(7)
    lib = harness["real"]
    for n in (1, 2, 3, 7, 16, 41):
        if kind == "real":
            A = rng.standard_normal((n, n)).astype(np.complex128)
        elif kind == "complex":
            A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        elif kind == "nonnormal":
            A = (np.triu(rng.standard_normal((n, n)), 0) * 5 + 1e-3 * rng.standard_normal((n, n))).astype(np.complex128)
        else:   # companion matrix of prod (x - j): Hessenberg already, well separated real spectrum
            A = np.zeros((n, n), dtype=np.complex128)
            A[np.arange(1, n), np.arange(n - 1)] = 1.0
            A[:, -1] = -np.poly(np.arange(1, n + 1) / n)[::-1][:n]
        a = np.asfortranarray(A)
        w = np.zeros(n, dtype=np.complex128)
        X = np.zeros((n, n), dtype=np.complex128, order="F")
        assert lib.gcrodr_host_eig(n, a.ctypes.data, w.ctypes.data, X.ctypes.data) == 0
        scale = max(np.abs(A).max(), 1.0)
        assert np.abs(A @ X - X * w[None, :]).max() < (1e-6 if kind == "companion" else 1e-10) * scale * n        # A x = lambda x
        assert np.allclose(np.linalg.norm(X, axis=0), 1.0)
        if kind != "companion":
            want = np.linalg.eigvals(A)
            assert max(np.abs(want - z).min() for z in w) < 1e-8 * scale


# ---- the reference's own known-answer test of its recycling driver: examples/driver.cpp on the in-tree 40X sequence ----------------
class _CsrOperator:
    """examples/driver.cpp's CustomOperator (lines 45-62): no domain decomposition, GMV = csrmm, apply = copy or, with
    -diagonal_scaling, division by the diagonal; unweighted inner products (getScaling() == nullptr); start does nothing."""

    def __init__(self, A, jacobi):
        self.A, self.jacobi, self.dg = A, jacobi, A.diagonal()

    def start(self, b, x):
        return x

    def apply(self, v):
        return [np.asfortranarray(v[0] / self.dg[:, None]) if self.jacobi else np.array(v[0], order="F", copy=True)]

    def GMV(self, v):
        return [np.asfortranarray(self.A @ v[0])]

    def dot(self, x, y):
        return (np.conj(x[0]) * y[0]).sum(axis=0)

    def rhs_norm(self, b):
        return np.sqrt(np.real(self.dot(b, b)))

    def gram(self, X, Y):
        return X[0].conj().T @ Y[0]


def _sequence_40x():
    import scipy.sparse as sp
    from tests.golden_util import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "refdata_40X_sequence.npz"))
    n = int(z["n"])
    mats = [sp.csr_matrix((z["a"][i], z["ja"] - 1, z["ia"] - 1), shape=(n, n)) for i in range(10)]
    return z, mats, [np.asfortranarray(z["rhs"][i].reshape(-1, 1)) for i in range(10)]


# pass windows of the reference's test (examples/driver.cpp:152-155, right preconditioning): total iterations over the ten systems
WINDOW = {False: (2346, 2366), True: (2055, 2075)}


@pytest.mark.parametrize("jacobi", [False, True])
def test_reference_known_answer_40X_sequence(harness, jacobi):
    """GCRO-DR(40, 20), tol 1e-10, max_it 1000 on ten slowly changing systems, the recycled pair carried from one system to the next
    (the operator changes: C = A M^-1 U is recomputed at every new system, GCRODR.hpp:94-130).  Reference (oracle/_ref/driver_ref, the
    unmodified examples/driver.cpp): 497 231 206 198 198 199 206 208 207 206 = 2356 iterations, 2065 with -diagonal_scaling; its own
    acceptance: every residual <= 1e-7 and the total inside a window of +-10.  Required here, for the oracle restatement AND the product's
    driver on the host backend: residuals <= 1e-7, the total inside the reference's window, and per system exactly the reference's
    count without scaling; with -diagonal_scaling two of the ten systems end one or two iterations apart (179 183 vs 178 181-182: no
    near-tie or cut conjugate pair in the harmonic Ritz selection, the residual curve is flat around 1e-10 there), inside the reference's
    own tolerance of +-10 on the total."""
    z, mats, rhs = _sequence_40x()
    want = z["gcrodr_40_20_tol1e10_diagonal_scaling" if jacobi else "gcrodr_40_20_tol1e10"]
    assert WINDOW[jacobi][0] < int(want.sum()) < WINDOW[jacobi][1]          # the stored reference run passed its own test
    n = mats[0].shape[0]
    state = None
    h = None
    got_oracle, got_product = [], []
    for i in range(10):
        op = _CsrOperator(mats[i], jacobi)
        it, x, state = oracle_gcrodr(op, [rhs[i]], tol=1e-10, max_it=1000, restart=40, recycle=20, state=state)
        got_oracle.append(it)
        assert np.linalg.norm(mats[i] @ x[0][:, 0] - rhs[i][:, 0]) <= 1e-7 * np.linalg.norm(rhs[i])      # driver.cpp:140
        if h is None:
            h = HostGcrodr(harness["real"], op, [n], [np.ones(n)], 1, np.float64)
        h.op = op
        it, x, _, _ = h.solve([rhs[i]], 40, 20, max_it=1000, tol=1e-10)
        got_product.append(it)
        assert np.linalg.norm(mats[i] @ x[0][:, 0] - rhs[i][:, 0]) <= 1e-7 * np.linalg.norm(rhs[i])
    h.close()
    for got in (got_oracle, got_product):
        assert WINDOW[jacobi][0] < sum(got) < WINDOW[jacobi][1], got
        assert np.abs(np.array(got) - want).max() <= (3 if jacobi else 0), (got, want.tolist())
    if not jacobi:
        assert got_oracle == want.tolist() and got_product == want.tolist()


def test_reference_40X_sequence_three_identical_columns(harness):
    """the reference's test line also runs -mu 3 (the right-hand side copied three times, examples/driver.cpp:119-120): every column has
    its own Krylov space and recycled pair, so each must take exactly the single-column count -- first three systems of the sequence"""
    z, mats, rhs = _sequence_40x()
    n = mats[0].shape[0]
    h = None
    for i in range(3):
        op = _CsrOperator(mats[i], False)
        if h is None:
            h = HostGcrodr(harness["real"], op, [n], [np.ones(n)], 3, np.float64)
        h.op = op
        b = np.asfortranarray(np.repeat(rhs[i], 3, axis=1))
        it, x, _, _ = h.solve([b], 40, 20, max_it=1000, tol=1e-10)
        assert it == int(z["gcrodr_40_20_tol1e10"][i])
        assert np.abs(x[0][:, 1] - x[0][:, 0]).max() == 0 and np.abs(x[0][:, 2] - x[0][:, 0]).max() == 0
    h.close()


def test_reference_known_answer_40X_sequence_block_driver():
    """the same acceptance test for IterativeMethod::BGCRODR (Makefile:384: -hpddm_krylov_method bgcrodr, one right-hand side): the
    unmodified driver needs the same 2356 iterations, system by system; so does the oracle restatement of the block driver"""
    z, mats, rhs = _sequence_40x()
    want = z["bgcrodr_40_20_tol1e10"]
    assert WINDOW[False][0] < int(want.sum()) < WINDOW[False][1]
    state = None
    got = []
    for i in range(10):
        it, x, state = oracle_bgcrodr(_CsrOperator(mats[i], False), [rhs[i]], tol=1e-10, max_it=1000, restart=40, recycle=20, state=state)
        got.append(it)
        assert np.linalg.norm(mats[i] @ x[0][:, 0] - rhs[i][:, 0]) <= 1e-7 * np.linalg.norm(rhs[i])
    assert got == want.tolist()


@pytest.mark.parametrize("name", [n for n in cases() if "_bgcrodr_" in n])
def test_product_block_driver_reproduces_the_reference_bgcrodr(harness, name):
    """gcro::run_block (IterativeMethod::BGCRODR) on the host backend against goldens of the unmodified reference: one, two and three
    right-hand sides, complex scalars, sequences of solves sharing the block pair"""
    parts, ref, meta, w, op = _world(name)
    P = meta["P"]
    dtype = np.complex128 if meta["complex"] else np.float64
    h = HostGcrodr(harness["complex" if meta["complex"] else "real"], op, [p["ndof"] for p in parts], w.d, meta["mu"], dtype)
    for s in range(1, meta["solves"] + 1):
        tag = "" if s == 1 else str(s)
        b = [parts[r]["f"] if s == 1 else ref[r]["f" + tag] for r in range(P)]
        it, x, res, _ = h.solve(b, meta["restart"], meta["recycle"], max_it=meta["max_it"], tol=meta["tol"], block=True)
        assert it == int(ref[0]["iterations" + tag][0]), (s, it)
        assert max(rel(x[r], ref[r]["sol" + tag]) for r in range(P)) < 1e-8, s
    assert h.lib.gcrodr_host_recycled_dim(h.state) == min(meta["recycle"], meta["restart"] - 1)
    # the two drivers do not share a pair: the non-block driver drops the block pair and builds its own
    it, x, _, _ = h.solve([parts[r]["f"] for r in range(P)], meta["restart"], meta["recycle"], max_it=meta["max_it"], tol=meta["tol"])
    res = w.compute_residual(x, [parts[r]["f"] for r in range(P)])
    assert np.all(res[:, 1] <= 3 * meta["tol"] * res[:, 0])
    h.close()


def test_product_block_driver_passes_the_reference_known_answer_test(harness):
    """first four systems of the 40X sequence through gcro::run_block (one right-hand side): 497 231 206 198, as the unmodified driver"""
    z, mats, rhs = _sequence_40x()
    n = mats[0].shape[0]
    h = None
    for i in range(4):
        op = _CsrOperator(mats[i], False)
        if h is None:
            h = HostGcrodr(harness["real"], op, [n], [np.ones(n)], 1, np.float64)
        h.op = op
        it, x, _, _ = h.solve([rhs[i]], 40, 20, max_it=1000, tol=1e-10, block=True)
        assert it == int(z["bgcrodr_40_20_tol1e10"][i])
        assert np.linalg.norm(mats[i] @ x[0][:, 0] - rhs[i][:, 0]) <= 1e-7 * np.linalg.norm(rhs[i])
    h.close()
