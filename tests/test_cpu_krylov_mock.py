"""CPU: hpddm_b200/csrc/hb_krylov.cu ITSELF -- the device Krylov drivers (GMRES, CG, BGMRES), the GCRO-DR DeviceBackend, krylov_entry
and the exported hpddm_b200[z]_solve* entry points -- compiled with g++ against tests/native/krylov_mock.cpp, a host stand-in for
the CUDA runtime and for the kernel launchers the file calls (each launcher a plain loop doing what hb_internal.h says the kernel
does).  What this pins without a GPU: the host side of the device drivers -- pointer arithmetic over block bases, product /
coefficient layouts, staging, the life cycle of the recycled pair in the context -- on several row blocks of uneven size.  What it
cannot pin: the kernels themselves (they have their own GPU tests and are shared with the GPU-verified GMRES / BGMRES drivers)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.gcrodr import bgcrodr as oracle_bgcrodr
from oracle.gcrodr import gcrodr as oracle_gcrodr
from oracle.krylov import bgmres, cg, gmres
from tests.test_cpu_gcrodr import _CsrOperator, _sequence_40x

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST, NONE = 0, -1   # HPDDM_B200_HOST, HPDDM_B200_CORRECTION_NONE


@pytest.fixture(scope="module")
def mock(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("krylov_mock")
    csrc = os.path.join(ROOT, "hpddm_b200", "csrc")
    out = {}
    for name, flags, prefix in (("real", [], "hpddm_b200_"), ("complex", ["-DHB_COMPLEX"], "hpddm_b200z_")):
        so = str(tmp / f"libkrylov_mock_{name}.so")
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-Wl,-Bsymbolic", "-I/usr/local/cuda/include", "-I", os.path.join(ROOT, "include")] + os.environ.get("HB_TEST_CXXFLAGS", "").split() + flags +
                              ["-o", so, "-x", "c++", os.path.join(csrc, "hb_krylov.cu"), os.path.join(csrc, "hb_gcrodr.cpp"), os.path.join(ROOT, "tests", "native", "krylov_mock.cpp")])
        lib = C.CDLL(so)
        lib.krylov_mock_create.restype = C.c_void_p
        lib.krylov_mock_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.krylov_mock_set_values.argtypes = [C.c_void_p]
        lib.krylov_mock_destroy.argtypes = [C.c_void_p]
        lib.krylov_mock_launches.argtypes = [C.c_void_p]
        lib.krylov_mock_set_prcndtnr.argtypes = [C.c_void_p, C.c_int]
        lib.krylov_mock_launches.restype = C.c_long
        lib.krylov_mock_error.restype = C.c_char_p
        out[name] = (lib, prefix)
    return out


class MockDeco:
    def __init__(self, mock, A, sizes, jacobi=False):
        self.cplx = np.iscomplexobj(A.data)
        self.lib, self.prefix = mock["complex" if self.cplx else "real"]
        self.dtype = np.complex128 if self.cplx else np.float64
        A = sp.csr_matrix(A)
        A.sort_indices()
        self.n, self.sizes = A.shape[0], list(sizes)
        assert sum(sizes) == self.n
        self.off = np.concatenate([[0], np.cumsum(sizes)])
        ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(self.dtype)
        sz = np.array(sizes, dtype=np.int32)
        self.ctx = C.c_void_p(self.lib.krylov_mock_create(self.n, ia.ctypes.data, ja.ctypes.data, a.ctypes.data, len(sizes), sz.ctypes.data, int(jacobi)))

    def set_values(self, A):
        A = sp.csr_matrix(A)
        A.sort_indices()
        a = A.data.astype(self.dtype)
        self.lib.krylov_mock_set_values(a.ctypes.data)

    def _call(self, fn, b, *args):
        """b: n x mu; the entry points take one pointer per block, column-major n_q x mu"""
        mu = b.shape[1]
        bs = [np.asfortranarray(b[self.off[q]:self.off[q + 1]].astype(self.dtype)) for q in range(len(self.sizes))]
        xs = [np.zeros_like(v, order="F") for v in bs]
        arr = lambda vs: (C.c_void_p * len(vs))(*[v.ctypes.data for v in vs])
        it = C.c_int(-1)
        res = np.zeros(mu)
        f = getattr(self.lib, self.prefix + fn)
        f.restype = C.c_int
        rc = f(self.ctx, arr(bs), arr(xs), C.c_int(mu), C.c_int(NONE), *args, C.c_int(HOST), C.byref(it), res.ctypes.data_as(C.c_void_p))
        assert rc == 0, (rc, self.lib.krylov_mock_error())
        return it.value, np.vstack(xs), res

    def solve(self, b, restart=40, max_it=100, tol=1e-6):
        return self._call("solve", b, C.c_int(restart), C.c_int(max_it), C.c_double(tol))

    def solve_bgmres(self, b, restart=40, max_it=100, tol=1e-6):
        return self._call("solve_bgmres", b, C.c_int(restart), C.c_int(max_it), C.c_double(tol))

    def solve_cg(self, b, max_it=100, tol=1e-6):
        return self._call("solve_cg", b, C.c_int(max_it), C.c_double(tol))

    def solve_gcrodr(self, b, restart=40, recycle=10, max_it=100, tol=1e-6, target=0, strategy=0, same_system=0):
        return self._call("solve_gcrodr", b, C.c_int(restart), C.c_int(recycle), C.c_int(target), C.c_int(strategy), C.c_int(same_system), C.c_int(max_it), C.c_double(tol))

    def solve_bgcrodr(self, b, restart=40, recycle=10, max_it=100, tol=1e-6, target=0, strategy=0, same_system=0):
        return self._call("solve_bgcrodr", b, C.c_int(restart), C.c_int(recycle), C.c_int(target), C.c_int(strategy), C.c_int(same_system), C.c_int(max_it), C.c_double(tol))

    def recycle_dim(self):
        return getattr(self.lib, self.prefix + "recycle_dim")(self.ctx)

    def recycle_destroy(self):
        return getattr(self.lib, self.prefix + "recycle_destroy")(self.ctx)

    def close(self):
        self.lib.krylov_mock_destroy(self.ctx)


@pytest.mark.parametrize("sizes", [None, (1000, 1, 2987)])
def test_exported_gcrodr_entry_point_passes_the_reference_known_answer_test(mock, sizes):
    """hpddm_b200_solve_gcrodr (entry point -> krylov_entry -> gcrodr_device -> DeviceBackend -> launchers) on the reference's 40X
    sequence: the unmodified examples/driver.cpp needs 497 231 206 198 198 199 206 208 207 206 iterations (Makefile:380); so must this,
    with the vectors in one block or split over three blocks of 1000 / 1 / 2987 rows, the pair carried from system to system in the context"""
    z, mats, rhs = _sequence_40x()
    n = mats[0].shape[0]
    d = MockDeco(mock, mats[0], sizes or (n,))
    got = []
    for i in range(10):
        d.set_values(mats[i])
        it, x, res = d.solve_gcrodr(rhs[i], restart=40, recycle=20, max_it=1000, tol=1e-10)
        got.append(it)
        assert np.linalg.norm(mats[i] @ x[:, 0] - rhs[i][:, 0]) <= 1e-7 * np.linalg.norm(rhs[i])          # examples/driver.cpp:140
        assert res[0] <= 1e-10 and d.recycle_dim() == 20
    assert got == z["gcrodr_40_20_tol1e10"].tolist()
    assert d.lib.krylov_mock_launches(d.ctx) > 0
    d.recycle_destroy()
    assert d.recycle_dim() == 0
    d.close()


def test_exported_gcrodr_entry_point_three_columns_and_same_system(mock):
    """mu = 3 (the reference's -mu 3 line: three copies of the right-hand side) on two blocks; then recycle_same_system 1 -> 2 against
    the oracle restatement on the same operator"""
    z, mats, rhs = _sequence_40x()
    n = mats[0].shape[0]
    d = MockDeco(mock, mats[0], (n - 1500, 1500))
    for i in range(2):
        d.set_values(mats[i])
        it, x, _ = d.solve_gcrodr(np.repeat(rhs[i], 3, axis=1), restart=40, recycle=20, max_it=1000, tol=1e-10)
        assert it == int(z["gcrodr_40_20_tol1e10"][i])
        assert np.abs(x[:, 1] - x[:, 0]).max() == 0 and np.abs(x[:, 2] - x[:, 0]).max() == 0
    d.close()
    d = MockDeco(mock, mats[0], (n,), jacobi=True)
    op = _CsrOperator(mats[0], True)
    state = None
    for s in (1, 2, 3):
        b = np.asfortranarray(rhs[0] * s + np.sin(np.arange(n) * 0.01 * s)[:, None] * 1e3)
        it0, x0, state = oracle_gcrodr(op, [b], restart=30, recycle=8, max_it=600, tol=1e-8, state=state, same_system=min(s, 2))
        it, x, _ = d.solve_gcrodr(b, restart=30, recycle=8, max_it=600, tol=1e-8, same_system=min(s, 2))
        assert abs(it - it0) <= 1, (s, it, it0)
        assert np.abs(x - x0[0]).max() <= 1e-6 * np.abs(x0[0]).max()
    d.close()


def _poisson2d(m, shift=0.0):
    T = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m))
    A = sp.kronsum(T, T).tocsr()
    return (A + shift * sp.identity(m * m)).tocsr()


def test_exported_gcrodr_entry_point_complex(mock):
    """K = std::complex<double> (hpddm_b200z_solve_gcrodr): shifted complex non-Hermitian operator, two right-hand sides, two solves"""
    A = (_poisson2d(24, -0.6 + 0.35j)).astype(np.complex128)
    n = A.shape[0]
    rs = np.random.RandomState(5)
    d = MockDeco(mock, A, (200, n - 200), jacobi=True)
    op = _CsrOperator(A, True)
    state = None
    for s in range(2):
        b = np.asfortranarray(rs.uniform(size=(n, 2)) + 1j * rs.uniform(size=(n, 2)))
        it0, x0, state = oracle_gcrodr(op, [b], restart=12, recycle=4, max_it=400, tol=1e-8, state=state)
        it, x, _ = d.solve_gcrodr(b, restart=12, recycle=4, max_it=400, tol=1e-8)
        assert it == it0, (s, it, it0)
        assert np.abs(x - x0[0]).max() <= 1e-7 * np.abs(x0[0]).max()
    d.close()


def test_gmres_cg_bgmres_drivers_on_the_mock(mock):
    """the three GPU-verified device drivers through the same stand-in: iteration counts of the oracle restatements (which reproduce the
    reference's on the goldens) -- a CPU regression test of their host logic"""
    A = _poisson2d(20)
    n = A.shape[0]
    rs = np.random.RandomState(3)
    b = np.asfortranarray(rs.uniform(size=(n, 3)))
    op = _CsrOperator(A, True)
    op.gram = lambda X, Y: X[0].conj().T @ Y[0]
    d = MockDeco(mock, A, (150, n - 150), jacobi=True)
    it0, x0, _ = gmres(op, [b], restart=15, max_it=300, tol=1e-8)
    it, x, _ = d.solve(b, restart=15, max_it=300, tol=1e-8)
    assert it == it0 and np.abs(x - x0[0]).max() <= 1e-8 * np.abs(x0[0]).max()
    # CG: only for symmetric preconditioners (Prcndtnr SY = 1: ASM); with the default (GE: RAS) the entry point forwards to GMRES(40)
    it_ge, _, _ = d.solve_cg(b, max_it=300, tol=1e-8)
    assert it_ge == gmres(op, [b], restart=40, max_it=300, tol=1e-8)[0]
    d.lib.krylov_mock_set_prcndtnr(d.ctx, 1)
    it0, x0 = cg(op, [b], max_it=300, tol=1e-8)
    it, x, _ = d.solve_cg(b, max_it=300, tol=1e-8)
    assert it == it0 and np.abs(x - x0[0]).max() <= 1e-8 * np.abs(x0[0]).max()
    it0, x0 = bgmres(op, [b], restart=15, max_it=300, tol=1e-8)
    it, x, _ = d.solve_bgmres(b, restart=15, max_it=300, tol=1e-8)
    assert abs(it - it0) <= 1 and np.abs(x - x0[0]).max() <= 1e-6 * np.abs(x0[0]).max()
    d.close()


def test_exported_bgcrodr_entry_point(mock):
    """hpddm_b200[z]_solve_bgcrodr (IterativeMethod::BGCRODR: one block Krylov space, one recycled pair of mu k columns): the reference's
    40X known-answer counts with one right-hand side on three row blocks; three right-hand sides and two solves against the oracle
    restatement (itself pinned by goldens of the reference); complex scalars; the block and the non-block driver swap their pairs"""
    z, mats, rhs = _sequence_40x()
    n = mats[0].shape[0]
    d = MockDeco(mock, mats[0], (1000, 1, 2987))
    for i in range(4):
        d.set_values(mats[i])
        it, x, _ = d.solve_bgcrodr(rhs[i], restart=40, recycle=20, max_it=1000, tol=1e-10)
        assert it == int(z["bgcrodr_40_20_tol1e10"][i]) and d.recycle_dim() == 20
        assert np.linalg.norm(mats[i] @ x[:, 0] - rhs[i][:, 0]) <= 1e-7 * np.linalg.norm(rhs[i])
    d.close()
    for A, cplx in ((_poisson2d(22), False), (_poisson2d(22, -0.6 + 0.35j).astype(np.complex128), True)):
        n = A.shape[0]
        rs = np.random.RandomState(11)
        dd = MockDeco(mock, A, (n - 170, 170), jacobi=True)
        op = _CsrOperator(A, True)
        state = None
        for s in range(2):
            b = np.asfortranarray(rs.uniform(size=(n, 3)) + (1j * rs.uniform(size=(n, 3)) if cplx else 0.0))
            it0, x0, state = oracle_bgcrodr(op, [b], restart=10, recycle=3, max_it=300, tol=1e-8, state=state)
            it, x, _ = dd.solve_bgcrodr(b, restart=10, recycle=3, max_it=300, tol=1e-8)
            assert it == it0, (cplx, s, it, it0)
            assert np.abs(x - x0[0]).max() <= 1e-7 * np.abs(x0[0]).max()
        assert dd.recycle_dim() == 3
        it, x, _ = dd.solve_gcrodr(b, restart=10, recycle=3, max_it=300, tol=1e-8)      # drops the block pair, builds one pair per column
        assert np.abs(A @ x - b).max() <= 1e-6 * np.abs(b).max() and dd.recycle_dim() == 3
        it, x, _ = dd.solve_bgcrodr(b, restart=10, recycle=0, max_it=300, tol=1e-8)     # recycle = 0: BGMRES (GCRODR.hpp:460-465)
        assert it == dd.solve_bgmres(b, restart=10, max_it=300, tol=1e-8)[0]
        dd.close()
