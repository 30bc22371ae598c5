"""Python host mirror of the reference's Schwarz surface over the C ABI.

The reference exposes the same operations to Python through ctypes
(interface/hpddm.py:96-275 over interface/HPDDM.h:96-112: HpddmSchwarzCreate /
Initialize / MultiplicityScaling / CallNumfact / BuildCoarseOperator /
ComputeResidual, HpddmSolve); this module keeps the method names of
HPDDM::Schwarz (include/HPDDM_schwarz.hpp) and forwards every call to
libhpddm_b200.so.  No arithmetic happens here.

One `Decomposition` = the subdomains hosted by this process on one GPU (one MPI
rank = one subdomain in the reference; several per process are allowed so a
whole decomposition can be exercised on one GPU).  Collective calls take / return
one array per local subdomain, column-major (n_loc, mu), float64 or -- Decomposition(dtype=np.complex128),
the hpddm_b200z_* instantiation of the C ABI -- complex128.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import capi


def _f(a, dtype=np.float64):
    a = np.asarray(a, dtype=dtype)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


class Schwarz:
    """One subdomain (HPDDM::Schwarz object, examples/schwarz.cpp:90)."""

    def __init__(self, deco, global_rank):
        self.deco = deco
        self.api = deco.api
        self.dtype = deco.dtype
        self.rank = int(global_rank)
        h = C.c_void_p()
        self.api.check(self.api.sub_create(deco.ctx, self.rank, C.byref(h)))
        self.h = h
        self.n = 0
        self._keep = []

    # Subdomain::initialize(a, o, r)  (include/HPDDM_subdomain.hpp:165-259)
    def initialize(self, Mat, o, mapping, sym=False, numbering="C"):
        A = sp.csr_matrix(Mat)
        ia = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ja = np.ascontiguousarray(A.indices, dtype=np.int32)
        a = np.ascontiguousarray(A.data, dtype=self.dtype)
        self.n = A.shape[0]
        L = self.api
        self.api.check(L.sub_set_matrix(self.h, self.n, int(a.size), capi.ptr(ia), capi.ptr(ja), capi.ptr(a), int(bool(sym)), numbering.encode()))
        ranks = np.ascontiguousarray(list(o), dtype=np.int32)
        sizes = np.ascontiguousarray([len(m) for m in mapping], dtype=np.int32)
        idx = np.ascontiguousarray(np.concatenate([np.asarray(m, dtype=np.int32) for m in mapping]) if len(mapping) else np.zeros(0, np.int32), dtype=np.int32)
        self.api.check(L.sub_set_neighbors(self.h, len(ranks), capi.ptr(ranks), capi.ptr(sizes), capi.ptr(idx)))
        return self

    def setGridHint(self, nx, ny, nz=1, dof=1):
        self.api.check(self.api.sub_set_grid_hint(self.h, int(nx), int(ny), int(nz), int(dof)))

    # Schwarz::initialize(d)  (include/HPDDM_schwarz.hpp:178)
    def setScaling(self, d):
        d = np.ascontiguousarray(d, dtype=np.float64)
        assert d.size == self.n
        self.api.check(self.api.sub_set_scaling(self.h, capi.ptr(d)))

    # Schwarz::callNumfact(A = nullptr)  (include/HPDDM_schwarz.hpp:337-368)
    def callNumfact(self, A=None, method="ras", sym=False):
        L = self.api
        if method == "none":
            t = capi.PRCNDTNR["NO"]
        elif method == "asm":
            t = capi.PRCNDTNR["SY"]
        elif method == "soras":
            t = capi.PRCNDTNR["OS"] if A is not None else capi.PRCNDTNR["SY"]
        elif method in ("oras", "osm") and A is not None:
            t = capi.PRCNDTNR["OG"]
        else:
            t = capi.PRCNDTNR["GE"]
        if A is None:
            self.api.check(L.sub_numfact(self.h, t, 0, 0, None, None, None, 0, b"C"))
        else:
            A = sp.csr_matrix(A)
            ia = np.ascontiguousarray(A.indptr, dtype=np.int32)
            ja = np.ascontiguousarray(A.indices, dtype=np.int32)
            a = np.ascontiguousarray(A.data, dtype=self.dtype)
            self.api.check(L.sub_numfact(self.h, t, A.shape[0], int(a.size), capi.ptr(ia), capi.ptr(ja), capi.ptr(a), int(bool(sym)), b"C"))

    # Preconditioner::setVectors  (include/HPDDM_preconditioner.hpp:358-362)
    def setVectors(self, Z):
        Z = _f(Z, self.dtype)
        assert Z.shape[0] == self.n
        self.api.check(self.api.sub_set_vectors(self.h, capi.ptr(Z), int(Z.shape[1])))

    # Schwarz::solveGEVP<EIGENSOLVER>(MatNeumann)  (include/HPDDM_schwarz.hpp:665-715), on the GPU
    def solveGEVP(self, MatNeumann, nu=20, tol=1e-6, max_it=100, sym=False, threshold=None):
        A = sp.csr_matrix(MatNeumann)
        ia = np.ascontiguousarray(A.indptr, dtype=np.int32)
        ja = np.ascontiguousarray(A.indices, dtype=np.int32)
        a = np.ascontiguousarray(A.data, dtype=self.dtype)
        lam = np.zeros(nu)
        it = self.api.check(self.api.sub_solve_gevp(self.h, A.shape[0], int(a.size), capi.ptr(ia), capi.ptr(ja), capi.ptr(a), int(bool(sym)), b"C",
                                                               int(nu), float(tol), int(max_it), capi.ptr(lam)))
        if threshold is not None and threshold > 0.0:
            # -hpddm_geneo_threshold (include/HPDDM_eigensolver.hpp:69-159 selectNu): of the nu computed pairs keep those whose
            # eigenvalue is below the threshold (at least one) -- the number of deflation vectors then differs per subdomain
            keep = max(1, int(np.count_nonzero(lam < threshold)))
            if keep < nu:
                self.setVectors(self.getVectors()[:, :keep])
                lam = lam[:keep]
        return lam, it

    def getVectors(self):
        nu = C.c_int(0)
        self.api.check(self.api.sub_get_vectors(self.h, None, C.byref(nu)))
        Z = np.zeros((self.n, nu.value), order="F", dtype=self.dtype)
        if nu.value:
            self.api.check(self.api.sub_get_vectors(self.h, capi.ptr(Z), C.byref(nu)))
        return Z

    # SUBDOMAIN::solve(b, x, n)  (e.g. include/HPDDM_SuiteSparse.hpp:388-423)
    def solve(self, b):
        b = _f(b, self.dtype)
        x = np.empty_like(b, order="F")
        self.api.check(self.api.sub_solve(self.h, capi.ptr(b), capi.ptr(x), int(b.shape[1]), capi.HOST))
        return x

    # Subdomain::boundaryConditions  (include/HPDDM_subdomain.hpp:327-336)
    def boundaryConditions(self):
        cnt = C.c_int(0)
        self.api.check(self.api.sub_boundary_conditions(self.h, None, None, C.byref(cnt)))
        idx = np.zeros(cnt.value, dtype=np.int32)
        val = np.zeros(cnt.value, dtype=self.dtype)
        if cnt.value:
            self.api.check(self.api.sub_boundary_conditions(self.h, capi.ptr(idx), capi.ptr(val), C.byref(cnt)))
        return dict(zip(idx.tolist(), val.tolist()))

    def statistics(self):
        st = capi.Stats()
        self.api.check(self.api.sub_stats(self.h, C.byref(st)))
        return {k: getattr(st, k) for k, _ in capi.Stats._fields_}


class Decomposition:
    """The subdomains of this process + the collective hot-path calls."""


    def _vecs(self, vs):
        """one column-major n_i x mu block per local subdomain: the C ABI takes a bare pointer array, so a list of the wrong length or
        a block with the wrong number of rows would be read out of bounds -- refused here"""
        vs = [_f(v, self.dtype) for v in vs]
        if len(vs) != len(self.subs):
            raise capi.HpddmB200Error(f"{len(vs)} vectors for {len(self.subs)} local subdomains")
        for v, s in zip(vs, self.subs):
            if v.shape[0] != s.n or v.shape[1] != vs[0].shape[1]:
                raise capi.HpddmB200Error(f"vector of shape {v.shape} for a subdomain of {s.n} unknowns ({vs[0].shape[1]} columns expected)")
        return vs
    def __init__(self, device=0, dtype=np.float64):
        self.api = capi.api(dtype)
        self.dtype = self.api.dtype
        self.ctx = C.c_void_p()
        self.api.check(self.api.ctx_create(int(device), C.byref(self.ctx)))
        self.subs = []
        self.device = device
        self.correction = None

    def close(self):
        if self.ctx:
            self.api.ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, global_rank):
        s = Schwarz(self, global_rank)
        self.subs.append(s)
        return s

    # --- communicator bootstrap over torch.distributed (any backend)
    def comm_init_torch(self):
        import torch
        import torch.distributed as dist
        rank, size = dist.get_rank(), dist.get_world_size()
        buf = (C.c_char * 128)()
        if rank == 0:
            self.api.check(self.api.nccl_unique_id(buf))
        t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().numpy().tobytes())
        self.api.check(self.api.ctx_comm_init(self.ctx, raw, rank, size))

    # --- host-bootstrapped communicator: the host program's own all-gather carries the control plane (CUDA-IPC handles, coarse
    # sizes), the library's peer-memory fabric carries the hot path; no NCCL needed, several processes may share one GPU
    def comm_init_host(self, rank, size, allgather):
        """allgather(bytes) -> list of `size` bytes objects in rank order (e.g. built on MPI_Allgather / torch.distributed)."""
        def cb(send, recv, nbytes, user):
            try:
                parts = allgather(C.string_at(send, nbytes))
                C.memmove(recv, b"".join(parts), nbytes * size)
                return 0
            except Exception as e:   # never let an exception cross the C boundary
                print(f"[hpddm_b200] host all-gather callback failed: {e!r}")
                return -1
        self._allgather_cb = capi.ALLGATHER_FN(cb)   # keep the trampoline alive as long as the context
        self.api.check(self.api.ctx_comm_init_host(self.ctx, int(rank), int(size), C.cast(self._allgather_cb, C.c_void_p), None))

    def comm_init_host_torch(self):
        """control plane over torch.distributed (any backend, e.g. gloo between processes that share one GPU)"""
        import torch
        import torch.distributed as dist
        rank, size = dist.get_rank(), dist.get_world_size()
        on_gpu = dist.get_backend() == "nccl"

        def allgather(data):
            t = torch.frombuffer(bytearray(data), dtype=torch.uint8)
            if on_gpu:
                t = t.cuda()
            out = [torch.empty_like(t) for _ in range(size)]
            dist.all_gather(out, t)
            return [bytes(o.cpu().numpy().tobytes()) for o in out]
        self.comm_init_host(rank, size, allgather)

    @property
    def transport(self):
        return {0: "single process", 1: "nccl", 2: "peer-memory fabric"}[int(self.api.ctx_transport(self.ctx))]

    def synchronize(self):
        self.api.check(self.api.ctx_synchronize(self.ctx))

    @property
    def stream(self):
        return self.api.ctx_stream(self.ctx)

    @property
    def launches(self):
        return int(self.api.ctx_launch_count(self.ctx))

    # Schwarz::multiplicityScaling  (include/HPDDM_schwarz.hpp:381-404), then initialize(d)
    def multiplicityScaling(self, ds):
        ds = [np.ascontiguousarray(d, dtype=np.float64).copy() for d in ds]
        self.api.check(self.api.multiplicity_scaling(self.ctx, capi.ptr_array(ds)))
        for s, d in zip(self.subs, ds):
            s.setScaling(d)
        return ds

    # Schwarz::buildTwo  (include/HPDDM_schwarz.hpp:440-495)
    def buildTwo(self):
        self.api.check(self.api.build_coarse(self.ctx))

    def setCoarse(self, E):
        E = _f(E, self.dtype)
        self.api.check(self.api.set_coarse(self.ctx, capi.ptr(E), int(E.shape[0])))

    def getCoarse(self):
        n = C.c_int(0)
        self.api.check(self.api.get_coarse(self.ctx, None, C.byref(n)))
        E = np.zeros((n.value, n.value), order="F", dtype=self.dtype)
        self.api.check(self.api.get_coarse(self.ctx, capi.ptr(E), C.byref(n)))
        return E

    # --- hot path, host vectors (what an unchanged Krylov driver passes)
    def _outs(self, ins):
        return [np.empty_like(v, order="F") for v in ins]

    def start(self, b, x):
        b = self._vecs(b)
        x = [_f(v, self.dtype).copy(order="F") for v in x]
        mu = b[0].shape[1]
        self.api.check(self.api.start(self.ctx, capi.ptr_array(b), capi.ptr_array(x), mu, capi.HOST))
        return x

    def end(self):
        self.api.check(self.api.end(self.ctx))

    def apply(self, ins, correction="__default__"):
        corr = self.correction if correction == "__default__" else correction
        ins = self._vecs(ins)
        outs = self._outs(ins)
        self.api.check(self.api.apply(self.ctx, capi.ptr_array(ins), capi.ptr_array(outs), ins[0].shape[1], capi.CORRECTION[corr], capi.HOST))
        return outs

    def deflation(self, ins):
        ins = self._vecs(ins)
        outs = self._outs(ins)
        self.api.check(self.api.deflation(self.ctx, capi.ptr_array(ins), capi.ptr_array(outs), ins[0].shape[1], capi.HOST))
        return outs

    def exchange(self, xs, scaled=True):
        xs = [_f(v, self.dtype).copy(order="F") for v in xs]
        self.api.check(self.api.exchange(self.ctx, capi.ptr_array(xs), xs[0].shape[1], int(bool(scaled)), capi.HOST))
        return xs

    def GMV(self, ins):
        ins = self._vecs(ins)
        outs = self._outs(ins)
        self.api.check(self.api.gmv(self.ctx, capi.ptr_array(ins), capi.ptr_array(outs), ins[0].shape[1], capi.HOST))
        return outs

    def callSolver(self, rhs):
        rhs = [_f(v, self.dtype).copy(order="F") for v in rhs]
        self.api.check(self.api.coarse_solve(self.ctx, capi.ptr_array(rhs), rhs[0].shape[1], capi.HOST))
        return rhs

    def dot(self, x, y):
        x = self._vecs(x)
        y = self._vecs(y)
        mu = x[0].shape[1]
        res = np.zeros(mu, dtype=self.dtype)
        self.api.check(self.api.dot(self.ctx, capi.ptr_array(x), capi.ptr_array(y), mu, capi.ptr(res), capi.HOST))
        return res

    # ||b|| of IterativeMethod::initializeNorm (include/HPDDM_iterative.hpp:455-468): penalised boundary rows count as b_i / HPDDM_PEN
    def rhs_norm(self, b):
        b = self._vecs(b)
        mu = b[0].shape[1]
        out = np.zeros(mu)
        self.api.check(self.api.rhs_norm(self.ctx, capi.ptr_array(b), mu, capi.ptr(out), capi.HOST))
        return out

    # Schwarz::computeResidual (include/HPDDM_schwarz.hpp:761-803): (mu, 2) array of ||f||, ||A x - f||; norm in {"l2", "l1", "linfty"}
    def computeResidual(self, x, f, norm="l2"):
        x = self._vecs(x)
        f = self._vecs(f)
        mu = x[0].shape[1]
        st = np.zeros(2 * mu)
        self.api.check(self.api.compute_residual(self.ctx, capi.ptr_array(x), capi.ptr_array(f), capi.ptr(st), mu, {"l2": 0, "l1": 1, "linfty": 2}[norm], capi.HOST))
        return st.reshape(mu, 2)

    # IterativeMethod::solve (include/HPDDM_iterative.hpp:1013-1111) with the Krylov basis resident in HBM
    def solve(self, b, x0=None, correction="__default__", restart=40, max_it=100, tol=1e-6):
        corr = self.correction if correction == "__default__" else correction
        b = self._vecs(b)
        x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [_f(v, self.dtype).copy(order="F") for v in x0]
        mu = b[0].shape[1]
        it = C.c_int(0)
        res = np.zeros(mu)
        self.api.check(self.api.solve(self.ctx, capi.ptr_array(b), capi.ptr_array(x), mu, capi.CORRECTION[corr], int(restart), int(max_it), float(tol),
                                               capi.HOST, C.byref(it), capi.ptr(res)))
        return it.value, x, res

    # IterativeMethod::BGMRES (include/HPDDM_GMRES.hpp:160-313) on the device: one block Krylov space for all right-hand sides
    def solve_bgmres(self, b, x0=None, correction="__default__", restart=40, max_it=100, tol=1e-6):
        corr = self.correction if correction == "__default__" else correction
        b = self._vecs(b)
        x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [_f(v, self.dtype).copy(order="F") for v in x0]
        mu = b[0].shape[1]
        it = C.c_int(0)
        res = np.zeros(mu)
        self.api.check(self.api.solve_bgmres(self.ctx, capi.ptr_array(b), capi.ptr_array(x), mu, capi.CORRECTION[corr], int(restart), int(max_it), float(tol),
                                             capi.HOST, C.byref(it), capi.ptr(res)))
        return it.value, x, res

    # IterativeMethod::GCRODR (include/HPDDM_GCRODR.hpp:35-444) on the device: Krylov basis and recycled pair (U, C) in HBM; the pair
    # stays in the context between calls (the reference keeps it in A.storage()), recycle_destroy() drops it
    RECYCLE_TARGET = {"SM": 0, "LM": 1, "SR": 2, "LR": 3, "SI": 4, "LI": 5}

    def solve_gcrodr(self, b, x0=None, correction="__default__", restart=40, recycle=10, max_it=100, tol=1e-6, target="SM", strategy="A", same_system=0):
        """same_system: the value of -hpddm_recycle_same_system (0: operator may have changed; 1: same operator, pair built / updated; 2: pair used as is)"""
        corr = self.correction if correction == "__default__" else correction
        b = self._vecs(b)
        x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [_f(v, self.dtype).copy(order="F") for v in x0]
        mu = b[0].shape[1]
        it = C.c_int(0)
        res = np.zeros(mu)
        self.api.check(self.api.solve_gcrodr(self.ctx, capi.ptr_array(b), capi.ptr_array(x), mu, capi.CORRECTION[corr], int(restart), int(recycle),
                                             self.RECYCLE_TARGET[target], {"A": 0, "B": 1}[strategy], int(same_system), int(max_it), float(tol), capi.HOST, C.byref(it), capi.ptr(res)))
        return it.value, x, res

    # IterativeMethod::BGCRODR (include/HPDDM_GCRODR.hpp:445-907): block version, one recycled pair of mu k columns for all right-hand sides
    def solve_bgcrodr(self, b, x0=None, correction="__default__", restart=40, recycle=10, max_it=100, tol=1e-6, target="SM", strategy="A", same_system=0):
        corr = self.correction if correction == "__default__" else correction
        b = self._vecs(b)
        x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [_f(v, self.dtype).copy(order="F") for v in x0]
        mu = b[0].shape[1]
        it = C.c_int(0)
        res = np.zeros(mu)
        self.api.check(self.api.solve_bgcrodr(self.ctx, capi.ptr_array(b), capi.ptr_array(x), mu, capi.CORRECTION[corr], int(restart), int(recycle),
                                              self.RECYCLE_TARGET[target], {"A": 0, "B": 1}[strategy], int(same_system), int(max_it), float(tol), capi.HOST, C.byref(it),
                                              capi.ptr(res)))
        return it.value, x, res

    def recycle_dim(self):
        return int(self.api.recycle_dim(self.ctx))

    def recycle_destroy(self):
        self.api.check(self.api.recycle_destroy(self.ctx))

    # IterativeMethod::CG (include/HPDDM_CG.hpp:31-168) on the device (falls back to GMRES like the reference when the
    # preconditioner is not symmetric: RAS / ORAS or a deflated correction)
    def solve_cg(self, b, x0=None, correction="__default__", max_it=100, tol=1e-6):
        corr = self.correction if correction == "__default__" else correction
        b = self._vecs(b)
        x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [_f(v, self.dtype).copy(order="F") for v in x0]
        mu = b[0].shape[1]
        it = C.c_int(0)
        res = np.zeros(mu)
        self.api.check(self.api.solve_cg(self.ctx, capi.ptr_array(b), capi.ptr_array(x), mu, capi.CORRECTION[corr], int(max_it), float(tol), capi.HOST,
                                         C.byref(it), capi.ptr(res)))
        return it.value, x, res

    # --- hot path, device-resident vectors (raw addresses / torch tensors)
    def apply_device(self, ins, outs, mu, correction="__default__"):
        corr = self.correction if correction == "__default__" else correction
        self.api.check(self.api.apply(self.ctx, capi.ptr_array(ins), capi.ptr_array(outs), int(mu), capi.CORRECTION[corr], capi.DEVICE))

    def apply_host_inplace(self, ins, outs, mu, correction="__default__"):
        """apply on caller-owned host buffers (numpy or pinned torch tensors), no allocation."""
        corr = self.correction if correction == "__default__" else correction
        self.api.check(self.api.apply(self.ctx, capi.ptr_array(ins), capi.ptr_array(outs), int(mu), capi.CORRECTION[corr], capi.HOST))


class KrylovOperator:
    """Adapter to the duck-typed Operator concept of the Krylov drivers
    (include/HPDDM_GMRES.hpp:57-62,113-117): start / apply / GMV / dot."""

    def __init__(self, deco, correction=None):
        self.deco = deco
        self.correction = correction

    def start(self, b, x):
        return self.deco.start(b, x)

    def apply(self, v):
        return self.deco.apply(v, self.correction)

    def GMV(self, v):
        return self.deco.GMV(v)

    def dot(self, x, y):
        return self.deco.dot(x, y)

    def rhs_norm(self, b):
        return self.deco.rhs_norm(b)
