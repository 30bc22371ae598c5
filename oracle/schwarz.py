"""CPU restatement of the HPDDM RAS preconditioner-apply hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

All "ranks" live in one process: a vector is a list (one entry per rank) of
``(n_loc, mu)`` Fortran-ordered float64 (or complex128: K = std::complex<double>,
the reference's FORCE_COMPLEX build / BASELINE config 5) arrays, exactly the column-major
``n_loc x mu`` blocks the reference passes around.  Complex scalars
follow the reference's conventions: conjugate transposes where it uses Wrapper<K>::transc
(Z^H in deflation and in the Galerkin operator, conj on the first argument of inner
products), real partition of unity (underlying_type<K>).

Restated functions (reference file:line):
  Subdomain.initialize        include/HPDDM_subdomain.hpp:165-236
  Subdomain.exchange          include/HPDDM_subdomain.hpp:115-130
  Schwarz.multiplicityScaling include/HPDDM_schwarz.hpp:381-404
  Schwarz.exchange            include/HPDDM_schwarz.hpp:180-188
  Schwarz.start               include/HPDDM_schwarz.hpp:496-514
  Schwarz.apply               include/HPDDM_schwarz.hpp:527-612
  Schwarz.deflation           include/HPDDM_schwarz.hpp:1602-1622
  Schwarz.GMV                 include/HPDDM_schwarz.hpp:726-747
  Schwarz.computeResidual     include/HPDDM_schwarz.hpp:761-803
  Schwarz.scaleIntoOverlap    include/HPDDM_schwarz.hpp:622-657
  Schwarz.solveGEVP           include/HPDDM_schwarz.hpp:665-715 (+ HPDDM_ARPACK.hpp:84-178)
  CoarseOperator.callSolver   include/HPDDM_coarse_operator_impl.hpp:1630-1732
  coarse assembly E           include/HPDDM_operator.hpp:395-403,440-528
  Wrapper.diag/gthr/csrmm     include/HPDDM_wrapper.hpp:310-318,697-733,820-831
  Subdomain.boundaryCond      include/HPDDM_subdomain.hpp:310-336

The local solve ``s_.solve`` (HPDDM_schwarz.hpp:535,542,544,557,567,590) is
third-party in the reference (MUMPS / SuiteSparse, un-vendored, unpinned:
SURVEY.md section 0 item 3); here it is SuperLU through scipy (``splu``) or a
dense LAPACK factorisation for tiny blocks -- any backward-stable direct solve
yields the same operator to O(kappa * eps).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import scipy.linalg as sla

HPDDM_EPS = 1.0e-12  # include/HPDDM_define.hpp:46
HPDDM_PEN = 1.0e30   # include/HPDDM_define.hpp:47

# Prcndtnr (include/HPDDM_enum.hpp), coarse corrections (HPDDM_define.hpp:141-199)
NO, SY, GE, OS, OG = "NO", "SY", "GE", "OS", "OG"
DEFLATED, ADDITIVE, BALANCED = "deflated", "additive", "balanced"


def full_csr(A, sym):
    """MatrixCSR with sym_=true stores the lower triangle (matrix.hpp:33-57);
    expand to a full CSR for the oracle's arithmetic."""
    A = sp.csr_matrix(A)
    if sym:
        A = A + sp.tril(A, -1).T
    return sp.csr_matrix(A)


class LocalSolver:
    """SUBDOMAIN<K> concept: numfact / solve (HPDDM_SuiteSparse.hpp:224-424)."""

    def __init__(self, A, dense_below=0):
        A = sp.csc_matrix(A)
        self.n = A.shape[0]
        self.dtype = np.complex128 if np.iscomplexobj(A.data) else np.float64
        if self.n <= dense_below:
            self.lu = sla.lu_factor(A.toarray())
            self.dense = True
        else:
            self.lu = spla.splu(A, permc_spec="COLAMD", options=dict(SymmetricMode=False))
            self.dense = False

    def solve(self, b):
        b = np.asarray(b, dtype=np.result_type(self.dtype, np.asarray(b).dtype))
        if self.dense:
            return np.asfortranarray(sla.lu_solve(self.lu, b))
        out = np.empty_like(b, order="F")
        if b.ndim == 1:
            return self.lu.solve(b)
        for j in range(b.shape[1]):  # UMFPACK path is column-at-a-time too (SuiteSparse.hpp:401-406)
            out[:, j] = self.lu.solve(np.ascontiguousarray(b[:, j]))
        return out


class SchwarzWorld:
    """All ranks of a HPDDM::Schwarz decomposition in one process."""

    def __init__(self, parts, method="ras"):
        """parts: list (per rank) of dicts as produced by oracle.generate."""
        self.P = len(parts)
        self.A = [full_csr(p["Mat"], p.get("sym", False)) for p in parts]
        self.n = [a.shape[0] for a in self.A]
        # Subdomain::initialize (subdomain.hpp:173-188): neighbours sorted by rank,
        # empty mappings dropped.  NOTE the reference indexes r[idx[j]] while
        # iterating idx -- identical to r[i]; restated faithfully.
        self.map = []
        for p in parts:
            order = sorted(range(len(p["o"])), key=lambda q: p["o"][q])
            m = [(int(p["o"][q]), np.asarray(p["mapping"][q], dtype=np.int64)) for q in order if len(p["mapping"][q]) > 0]
            self.map.append(m)
        self.d = [np.array(p["d"], dtype=np.float64) for p in parts]
        self.type = GE if method == "ras" else (SY if method == "asm" else method)
        self.solver = None
        self.Z = None     # ev_: per rank (n_loc, nu_i) F-order  (preconditioner.hpp:106)
        self.E = None     # assembled coarse operator (dense)
        self.Elu = None
        self.nu = None
        self.bc = None

    # ------------------------------------------------------------------ setup
    def multiplicity_scaling(self):
        """schwarz.hpp:381-404: d <- d / (sum over sharing ranks of their d);
        processed in neighbour order (the reference uses MPI_Waitany order)."""
        send = [[self.d[r][idx].copy() for (_, idx) in self.map[r]] for r in range(self.P)]
        new = [np.ones_like(d) for d in self.d]
        for r in range(self.P):
            for k, (nb, idx) in enumerate(self.map[r]):
                kk = [q for q, (rr, _) in enumerate(self.map[nb]) if rr == r][0]
                recv = send[nb][kk]
                s = send[r][k]
                for j in range(idx.size):
                    if abs(s[j]) < HPDDM_EPS:
                        new[r][idx[j]] = 0.0
                    else:
                        new[r][idx[j]] /= 1.0 + new[r][idx[j]] * recv[j] / s[j]
        self.d = new

    def numfact(self, mats=None, dense_below=0):
        """Schwarz::callNumfact (schwarz.hpp:337-368)."""
        mats = self.A if mats is None else mats
        self.solver = [LocalSolver(a, dense_below) for a in mats]

    # --------------------------------------------------------------- exchange
    def subdomain_exchange(self, x):
        """Subdomain::exchange (subdomain.hpp:115-130): in[map] += neighbour's
        gathered values, per column.  Sends are gathered BEFORE any add
        (gthr precedes the Waitany loop for every neighbour)."""
        send = [[x[r][idx, :].copy() for (_, idx) in self.map[r]] for r in range(self.P)]
        for r in range(self.P):
            for k, (nb, idx) in enumerate(self.map[r]):
                kk = [q for q, (rr, _) in enumerate(self.map[nb]) if rr == r][0]
                np.add.at(x[r], (idx, slice(None)), send[nb][kk])
        return x

    def exchange(self, x):
        """Schwarz::exchange (schwarz.hpp:180-188) = Wrapper::diag then halo sum."""
        for r in range(self.P):
            x[r] *= self.d[r][:, None]
        return self.subdomain_exchange(x)

    # ------------------------------------------------------------------- SpMV
    def csrmm(self, x, r):
        return np.asfortranarray(self.A[r] @ x)

    def GMV(self, x):
        """schwarz.hpp:726-747: out = exchange(A_loc * in)."""
        out = [self.csrmm(x[r], r) for r in range(self.P)]
        return self.exchange(out)

    # ------------------------------------------------------ boundary conditions
    def boundary_conditions(self, r):
        """Subdomain::boundaryCond(itions) (subdomain.hpp:310-336): row i is a boundary
        condition when its diagonal is penalised (|a_ii| >= EPS*PEN) or when the part of the
        row left of / on the diagonal is the identity; value = a_ii.  None for the Poisson
        generators (diagonals O(1e2), non-trivial rows)."""
        A = self.A[r]
        out = {}
        dg = A.diagonal()
        for i in np.nonzero(np.abs(dg) >= HPDDM_EPS * HPDDM_PEN)[0]:
            out[int(i)] = dg[i]
        for i in np.nonzero(np.abs(dg - 1.0) <= HPDDM_EPS)[0]:
            lo, hi = A.indptr[i], A.indptr[i + 1]
            cols, vals = A.indices[lo:hi], A.data[lo:hi]
            left = cols < i
            if not np.any(np.abs(vals[left]) > HPDDM_EPS):
                out[int(i)] = dg[i]
        return out

    def start(self, b, x):
        """schwarz.hpp:496-514 (halo buffers are implicit here)."""
        for r in range(self.P):
            for i, v in self.boundary_conditions(r).items():
                x[r][i, :] = b[r][i, :] / v
        self.exchange(x)
        return x

    # --------------------------------------------------------------- two-level
    def set_vectors(self, Z):
        self.Z = [np.asfortranarray(z, dtype=np.complex128 if np.iscomplexobj(z) else np.float64) for z in Z]
        self.nu = [z.shape[1] for z in self.Z]

    def scale_into_overlap(self, A, r):
        """schwarz.hpp:622-657: B = D A D restricted to overlap dofs with d>EPS."""
        into = set()
        for (_, idx) in self.map[r]:
            for i in idx:
                if self.d[r][i] > HPDDM_EPS:
                    into.add(int(i))
        into = np.array(sorted(into), dtype=np.int64)
        n = self.n[r]
        mask = np.zeros(n, dtype=bool)
        mask[into] = True
        D = sp.diags(self.d[r] * mask)
        B = sp.csr_matrix(D @ sp.csr_matrix(A) @ D)
        B.data[np.abs(B.data) <= HPDDM_EPS] = 0.0
        B.eliminate_zeros()
        return B

    def solve_gevp(self, neumann, nu=20, dense_limit=6000):
        """Schwarz::solveGEVP (schwarz.hpp:665-715) with EIGENSOLVER=ARPACK
        (HPDDM_ARPACK.hpp:84-178): A_Neu x = lambda B x, nu smallest lambda via
        shift-invert at sigma = 0 (largest 1/lambda of A^{-1} B).  Entries below
        1/(EPS*PEN)=1e-18 are zeroed (schwarz.hpp:713)."""
        Z = []
        for r in range(self.P):
            A = sp.csr_matrix(neumann[r])
            B = self.scale_into_overlap(A, r)
            n = self.n[r]
            k = min(nu, n - 1)
            symmetric = abs(A - A.T).max() <= 1e-12 * abs(A).max()
            V = None
            if n <= dense_limit and symmetric:
                try:
                    wv, V = sla.eigh(B.toarray(), A.toarray(), subset_by_index=[n - k, n - 1])
                    V = V[:, ::-1]
                except np.linalg.LinAlgError:
                    V = None
            if V is None:
                lu = spla.splu(sp.csc_matrix(A + 1e-12 * abs(A.diagonal()).max() * sp.eye(n)))
                op = spla.LinearOperator((n, n), matvec=lambda v: lu.solve(B @ v), dtype=np.float64)
                rs = np.random.RandomState(4321 + r)
                wv, V = spla.eigs(op, k=k, which="LM", v0=rs.uniform(size=n), tol=1e-10)
                order = np.argsort(-np.abs(wv))
                cols = []
                skip = False
                for q in order:  # nonsymmetric pencils (2-D generator quirk): span{Re,Im} of a conjugate pair
                    if skip:
                        skip = False
                        continue
                    if abs(wv[q].imag) > 1e-10 * abs(wv[q]):
                        cols += [np.real(V[:, q]), np.imag(V[:, q])]
                        skip = True
                    else:
                        cols.append(np.real(V[:, q]))
                V = np.stack(cols[:k], axis=1)
            V = V / np.linalg.norm(V, axis=0, keepdims=True)
            V[np.abs(V) < 1.0 / (HPDDM_EPS * HPDDM_PEN)] = 0.0
            Z.append(np.asfortranarray(V))
        self.set_vectors(Z)
        return Z

    def build_coarse(self, lapack_tr_quirk=False):
        """Galerkin coarse operator E = Z^H A Z, block (i,j):
        E_ii = Z_i^H D_i A_i D_i Z_i ; E_ij = Z_i^H D_i R_ij (A_j D_j Z_j)
        (operator.hpp:395-403 applyFromNeighbor, 440-503 C = A*D with columns
        where D<=EPS dropped, 505-528 applyToNeighbor).  Dense N_c x N_c here;
        the reference stores it sparse by blocks and hands it to the coarse
        solver (coarse_operator_impl.hpp:282-1247)."""
        off = np.concatenate([[0], np.cumsum(self.nu)]).astype(int)
        Nc = int(off[-1])
        dtype = np.result_type(*[z.dtype for z in self.Z], *[a.dtype for a in self.A])
        E = np.zeros((Nc, Nc), dtype=dtype)
        W = []
        for j in range(self.P):
            dj = np.where(self.d[j] > HPDDM_EPS, self.d[j], 0.0)
            C = sp.csr_matrix(self.A[j] @ sp.diags(dj))
            W.append(np.asfortranarray(C @ self.Z[j]))  # work_ = A_j D_j Z_j
        for i in range(self.P):
            E[off[i]:off[i + 1], off[i]:off[i + 1]] = self.Z[i].conj().T @ (self.d[i][:, None] * W[i])
            for (j, idx) in self.map[i]:
                kk = [q for q, (rr, _) in enumerate(self.map[j]) if rr == i][0]
                jdx = self.map[j][kk][1]
                tmp = np.zeros((self.n[i], self.nu[j]), dtype=dtype)
                tmp[idx, :] = self.d[i][idx, None] * W[j][jdx, :]
                E[off[i]:off[i + 1], off[j]:off[j + 1]] = self.Z[i].conj().T @ tmp
        self.E = E
        self.off = off
        # Plugin quirk, reproduced only on request: the reference's dense LAPACK coarse solver
        # (LapackTR, HPDDM_LAPACK.hpp:349-356 via numfact<numbering_, true> at :426) fills its dense
        # array transposed when the coarse pattern is not full, and then calls getrs("N"): it solves
        # E^T y = rhs.  Invisible for symmetric E (every SPD configuration), visible with the
        # non-symmetric matrices of the 2-D generator on > 4 ranks.  The sparse coarse solvers
        # (MUMPS / SuiteSparse) and this repo's CUDA path solve E y = rhs.
        full = all(len(self.map[i]) == self.P - 1 for i in range(self.P))
        self.Elu = sla.lu_factor(E.T if (lapack_tr_quirk and not full) else E)
        return E

    def call_solver(self, uc):
        """CoarseOperator::callSolver (coarse_operator_impl.hpp:1630-1732):
        gather nu_i x mu blocks -> N_c x mu RHS, solve E y = rhs, scatter."""
        rhs = np.concatenate(uc, axis=0)
        y = sla.lu_solve(self.Elu, rhs)
        return [np.asfortranarray(y[self.off[r]:self.off[r + 1], :]) for r in range(self.P)]

    def deflation(self, x):
        """schwarz.hpp:1602-1622: out = exchange(Z E^{-1} Z^H D in)  (gemm with Wrapper<K>::transc, 1616)."""
        uc = [self.Z[r].conj().T @ (self.d[r][:, None] * x[r]) for r in range(self.P)]
        uc = self.call_solver(uc)
        out = [np.asfortranarray(self.Z[r] @ uc[r]) for r in range(self.P)]
        return self.exchange(out)

    # ------------------------------------------------------------------- apply
    def apply(self, x, correction=None):
        """Schwarz::apply (schwarz.hpp:527-612), every branch; `x` is not
        modified (the reference clobbers `in` when no workspace is passed)."""
        P = self.P
        if self.E is None or correction is None:
            if self.type == NO:
                return [v.copy() for v in x]
            if self.type in (GE, OG):
                out = [self.solver[r].solve(x[r]) for r in range(P)]
                return self.exchange(out)
            if self.type == OS:
                out = [self.solver[r].solve(self.d[r][:, None] * x[r]) * self.d[r][:, None] for r in range(P)]
            else:
                out = [self.solver[r].solve(x[r]) for r in range(P)]
            return self.subdomain_exchange(out)
        work = [np.array(v, order="F", copy=True) for v in x]
        if correction == ADDITIVE:
            out = self.deflation(x)
            for r in range(P):
                out[r] += self.solver[r].solve(work[r])
            return self.exchange(out)
        out = self.deflation(x)
        for r in range(P):
            work[r] -= self.A[r] @ out[r]          # csrmm alpha=-1, beta=1 (581-586)
        self.exchange(work)                        # (I - A Q) in, consistent (588)
        if self.type == OS:
            for r in range(P):
                work[r] *= self.d[r][:, None]
        work = [self.solver[r].solve(work[r]) for r in range(P)]  # (590)
        self.exchange(work)                        # (591)
        if correction == BALANCED:                 # (593-606)
            tmp = self.GMV(work)
            tmp = self.deflation(tmp)
            for r in range(P):
                work[r] -= tmp[r]
        for r in range(P):
            out[r] += work[r]                      # (607)
        return out

    # --------------------------------------------------------------- residual
    def compute_residual(self, x, f, norm="l2"):
        """schwarz.hpp:761-803: returns (||f||, ||A x - f||) per column; D-weighted l2 / l1 sums or the plain l-infinity norm.  Rows
        that carry a boundary condition are left out of the residual, entries of f larger than EPS * PEN are divided by PEN."""
        tmp = self.GMV(x)
        mu = x[0].shape[1]
        st = np.zeros((mu, 2))
        for r in range(self.P):
            tmp[r] -= f[r]
            bc = self.boundary_conditions(r)
            notb = np.ones(self.n[r])
            for i in bc:
                notb[i] = 0.0
            fr = np.abs(np.where(np.abs(f[r]) > HPDDM_EPS * HPDDM_PEN, f[r] / HPDDM_PEN, f[r]))
            tr = np.abs(tmp[r]) * notb[:, None]
            if norm == "l2":
                st[:, 1] += self.d[r] @ tr ** 2
                st[:, 0] += self.d[r] @ fr ** 2
            elif norm == "l1":
                st[:, 1] += self.d[r] @ tr
                st[:, 0] += self.d[r] @ fr
            else:
                st[:, 1] = np.maximum(st[:, 1], tr.max(axis=0))
                st[:, 0] = np.maximum(st[:, 0], fr.max(axis=0))
        return np.sqrt(st) if norm == "l2" else st

    def rhs_norm(self, b):
        """||b|| of IterativeMethod::initializeNorm, right-preconditioned branch (iterative.hpp:455-468): D-weighted l2 norm per
        column in which an entry on a boundary-condition row (Subdomain::boundaryConditions) larger than PEN * EPS is first
        divided by HPDDM_PEN -- a penalised Dirichlet value b_i = 1e30 g_i counts as g_i."""
        mu = b[0].shape[1]
        out = np.zeros(mu)
        for r in range(self.P):
            br = np.array(b[r], copy=True)
            for i in self.boundary_conditions(r):
                big = np.abs(br[i, :]) > HPDDM_PEN * HPDDM_EPS
                br[i, big] = br[i, big] / HPDDM_PEN
            out += self.d[r] @ (np.abs(br) ** 2)
        return np.sqrt(out)

    # ------------------------------------------------------------------ helpers
    def dot(self, x, y):
        """D-weighted global inner products per column (iterative.hpp:455-468,
        GMRES.hpp:59-68): sum_r sum_i d_i conj(x_i) y_i  (iterative.hpp:503)."""
        return sum((self.d[r][:, None] * np.conj(x[r]) * y[r]).sum(axis=0) for r in range(self.P))
