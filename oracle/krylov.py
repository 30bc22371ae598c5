"""Krylov callers of the hot path, restated (TEST INFRASTRUCTURE ONLY).

``gmres`` follows ``IterativeMethod::GMRES`` (include/HPDDM_GMRES.hpp:31-158)
with the reference defaults (include/HPDDM_iterative.hpp:197-212): right
preconditioning, classical Gram-Schmidt, restart 40, tol 1e-6, max_it 100,
D-weighted inner products, convergence test |s_i| / ||b||_D <= tol
(iterative.hpp:98-127,455-468), Givens-rotated Hessenberg (Arnoldi,
iterative.hpp:669-710), solution update x += M^{-1} (V y) at restart/exit
(updateSol/addSol, iterative.hpp:272-336).

The operator is duck-typed like the reference's ``Operator`` concept
(GMRES.hpp:57-62,113-117): it needs ``start(b,x)``, ``apply(v)``, ``GMV(v)``,
``dot(x,y)``.  Vectors are per-rank lists of (n_loc, mu) arrays; every column
runs its own (pseudo-block) GMRES exactly like the reference's non-block
driver.  Returns (iterations, x, applies).
"""
import numpy as np


def _axpy(a, x, y):
    for r in range(len(x)):
        y[r] += x[r] * a


def _rhs_norm(op, b):
    """initializeNorm (iterative.hpp:441-470): operators that know their boundary conditions rescale penalised entries."""
    if hasattr(op, "rhs_norm"):
        return op.rhs_norm(b)
    return np.sqrt(np.real(op.dot(b, b)))


def _converged(res, norm, tol):
    """checkConvergence (iterative.hpp:98-103): tol > 0 relative to ||b||, tol < 0 absolute."""
    return res / norm <= tol if tol > 0 else res <= -tol


def gmres(op, b, x0=None, tol=1e-6, max_it=100, restart=40, verbose=False):
    P = len(b)
    mu = b[0].shape[1]
    x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [np.array(v, order="F", copy=True) for v in x0]
    x = op.start(b, x)                                   # initializeNorm -> A.start (iterative.hpp:444)
    dtype = np.result_type(*[v.dtype for v in b])        # K: float64 or complex128
    norm = _rhs_norm(op, b)                              # ||b||_D, penalised rows / PEN (iterative.hpp:455-468)
    norm = np.where(norm < 1e-12, 1.0, norm)             # GMRES.hpp:73
    m = restart
    applies = 0
    conv = np.full(mu, -m, dtype=int)                    # hasConverged sentinel
    j = 1
    while j <= max_it:
        Ax = op.GMV(x)
        v = [[b[r] - Ax[r] for r in range(P)]]           # right variant: v0 = b - A x
        sn0 = np.real(op.dot(v[0], v[0]))
        if j == 1 and np.any(sn0 < np.finfo(float).eps ** 2):
            j = 0
            break
        s = np.zeros((m + 1, mu), dtype=dtype)
        s[0] = np.sqrt(sn0)
        for r in range(P):
            v[0][r] = v[0][r] / np.sqrt(sn0)
        conv[conv > 0] = 0
        H = np.zeros((m + 1, m, mu), dtype=dtype)
        cs = np.zeros((m, mu), dtype=dtype)              # cosine lives in K, sine is real (iterative.hpp:690-710)
        sn = np.zeros((m, mu))
        i = 0
        done = False
        while i < m and j <= max_it:
            z = op.apply(v[i])                           # M^{-1} v_i       (GMRES.hpp:116)
            applies += 1
            w = op.GMV(z)                                # A M^{-1} v_i     (GMRES.hpp:117)
            # classical Gram-Schmidt (iterative.hpp:489-540): all dots first
            h = np.array([op.dot(v[k], w) for k in range(i + 1)])
            for k in range(i + 1):
                for r in range(P):
                    w[r] -= v[k][r] * h[k]
            hn = np.sqrt(np.real(op.dot(w, w)))
            H[:i + 1, i] = h
            H[i + 1, i] = hn
            if i < m - 1:
                for r in range(P):
                    w[r] = w[r] / np.where(hn == 0, 1.0, hn)
            v.append(w)
            for k in range(i):                           # previous rotations
                g = np.conj(cs[k]) * H[k, i] + sn[k] * H[k + 1, i]
                H[k + 1, i] = -sn[k] * H[k, i] + cs[k] * H[k + 1, i]
                H[k, i] = g
            delta = np.hypot(np.abs(H[i, i]), np.abs(H[i + 1, i]))
            sn[i] = np.real(H[i + 1, i]) / delta
            cs[i] = H[i, i] / delta
            H[i, i] = delta
            H[i + 1, i] = 0.0
            s[i + 1] = -sn[i] * s[i]
            s[i] = s[i] * np.conj(cs[i])
            i += 1
            res = np.abs(s[i])
            newly = (conv == -m) & _converged(res, norm, tol)
            conv[newly] = i
            if verbose:
                print(f"GMRES: {j:3d} {res.max():.6e} {(res / norm).max():.6e} < {tol}")
            if not np.any(conv == -m):
                done = True
                break
            j += 1
        # updateSol: y = H^{-1} s per column with its own dimension
        work = [np.zeros_like(x[r]) for r in range(P)]
        for nu in range(mu):
            dim = abs(conv[nu]) if conv[nu] != 0 else 0
            if not done and conv[nu] == -m:
                dim = i
            if dim == 0:
                continue
            y = np.linalg.solve(np.triu(H[:dim, :dim, nu]), s[:dim, nu])
            for k in range(dim):
                for r in range(P):
                    work[r][:, nu] += v[k][r][:, nu] * y[k]
        corr = op.apply(work)                            # addSol right variant (iterative.hpp:314-329)
        applies += 1
        for r in range(P):
            x[r] += corr[r]
        if done or j > max_it:
            break
    return min(j, max_it), x, applies


def cg(op, b, x0=None, tol=1e-6, max_it=100):
    """IterativeMethod::CG (include/HPDDM_CG.hpp:31-168), non-flexible variant, restated: D-weighted products
    sum_i d_i conj(x_i) y_i (Blas::dot conjugates its first argument for complex K, HPDDM_BLAS.hpp:173-179), all columns
    advance together, a converged column stops being updated (CG.hpp:99-105), convergence when
    ||M^-1 r||_D / ||M^-1 r_0||_D <= tol (CG.hpp:68,139-140).  Returns (iterations, x)."""
    P = len(b)
    mu = b[0].shape[1]
    x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [np.array(v, order="F", copy=True) for v in x0]
    x = op.start(b, x)
    z = op.GMV(x)
    r = [b[q] - z[q] for q in range(P)]
    p = op.apply(r)
    dirn = np.real(op.dot(p, p))
    res = np.sqrt(dirn)
    conv = np.full(mu, -max_it, dtype=int)
    i = 0
    if np.all(dirn >= np.finfo(float).eps ** 2):
        zd = p                                               # `trash` = D zd : p before the loop, z = M^-1 r afterwards
        while i < max_it:
            rz = np.real(op.dot(r, zd))                      # CG.hpp:87
            z = op.GMV(p)                                    # CG.hpp:88
            pAp = np.real(op.dot(z, p))                      # CG.hpp:90
            i += 1
            for nu in range(mu):
                if conv[nu] == -max_it:
                    alpha = rz[nu] / pAp[nu]
                    for q in range(P):
                        x[q][:, nu] += alpha * p[q][:, nu]
                        r[q][:, nu] -= alpha * z[q][:, nu]
            z = op.apply(r)                                  # CG.hpp:107
            zd = z
            beta = np.real(op.dot(r, z)) / rz                # CG.hpp:110
            dirn = np.real(op.dot(z, z))                     # CG.hpp:111
            for q in range(P):
                p[q] = z[q] + p[q] * beta[None, :]           # CG.hpp:115
            newly = (conv == -max_it) & (np.sqrt(dirn) / res <= tol)
            conv[newly] = i
            if not np.any(conv == -max_it):
                i -= 1
                break
    else:
        i = -1
    i += 1
    return min(i, max_it), x


def _gram(op, X, Y):
    """sum over ranks of X_r^H D_r Y_r  (mu_x x mu_y): the reduction of blockOrthogonalization / VR (iterative.hpp:523-583)"""
    return op.gram(X, Y)


def bgmres(op, b, x0=None, tol=1e-6, max_it=100, restart=40):
    """IterativeMethod::BGMRES (include/HPDDM_GMRES.hpp:160-313) with the reference defaults (iterative.hpp:192-212): right
    preconditioning, block classical Gram-Schmidt (blockOrthogonalization, iterative.hpp:523-556), CholQR of every new block
    (QR / VR, iterative.hpp:560-583,623-640), no deflation of right-hand sides (deflation_tol = -1), block Hessenberg reduced
    by LAPACK Householder QRs of 2mu x mu blocks (BlockArnoldi, iterative.hpp:714-737: mqr of the previous blocks, geqrf, mqr
    on s), convergence when, for every column nu, the norm of the FIRST nu+1 ENTRIES of column nu of the trailing mu x mu block
    of s, over ||b_nu||_D, is <= tol (checkBlockConvergence, iterative.hpp:139-146 -- a partial norm, restated as is),
    solution update by trtrs + gemm + one preconditioner apply (updateSol, iterative.hpp:272-336).  Returns (iterations, x)."""
    from scipy.linalg import lapack, solve_triangular
    P = len(b)
    mu = b[0].shape[1]
    dtype = np.result_type(*[v.dtype for v in b])
    cplx = np.issubdtype(dtype, np.complexfloating)
    geqrf = lapack.zgeqrf if cplx else lapack.dgeqrf
    mqr = lapack.zunmqr if cplx else lapack.dormqr
    x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [np.array(v, order="F", copy=True) for v in x0]
    x = op.start(b, x)
    norm = _rhs_norm(op, b)
    norm = np.where(norm < 1e-12, 1.0, norm)
    m = restart
    ldh = mu * (m + 1)

    def cholqr(W, update=True):
        G = _gram(op, W, W)
        try:
            R = np.linalg.cholesky(G).conj().T          # potrf("U"): G = R^H R
        except np.linalg.LinAlgError:
            return None
        if update:
            Rinv = solve_triangular(R, np.eye(mu, dtype=dtype))
            for r in range(P):
                W[r] = np.asfortranarray(W[r] @ Rinv)   # trsm("R", "U", "N", "N")
        return R

    def apply_qt(k, C):
        """C[k*mu:(k+2)*mu] <- Q_k^H C[...]  (mqr "L", transc) with the reflectors stored by geqrf in block column k of H"""
        blk = np.asfortranarray(H[k * mu:(k + 2) * mu, k * mu:(k + 1) * mu])
        out, _, info = mqr("L", "C" if cplx else "T", blk, tau[k], np.asfortranarray(C[k * mu:(k + 2) * mu]), 64 * mu)
        C[k * mu:(k + 2) * mu] = out

    j = 1
    dim = 0
    V = None
    H = None
    sv = None
    while j <= max_it:
        Ax = op.GMV(x)
        v0 = [np.asfortranarray(b[r] - Ax[r]) for r in range(P)]
        R0 = cholqr(v0)
        if R0 is None:
            raise RuntimeError("BGMRES: rank-deficient block residual (the reference falls back to GMRES)")
        V = [v0]
        H = np.zeros((ldh, m * mu), dtype=dtype)
        tau = [None] * m
        sv = np.zeros((ldh, mu), dtype=dtype)
        sv[:mu] = np.triu(R0)
        i = 0
        done = False
        while i < m and j <= max_it:
            z = op.apply(V[i])
            w = op.GMV(z)
            Hc = np.zeros((ldh, mu), dtype=dtype)
            # blockOrthogonalization, classical: all products first
            prods = [_gram(op, V[k], w) for k in range(i + 1)]
            for k in range(i + 1):
                Hc[k * mu:(k + 1) * mu] = prods[k]
                for r in range(P):
                    w[r] = np.asfortranarray(w[r] - V[k][r] @ prods[k])
            R = cholqr(w, update=i < m - 1)
            if R is None:
                raise RuntimeError("BGMRES: breakdown in BlockArnoldi")
            Hc[(i + 1) * mu:(i + 2) * mu] = np.triu(R)
            V.append(w)
            for k in range(i):
                apply_qt(k, Hc)
            blk, t, _, info = geqrf(np.asfortranarray(Hc[i * mu:(i + 2) * mu]))
            Hc[i * mu:(i + 2) * mu] = blk
            H[:, i * mu:(i + 1) * mu] = Hc
            tau[i] = t
            apply_qt(i, sv)
            i += 1
            res = sv[i * mu:(i + 1) * mu]
            pt = np.array([np.linalg.norm(res[:nu + 1, nu]) for nu in range(mu)])
            if np.all(_converged(pt, norm, tol)):
                dim = mu * i
                done = True
                break
            j += 1
        if not done:
            dim = mu * i                                   # restart (i == m) or iteration limit (i = max_it % m or m)
        if dim > 0:
            Y = solve_triangular(np.triu(H[:dim, :dim]), sv[:dim])
            work = [np.zeros_like(x[r]) for r in range(P)]
            for k in range(dim // mu):
                for r in range(P):
                    work[r] += V[k][r] @ Y[k * mu:(k + 1) * mu]
            corr = op.apply(work)
            for r in range(P):
                x[r] += corr[r]
        if done or j > max_it:
            break
    return min(j, max_it), x


class OracleOperator:
    """Adapter SchwarzWorld -> Krylov operator concept."""

    def __init__(self, world, correction=None):
        self.w = world
        self.correction = correction

    def start(self, b, x):
        return self.w.start(b, x)

    def apply(self, v):
        return self.w.apply(v, self.correction)

    def GMV(self, v):
        return self.w.GMV(v)

    def dot(self, x, y):
        return self.w.dot(x, y)

    def rhs_norm(self, b):
        return self.w.rhs_norm(b)

    def gram(self, X, Y):
        return sum(X[r].conj().T @ (self.w.d[r][:, None] * Y[r]) for r in range(self.w.P))
