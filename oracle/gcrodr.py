"""GCRO-DR (the Krylov method of BASELINE config 5), restated (TEST INFRASTRUCTURE ONLY).

``gcrodr`` follows ``IterativeMethod::GCRODR`` (include/HPDDM_GCRODR.hpp:35-444) with the reference defaults
(include/HPDDM_iterative.hpp:197-218): right preconditioning, classical Gram-Schmidt, CholQR, recycle target SM
(harmonic Ritz values of smallest magnitude), recycle strategy A, ``recycle_same_system`` 0.  Like the reference's
non-block driver every right-hand side has its own Krylov space, Hessenberg matrix and recycled pair (U, C); the
preconditioner and the operator are applied to all columns at once.

Structure of the reference, kept:

* first cycle of the first solve = GMRES(m) (GCRODR.hpp:179-214 with U == nullptr); afterwards k harmonic Ritz vectors of
  the (unrotated) Hessenberg matrix are turned into the pair  C = A M^-1 U,  C^H D C = I  (GCRODR.hpp:242-316);
* later cycles: Arnoldi on (I - C C^H D) A M^-1 with m - k new vectors (GCRODR.hpp:187-196, Arnoldi with `shift`),
  solution update through  x += M^-1 (U (C^H D r - B y) + V y)  (updateSolRecycling, iterative.hpp:338-393), then a new
  pair from the generalised harmonic Ritz problem  G^H G z = theta G^H W^H D [U~ V] z  (GCRODR.hpp:317-430);
* a later solve starts from the stored U: C = A M^-1 U re-orthonormalised by CholQR, the initial residual is projected
  (GCRODR.hpp:94-130).

Vectors are per-rank lists of (n_loc, mu) arrays; ``op`` is the operator concept of oracle/krylov.py.  The pair (U, C) is
returned in ``state`` and passed to the next solve (the reference keeps it in ``A.storage()``, HPDDM_option.hpp:445-454).

Only the span of the selected eigenvectors enters the iteration (a different basis of the same span changes U and C by the
same unitary diagonal factor, which cancels everywhere), so LAPACK ``geev`` / ``ggev`` through scipy stand in for the
reference's ``hseqr`` + ``hsein`` / ``ggev``.  The one exception is inherited from the reference: for real scalars a
complex-conjugate pair that straddles the k-th position is cut in the middle (GCRODR.hpp:283-296,307), and which real
combination survives depends on the eigenvector normalisation of the LAPACK build.
"""
import numpy as np
import scipy.linalg as sla

from oracle.krylov import _converged, _rhs_norm, gmres

_SORT = {
    "SM": lambda z: np.abs(z) ** 2, "LM": lambda z: -np.abs(z) ** 2, "SR": lambda z: z.real, "LR": lambda z: -z.real,
    "SI": lambda z: z.imag, "LI": lambda z: -z.imag,
}


def _order(vals, target):
    """selectNu (include/HPDDM_specifications.hpp:90-123): indices sorted by the recycle target"""
    return sorted(range(len(vals)), key=lambda i: _SORT[target](complex(vals[i])))


def _real_columns(w, X, sel, k, cplx):
    """Eigenvector block handed to the gemm's of GCRODR.hpp:307-308 / 403-419: for complex K the selected columns; for real K
    LAPACK's storage (real part, imaginary part in consecutive columns for a conjugate pair), first k columns."""
    if cplx:
        return X[:, sel[:k]]
    cols = []
    for j in sel:
        if abs(w[j].imag) < 1e-12 * max(1.0, abs(w[j])):
            cols.append(X[:, j].real)
        else:
            # pair (p, p + 1), positive imaginary part first; index p holds Re, index p + 1 holds Im of the eigenvector of w[p]
            p = j if w[j].imag > 0 else j - 1
            cols.append(X[:, p].real if j == p else X[:, p].imag)
    return np.array(cols[:k]).T.reshape(X.shape[0], -1)


def _combine(blocks, coef, nu, out):
    """out[:, nu] = sum_l blocks[l][:, nu] * coef[l] on every rank"""
    for r in range(len(out)):
        acc = np.zeros(out[r].shape[0], dtype=out[r].dtype)
        for l, blk in enumerate(blocks):
            acc += blk[r][:, nu] * coef[l]
        out[r][:, nu] = acc


def gcrodr(op, b, x0=None, tol=1e-6, max_it=100, restart=40, recycle=0, state=None, target="SM", strategy="A", same_system=0, verbose=False):
    """Returns (iterations, x, state).  same_system = value of -hpddm_recycle_same_system as IterativeMethod::options reads it
    (iterative.hpp:217; the reference raises it from 1 to 2 after a converged solve, GCRODR.hpp:435): 0 = the operator may have
    changed (C = A M^-1 U recomputed at the start of a later solve), 1 = same operator, the pair is still built / updated, C^H D r
    is taken as zero in the solution update, 2 = same operator, the stored pair is used as is (no product with A M^-1, no update)."""
    if recycle <= 0:                                              # GCRODR.hpp:50-55
        it, x, _ = gmres(op, b, x0=x0, tol=tol, max_it=max_it, restart=restart)
        return it, x, state
    P = len(b)
    mu = b[0].shape[1]
    dtype = np.result_type(*[v.dtype for v in b])
    cplx = np.issubdtype(dtype, np.complexfloating)
    m = min(restart, max_it)                                      # iterative.hpp:210
    k = min(m - 1, recycle)                                       # iterative.hpp:215
    U = C = None
    if state is not None and state.get("U") is not None:          # GCRODR.hpp:64-69
        assert state["mu"] == mu, "the oracle keeps one recycled pair per right-hand side"
        U, C, k = state["U"], state["C"], state["k"]
    zeros = lambda: [np.zeros((b[r].shape[0], mu), dtype=dtype, order="F") for r in range(P)]
    x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [np.array(v, order="F", copy=True) for v in x0]
    x = op.start(b, x)
    norm = _rhs_norm(op, b)
    norm = np.where(norm < 1e-12, 1.0, norm)
    conv = np.full(mu, -m, dtype=int)
    j = 1
    while j <= max_it:
        shift = i = k if U is not None else 0
        Ax = op.GMV(x)
        v = [None] * (m + 1)
        v[i] = [np.asfortranarray(b[r] - Ax[r]).astype(dtype) for r in range(P)]
        if j == 1 and U is not None and same_system:              # GCRODR.hpp:115-129 with id[4] / 4 != 0: C is still A M^-1 U
            h = np.array([op.dot(C[c], v[i]) for c in range(k)])
            wk = zeros()
            for c in range(k):
                for r in range(P):
                    v[i][r] -= C[c][r] * h[c][None, :]
                    wk[r] += U[c][r] * h[c][None, :]
            corr = op.apply(wk)
            for r in range(P):
                x[r] += corr[r]
        elif j == 1 and U is not None:                            # GCRODR.hpp:94-130: C = A M^-1 U, CholQR, projection
            pt = [op.apply(U[c]) for c in range(k)]
            C = [op.GMV(pt[c]) for c in range(k)]
            G = np.array([[op.dot(C[a], C[c]) for c in range(k)] for a in range(k)])      # k x k x mu
            for nu in range(mu):
                R = np.linalg.cholesky(G[:, :, nu]).conj().T                               # potrf("U")
                Rinv = sla.solve_triangular(R, np.eye(k, dtype=dtype))
                for blk in (C, pt, U):
                    old = [[blk[c][r][:, nu].copy() for r in range(P)] for c in range(k)]
                    for c in range(k):
                        for r in range(P):
                            blk[c][r][:, nu] = sum(old[l][r] * Rinv[l, c] for l in range(k))
            h = np.array([op.dot(C[c], v[i]) for c in range(k)])                           # k x mu
            for c in range(k):
                for r in range(P):
                    v[i][r] -= C[c][r] * h[c][None, :]
                    x[r] += pt[c][r] * h[c][None, :]
        sn0 = np.real(op.dot(v[i], v[i]))
        if j == 1 and np.any(sn0 < np.finfo(float).eps ** 2):
            j = 0
            break
        conv[conv > 0] = 0
        s = np.zeros((m + 1, mu), dtype=dtype)
        s[i] = np.sqrt(sn0)
        resnorm = np.sqrt(sn0)                                    # `sn[nu]`, the `norm` of updateSolRecycling
        for r in range(P):
            v[i][r] = v[i][r] / np.sqrt(sn0)
        R = np.zeros((m + 1, m, mu), dtype=dtype)                 # rotated Hessenberg (rows >= shift), cosines, sines
        cs = np.zeros((m, mu), dtype=dtype)
        sn = np.zeros((m, mu))
        save = np.zeros((m + 1, m, mu), dtype=dtype)              # unrotated Hessenberg of this cycle, column i - shift
        B = np.zeros((max(k, 1), m, mu), dtype=dtype)             # C^H D A M^-1 v_i
        while i < m and j <= max_it:
            z = op.apply(v[i])
            w = op.GMV(z)
            if U is not None:                                     # orthogonalization against C (GCRODR.hpp:191)
                hB = np.array([op.dot(C[c], w) for c in range(k)])
                for c in range(k):
                    for r in range(P):
                        w[r] -= C[c][r] * hB[c][None, :]
                B[:k, i] = hB
            h = np.array([op.dot(v[l], w) for l in range(shift, i + 1)])                   # Arnoldi, classical Gram-Schmidt
            for l in range(shift, i + 1):
                for r in range(P):
                    w[r] -= v[l][r] * h[l - shift][None, :]
            hn = np.sqrt(np.real(op.dot(w, w)))
            if i < m - 1:
                for r in range(P):
                    w[r] = w[r] / hn
            v[i + 1] = w
            save[:i + 1 - shift, i - shift] = h
            save[i + 1 - shift, i - shift] = hn
            R[shift:i + 1, i] = h
            R[i + 1, i] = hn
            for l in range(shift, i):
                g = np.conj(cs[l]) * R[l, i] + sn[l] * R[l + 1, i]
                R[l + 1, i] = -sn[l] * R[l, i] + cs[l] * R[l + 1, i]
                R[l, i] = g
            delta = np.hypot(np.abs(R[i, i]), np.abs(R[i + 1, i]))
            sn[i] = np.real(R[i + 1, i]) / delta
            cs[i] = R[i, i] / delta
            R[i, i] = delta
            R[i + 1, i] = 0.0
            s[i + 1] = -sn[i] * s[i]
            s[i] = s[i] * np.conj(cs[i])
            i += 1
            res = np.abs(s[i])
            conv[(conv == -m) & _converged(res, norm, tol)] = i
            if verbose:
                print(f"GCRODR: {j:3d} {res.max():.6e} {(res / norm).max():.6e} < {tol}")
            if not np.any(conv == -m):
                i += (m - k) if U is not None else m
                break
            j += 1
        if j != max_it + 1 and i == m:
            converged = False
        else:
            converged = True
            if j == max_it + 1:                                   # GCRODR.hpp:223-231
                rem = (max_it - m) % (m - k) if U is not None else max_it % m
                if rem:
                    conv[conv < 0] = rem + (k if U is not None else 0)
        # updateSolRecycling (iterative.hpp:338-393)
        work = zeros()
        for nu in range(mu):
            dim = abs(conv[nu])
            if dim == 0:
                continue
            y = sla.solve_triangular(R[shift:dim, shift:dim, nu], s[shift:dim, nu]) if dim > shift else np.zeros(0, dtype=dtype)
            if U is not None:
                su = -(B[:k, shift:dim, nu] @ y)                    # iterative.hpp:351: `same` drops the C^H D r term
                if not same_system:
                    su = su + np.array([resnorm[nu] * op.dot(C[c], v[shift])[nu] for c in range(k)])
                _combine(U[:k] + v[shift:dim], np.concatenate([su, y]), nu, work)
            else:
                _combine(v[:dim], y, nu, work)
        corr = op.apply(work)
        for nu in range(mu):
            if conv[nu] != 0:
                for r in range(P):
                    x[r][:, nu] += corr[r][:, nu]
        if i == m:                                                # GCRODR.hpp:234-237: the last basis vector is normalised here
            if U is not None:
                i -= k
            for r in range(P):
                v[m][r] = v[m][r] / save[i, i - 1][None, :]
        if same_system > 1:                                       # GCRODR.hpp:239: id[4] / 4 <= 1 guards both branches below
            pass
        elif U is None:                                           # GCRODR.hpp:242-316: first pair from GMRES(m)
            nz = conv[conv != 0]
            dim = abs(int(nz.min())) if len(nz) else 0
            if j < k or dim < k:
                k = dim
            U = [zeros() for _ in range(k)]
            C = [zeros() for _ in range(k)]
            for nu in range(mu):
                hq = cs[dim - 1, nu] / R[dim - 1, dim - 1, nu]
                f = np.zeros(dim, dtype=dtype)
                for l in range(dim - 1, 0, -1):
                    f[l] = cs[l - 1, nu] * hq
                    hq = hq * -sn[l - 1, nu]
                f[0] = hq
                Hbar = save[:dim + 1, :dim, nu]
                Hm = Hbar[:dim].copy()
                Hm[:, dim - 1] += Hbar[dim, dim - 1] * Hbar[dim, dim - 1] * f
                w_, X = sla.eig(Hm)
                q = _order(w_, target)[:k]
                sel = set()
                if cplx:
                    sel = set(q)
                else:                                             # GCRODR.hpp:283-296
                    mm, it_ = 0, 0
                    while it_ < len(q):
                        if abs(w_[q[it_]].imag) < 1e-12:
                            sel.add(q[it_])
                            mm += 1
                        elif mm < k + 1:
                            p = q[it_] if w_[q[it_]].imag > 0 else q[it_] - 1
                            sel.update((p, p + 1))
                            mm += 2
                            it_ += 1
                        else:
                            break
                        it_ += 1
                vr = _real_columns(w_, X, sorted(sel), k, cplx).astype(dtype)
                Q, Rr = np.linalg.qr(Hbar @ vr)
                Y = vr @ np.linalg.inv(Rr)                          # trsm("R", "U"): U = V vr R^-1
                for c in range(k):
                    _combine(v[:dim], Y[:, c], nu, U[c])
                    _combine(v[:dim + 1], Q[:, c], nu, C[c])
        elif j > m - k:                                           # GCRODR.hpp:317-430: new pair from [U, V]
            for nu in range(mu):
                if conv[nu] == 0:
                    continue
                dim = abs(conv[nu])
                diff = dim - k
                W = C[:k] + v[k:dim + 1]                           # dim + 1 vectors
                if strategy == "A":
                    un = np.array([np.real(op.dot(U[c], U[c])[nu]) for c in range(k)])
                    Du = 1.0 / np.sqrt(un)
                    Wt = np.array([[op.dot(W[l], U[c])[nu] for c in range(k)] for l in range(dim + 1)]) * Du[None, :]
                else:
                    Du = np.ones(k)
                Hbar = save[:diff + 1, :diff, nu]
                G = np.zeros((dim + 1, dim), dtype=dtype)
                G[:k, :k] = np.diag(Du)
                G[:k, k:] = B[:k, k:dim, nu]
                G[k:, k:] = Hbar
                Am = G.conj().T @ G
                Bm = np.zeros((dim, dim), dtype=dtype)
                if strategy == "A":
                    Bm[:, :k] = G.conj().T @ Wt
                else:                                             # strategy B (GCRODR.hpp:376-382): W^H D [U V] replaced by [I 0; 0 I; 0 0]
                    Bm[:k, :k] = np.eye(k)
                    Bm[k:, :k] = B[:k, k:dim, nu].conj().T
                Bm[k:, k:] = Hbar[:diff].conj().T
                w_, X = sla.eig(Am, Bm)
                q = _order(w_, target)[:k]
                if cplx:
                    vr = X[:, q]
                else:                                             # ggev storage: column p = Re, column p + 1 = Im of a pair
                    cols = []
                    for jj in q:
                        if abs(w_[jj].imag) < 1e-12 * max(1.0, abs(w_[jj])) or not np.isfinite(w_[jj]):
                            cols.append(X[:, jj].real)
                        else:
                            p = jj if w_[jj].imag > 0 else jj - 1
                            cols.append(X[:, p].real if jj == p else X[:, p].imag)
                    vr = np.array(cols).T
                vr = vr.astype(dtype)
                Q, Rr = np.linalg.qr(G @ vr)
                Y = vr @ np.linalg.inv(Rr)
                Y[:k] *= Du[:, None]                               # U~ = U Du
                basisU = U[:k] + v[k:dim]
                Un = [zeros() for _ in range(k)]
                Cn = [zeros() for _ in range(k)]
                for c in range(k):
                    _combine(basisU, Y[:, c], nu, Un[c])
                    _combine(W, Q[:, c], nu, Cn[c])
                for c in range(k):
                    for r in range(P):
                        U[c][r][:, nu] = Un[c][r][:, nu]
                        C[c][r][:, nu] = Cn[c][r][:, nu]
        if converged:
            break
    return min(j, max_it), x, dict(U=U, C=C, k=k, mu=mu)


def bgcrodr(op, b, x0=None, tol=1e-6, max_it=100, restart=40, recycle=0, state=None, target="SM", strategy="A", same_system=0, verbose=False):
    """IterativeMethod::BGCRODR (include/HPDDM_GCRODR.hpp:445-907) with the reference defaults (iterative.hpp:192-218): right
    preconditioning, block classical Gram-Schmidt, CholQR, no deflation of right-hand sides (deflation_tol = -1), no enlarged Krylov
    space, recycle target SM, strategy A.  One block Krylov space and ONE recycled pair (U, C) of mu * k columns for all right-hand
    sides: block Arnoldi with Householder-reduced block Hessenberg matrix as in BGMRES (BlockArnoldi, iterative.hpp:714-737, with
    `shift` = k), orthogonalisation of every new block against C, the block versions of the first pair (GCRODR.hpp:674-760), of the
    solution update (updateSolRecycling, iterative.hpp:372-391) and of the generalised harmonic Ritz update (GCRODR.hpp:761-884).
    U, C: lists of k blocks (per-rank lists of n x mu arrays); column c of the n x (mu k) matrix is column c % mu of block c // mu.
    Returns (iterations, x, state)."""
    from scipy.linalg import lapack
    from oracle.krylov import bgmres
    if recycle <= 0:                                              # GCRODR.hpp:460-465
        it, x = bgmres(op, b, x0=x0, tol=tol, max_it=max_it, restart=restart)
        return it, x, state
    P = len(b)
    mu = b[0].shape[1]
    dtype = np.result_type(*[v.dtype for v in b])
    cplx = np.issubdtype(dtype, np.complexfloating)
    geqrf = lapack.zgeqrf if cplx else lapack.dgeqrf
    mqr = lapack.zunmqr if cplx else lapack.dormqr
    m = min(restart, max_it)
    k = min(m - 1, recycle)
    U = C = None
    if state is not None and state.get("U") is not None:
        assert state["mu"] == mu
        U, C, k = state["U"], state["C"], state["k"]
    x = [np.zeros_like(v, order="F") for v in b] if x0 is None else [np.array(v, order="F", copy=True) for v in x0]
    x = op.start(b, x)
    norm = _rhs_norm(op, b)
    norm = np.where(norm < 1e-12, 1.0, norm)
    ldh = mu * (m + 1)

    def gram(X, W):                                               # stacked X_c^H D W, (len(X) mu) x mu
        return np.vstack([op.gram(Xc, W) for Xc in X]) if len(X) else np.zeros((0, mu), dtype=dtype)

    def combine(X, coef):                                         # [X_0 X_1 ...] (n x len(X) mu) times coef -> one block of coef.shape[1] columns
        return [sum(X[c][r] @ coef[c * mu:(c + 1) * mu] for c in range(len(X))) for r in range(P)]

    def split(Mat):                                               # per-rank n x (q mu) -> q blocks
        return [[np.asfortranarray(Mat[r][:, c * mu:(c + 1) * mu]) for r in range(P)] for c in range(Mat[0].shape[1] // mu)]

    def cholqr(W, update=True):
        G = op.gram(W, W)
        try:
            R = np.linalg.cholesky(G).conj().T
        except np.linalg.LinAlgError:
            return None
        if update:
            Rinv = sla.solve_triangular(R, np.eye(mu, dtype=dtype))
            for r in range(P):
                W[r] = np.asfortranarray(W[r] @ Rinv)
        return R

    j = 1
    dim = mu * m
    while j <= max_it:
        shift = k if U is not None else 0
        Ax = op.GMV(x)
        v0 = [np.asfortranarray(b[r] - Ax[r]).astype(dtype) for r in range(P)]
        if j == 1 and U is not None:                              # GCRODR.hpp:515-556
            bK = mu * k
            if not same_system:
                pt = [op.apply(U[c]) for c in range(k)]
                C = [op.GMV(pt[c]) for c in range(k)]
                G = np.hstack([gram(C, C[c]) for c in range(k)])   # bK x bK
                R = np.linalg.cholesky(G).conj().T
                Rinv = sla.solve_triangular(R, np.eye(bK, dtype=dtype))
                C, pt, U = (split(combine(blk, Rinv)) for blk in (C, pt, U))
            Hc = gram(C, v0)
            corr = combine(C, Hc)
            for r in range(P):
                v0[r] -= corr[r]
            if not same_system:
                upd = combine(pt, Hc)
            else:
                upd = op.apply(combine(U, Hc))
            for r in range(P):
                x[r] += upd[r]
        R0 = cholqr(v0)
        if R0 is None:
            raise RuntimeError("BGCRODR: rank-deficient block residual (the reference falls back to GCRODR)")
        v = [None] * (m + 1)
        v[shift] = v0
        H = np.zeros((ldh, m * mu), dtype=dtype)
        tau = [None] * m
        s = np.zeros((ldh, mu), dtype=dtype)
        s[shift * mu:(shift + 1) * mu] = np.triu(R0)
        save = np.zeros((ldh, m * mu), dtype=dtype)               # unreduced block Hessenberg matrix of this cycle, indices relative to shift
        Bm = np.zeros((max(k, 1) * mu, m * mu), dtype=dtype)       # C^H D A M^-1 v_i

        def apply_q(kk, Cm, trans):
            blk = np.asfortranarray(H[kk * mu:(kk + 2) * mu, kk * mu:(kk + 1) * mu])
            out, _, info = mqr("L", ("C" if cplx else "T") if trans else "N", blk, tau[kk], np.asfortranarray(Cm[kk * mu:(kk + 2) * mu]), 64 * mu)
            Cm[kk * mu:(kk + 2) * mu] = out

        i = shift
        conv_now = False
        while i < m and j <= max_it:
            z = op.apply(v[i])
            w = op.GMV(z)
            if U is not None:
                hB = gram(C, w)
                corr = combine(C, hB)
                for r in range(P):
                    w[r] = w[r] - corr[r]
                Bm[:k * mu, i * mu:(i + 1) * mu] = hB
            Hc = np.zeros((ldh, mu), dtype=dtype)
            prods = gram(v[shift:i + 1], w)
            Hc[shift * mu:(i + 1) * mu] = prods
            corr = combine(v[shift:i + 1], prods)
            w = [np.asfortranarray(w[r] - corr[r]) for r in range(P)]
            R = cholqr(w, update=i < m - 1)
            if R is None:
                raise RuntimeError("BGCRODR: breakdown in BlockArnoldi (the reference falls back to GCRODR)")
            Hc[(i + 1) * mu:(i + 2) * mu] = np.triu(R)
            save[:(i + 2 - shift) * mu, (i - shift) * mu:(i - shift + 1) * mu] = Hc[shift * mu:(i + 2) * mu]
            v[i + 1] = w
            for kk in range(shift, i):
                apply_q(kk, Hc, True)
            blk, t, _, info = geqrf(np.asfortranarray(Hc[i * mu:(i + 2) * mu]))
            Hc[i * mu:(i + 2) * mu] = blk
            H[:, i * mu:(i + 1) * mu] = Hc
            tau[i] = t
            apply_q(i, s, True)
            i += 1
            res = s[i * mu:(i + 1) * mu]
            pt_ = np.array([np.linalg.norm(res[:nu + 1, nu]) for nu in range(mu)])
            if verbose:
                print(f"BGCRODR: {j:3d} {(pt_ / norm).max():.6e}")
            if np.all(_converged(pt_, norm, tol)):
                dim = mu * i
                conv_now = True
                break
            j += 1
        if not conv_now and j != max_it + 1 and i == m:
            converged = False
        else:
            converged = True
            if j == max_it + 1:
                rem = (max_it - m) % (m - k) if U is not None else max_it % m
                if rem:
                    dim = mu * (rem + (k if U is not None else 0))
        # updateSolRecycling, block form (iterative.hpp:372-391)
        da = dim - mu * shift
        Y = sla.solve_triangular(np.triu(H[shift * mu:shift * mu + da, shift * mu:shift * mu + da]), s[shift * mu:shift * mu + da]) if da > 0 else np.zeros((0, mu), dtype=dtype)
        if U is not None:
            top = -(Bm[:k * mu, shift * mu:shift * mu + da] @ Y)
            if not same_system:
                r0 = [np.asfortranarray(v[shift][r] @ np.triu(R0)) for r in range(P)]      # the block residual V_k R_0
                top = top + gram(C, r0)
            work = combine(U[:k] + v[shift:shift + da // mu], np.vstack([top, Y]))
        else:
            work = combine(v[:da // mu], Y)
        corr = op.apply(work)
        for r in range(P):
            x[r] += corr[r]
        if not conv_now and i == m:                               # GCRODR.hpp:658-661: the last block is normalised here
            irel = m - shift
            Rl = np.triu(save[irel * mu:(irel + 1) * mu, (irel - 1) * mu:irel * mu])
            Rli = sla.solve_triangular(Rl, np.eye(mu, dtype=dtype))
            v[m] = [np.asfortranarray(v[m][r] @ Rli) for r in range(P)]
        if same_system > 1:
            pass
        elif U is None:                                           # GCRODR.hpp:674-760
            db = min(j, m)
            if db < k:
                k = db
            bK = mu * k
            df = db * mu
            Rd = save[db * mu:(db + 1) * mu, (db - 1) * mu:db * mu]
            Rx = save[db * mu:(db + 1) * mu, (m - 1) * mu:m * mu]  # GCRODR.hpp:682: block column m - 1 (zero when the cycle stopped before m)
            sb = np.zeros((ldh, mu), dtype=dtype)
            sb[(db - 1) * mu:db * mu] = Rd.conj().T @ Rx
            sb[:df] = sla.solve_triangular(np.triu(H[:df, :df]), sb[:df], trans="C")
            for kk in range(db - 1, -1, -1):
                apply_q(kk, sb, False)
            Hbar = save[:df + mu, :df].copy()
            Mh = Hbar[:df].copy()
            Mh[:, df - mu:df] += sb[:df]
            w_, X = sla.eig(Mh)
            q = _order(w_, target)
            vr = _ggev_columns(w_, X, q, bK, cplx).astype(dtype)
            Q, Rr = np.linalg.qr(Hbar @ vr)
            Yc = vr @ np.linalg.inv(Rr)
            U = split(combine(v[:db], Yc))
            C = split(combine(v[:db + 1], Q))
        elif j > m - k:                                           # GCRODR.hpp:761-884
            bK = mu * k
            diff = dim - bK
            nb = diff // mu
            W = C[:k] + v[k:k + nb + 1]
            if strategy == "A":
                un = np.concatenate([np.real(np.diag(op.gram(U[c], U[c]))) for c in range(k)])
                Du = 1.0 / np.sqrt(un)
                Wt = np.hstack([gram(W, U[c]) for c in range(k)]) * Du[None, :]
            else:
                Du = np.ones(bK)
            Hbar = save[:diff + mu, :diff]
            G = np.zeros((dim + mu, dim), dtype=dtype)
            G[:bK, :bK] = np.diag(Du)
            G[:bK, bK:] = Bm[:bK, k * mu:k * mu + diff]
            G[bK:, bK:] = Hbar
            Am = G.conj().T @ G
            Bmat = np.zeros((dim, dim), dtype=dtype)
            if strategy == "A":
                Bmat[:, :bK] = G.conj().T @ Wt
            else:
                Bmat[:bK, :bK] = np.eye(bK)
                Bmat[bK:, :bK] = Bm[:bK, k * mu:k * mu + diff].conj().T
            Bmat[bK:, bK:] = Hbar[:diff].conj().T
            w_, X = sla.eig(Am, Bmat)
            q = _order(w_, target)
            vr = _ggev_columns(w_, X, q, bK, cplx).astype(dtype)
            Q, Rr = np.linalg.qr(G @ vr)
            Yc = vr @ np.linalg.inv(Rr)
            Yc[:bK] *= Du[:, None]
            Un = split(combine(U[:k] + v[k:k + nb], Yc))
            Cn = split(combine(W, Q))
            U, C = Un, Cn
        if converged:
            break
    return min(j, max_it), x, dict(U=U, C=C, k=k, mu=mu)


def _ggev_columns(w, X, q, count, cplx):
    """first `count` entries of the ordering q as LAPACK's real eigenvector storage (column p = Re, p + 1 = Im of a conjugate pair)"""
    if cplx:
        return X[:, q[:count]]
    cols = []
    for jj in q[:count]:
        if abs(w[jj].imag) < 1e-12 * max(1.0, abs(w[jj])) or not np.isfinite(w[jj]):
            cols.append(X[:, jj].real)
        else:
            p = jj if w[jj].imag > 0 else jj - 1
            cols.append(X[:, p].real if jj == p else X[:, p].imag)
    return np.array(cols).T
