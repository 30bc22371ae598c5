"""Problem generators (driver side, like the reference's examples/generate.cpp): synthetic
inputs shared by the CUDA path, the oracle and the benchmarks.

``generate2d`` restates ``examples/generate.cpp:43-311`` of the reference
(2-D cell-centred 5-point Poisson on [0,10]^2, xGrid x yGrid ranks, `overlap`
extra cells on every interior side, ramp partition of unity, neighbour lists
enumerated in global lexicographic order), *including* its quirk that the
"row below/above" column offset is ``Nx/xGrid`` and not the local row length
(``generate.cpp:202,219,233``) -- reproduced, not fixed.

``generate3d`` is the 3-D generalisation used by BASELINE configs 2/3
(7-point stencil, px x py x pz ranks) following the same conventions
(SURVEY.md section 7 step 1); here the stencil offsets are the true local
strides.

Both return plain numpy/scipy objects, 'C' numbering.
"""
import numpy as np
import scipy.sparse as sp


def _split_grid_2d(size):
    # generate.cpp:51-53
    xg = int(np.sqrt(size))
    while size % xg != 0:
        xg -= 1
    return xg, size // xg


def _ramp(lo_art, hi_art, start, end, overlap):
    """distance (in cells) to the artificial faces of one axis, +inf if none.
    generate.cpp:96-186: d = layer_index / overlap on the first `overlap`
    layers next to an artificial (interior) face."""
    idx = np.arange(start, end)
    dist = np.full(idx.shape, np.inf)
    if lo_art:
        dist = np.minimum(dist, idx - start)
    if hi_art:
        dist = np.minimum(dist, end - 1 - idx)
    return dist


def generate2d(rank, size, Nx=100, Ny=100, overlap=1, mu=0, sym=False, neumann=False, seed=1234):
    """examples/generate.cpp:43-311.  Returns dict with keys
    o (neighbour ranks, in the order the reference pushes them), mapping (list
    of int32 arrays), ndof, Mat (csr; lower triangle with diagonal last in row if
    sym), MatNeumann (csr or None), d, f (ndof x max(mu,1), F-order), box."""
    xGrid, yGrid = _split_grid_2d(size)
    y = rank // xGrid
    x = rank - xGrid * y
    iStart = max(x * Nx // xGrid - overlap, 0)
    iEnd = min((x + 1) * Nx // xGrid + overlap, Nx)
    jStart = max(y * Ny // yGrid - overlap, 0)
    jEnd = min((y + 1) * Ny // yGrid + overlap, Ny)
    w, h = iEnd - iStart, jEnd - jStart
    ndof = w * h
    dx = 10.0 / Nx
    dy = 10.0 / Ny
    # --- right-hand side (generate.cpp:69-92)
    if mu == 0:
        xsc, ysc = [6.5, 2.0, 7.0], [8.0, 7.0, 3.0]
        rsc, asc = [0.3, 0.3, 0.4], [0.3, 0.2, -0.1]
        ii, jj = np.meshgrid(np.arange(iStart, iEnd), np.arange(jStart, jEnd))
        xx = dx * (ii + 0.5)
        yy = dy * (jj + 0.5)
        frs = np.ones_like(xx)
        for n in range(3):
            xd, yd = xx - xsc[n], yy - ysc[n]
            m = np.sqrt(xd * xd + yd * yd) <= rsc[n]
            frs = frs - np.where(m, asc[n] * np.cos(0.5 * np.pi * xd / rsc[n]) * np.cos(0.5 * np.pi * yd / rsc[n]), 0.0)
        f = frs.reshape(-1, 1).copy(order="F")
    else:
        # the reference draws from std::random_device (not reproducible); the
        # oracle uses a seeded stream instead
        rs = np.random.RandomState(seed + rank)
        f = np.asfortranarray(rs.uniform(0.0, 1.0, size=(mu, ndof)).T)
    # --- partition of unity ramp (generate.cpp:94-186)
    distx = _ramp(iStart != 0, iEnd != Nx, iStart, iEnd, overlap)
    disty = _ramp(jStart != 0, jEnd != Ny, jStart, jEnd, overlap)
    dist = np.minimum(distx[None, :], disty[:, None])
    d = np.minimum(dist / overlap, 1.0).reshape(-1)
    # --- neighbours + mappings, pushed in increasing-rank order
    o, mapping = [], []
    loc = np.arange(ndof).reshape(h, w)
    for dyy in (-1, 0, 1):
        for dxx in (-1, 0, 1):
            if dxx == 0 and dyy == 0:
                continue
            nx_, ny_ = x + dxx, y + dyy
            if not (0 <= nx_ < xGrid and 0 <= ny_ < yGrid):
                continue
            cols = slice(0, w) if dxx == 0 else (slice(0, 2 * overlap) if dxx < 0 else slice(w - 2 * overlap, w))
            rows = slice(0, h) if dyy == 0 else (slice(0, 2 * overlap) if dyy < 0 else slice(h - 2 * overlap, h))
            o.append(ny_ * xGrid + nx_)
            mapping.append(loc[rows, cols].reshape(-1).astype(np.int32))
    # --- matrix (generate.cpp:188-243), quirk: vertical offset = Nx / xGrid
    off = Nx // xGrid
    rows_, cols_, vals_ = [], [], []
    rowsN, colsN, valsN = [], [], []
    k = 0
    for j in range(jStart, jEnd):
        for i in range(iStart, iEnd):
            ent = []
            if j > jStart:
                ent.append((k - off, -1 / (dy * dy), -1 / (dx * dx) if i == iStart else 0.0))
            if i > iStart:
                ent.append((k - 1, -1 / (dx * dx), -1 / (dy * dy) if j == jStart else 0.0))
            ent.append((k, 2 / (dx * dx) + 2 / (dy * dy), 0.0))
            if not sym:
                if i < iEnd - 1:
                    ent.append((k + 1, -1 / (dx * dx), -1 / (dy * dy) if j == jEnd - 1 else 0.0))
                if j < jEnd - 1:
                    ent.append((k + off, -1 / (dy * dy), -1 / (dx * dx) if i == iEnd - 1 else 0.0))
            for c, v, _ in ent:
                rows_.append(k)
                cols_.append(c)
                vals_.append(v)
            if neumann:
                # generate.cpp:245-297 (both branches produce the same full matrix)
                entN = list(ent)
                if sym:
                    if i < iEnd - 1:
                        entN.append((k + 1, -1 / (dx * dx), -1 / (dy * dy) if j == jEnd - 1 else 0.0))
                    if j < jEnd - 1:
                        entN.append((k + off, -1 / (dy * dy), -1 / (dx * dx) if i == iEnd - 1 else 0.0))
                for c, v, extra in entN:
                    rowsN.append(k)
                    colsN.append(c)
                    valsN.append(v + extra)
            k += 1
    Mat = _csr_keep_order(rows_, cols_, vals_, ndof)
    MatN = _csr_keep_order(rowsN, colsN, valsN, ndof) if neumann else None
    return dict(o=o, mapping=mapping, ndof=ndof, Mat=Mat, MatNeumann=MatN, d=d, f=f, sym=sym,
                box=((iStart, iEnd), (jStart, jEnd)), grid=(xGrid, yGrid), dims=(w, h, 1))


def _csr_keep_order(r, c, v, n):
    """CSR with entries kept in insertion order (no sorting / duplicate merge);
    out-of-range columns produced by the reference quirk cannot occur because
    the reference guards with j>jStart / j<jEnd-1."""
    r = np.asarray(r, dtype=np.int64)
    ia = np.zeros(n + 1, dtype=np.int32)
    np.add.at(ia, r + 1, 1)
    ia = np.cumsum(ia).astype(np.int32)
    A = sp.csr_matrix((np.asarray(v, dtype=np.float64), np.asarray(c, dtype=np.int32), ia), shape=(n, n))
    return A


def split_grid_3d(size):
    """px*py*pz = size, as cubic as possible, px >= py >= pz (1,2,4,8 ->
    1x1x1, 2x1x1, 2x2x1, 2x2x2: SURVEY.md section 8e)."""
    best = None
    for pz in range(1, size + 1):
        if size % pz:
            continue
        for py in range(pz, size // pz + 1):
            if (size // pz) % py:
                continue
            px = size // (pz * py)
            if px < py:
                continue
            cost = px - pz
            if best is None or cost < best[0]:
                best = (cost, (px, py, pz))
    return best[1]


def generate3d(rank, size, N=(16, 16, 16), overlap=1, mu=1, grid=None, neumann=False, seed=1234, sym=False):
    """3-D 7-point cell-centred Poisson on [0,10]^3, N = (Nx,Ny,Nz) cells,
    grid = (px,py,pz) ranks (x fastest in rank numbering, as generate.cpp:55-56
    does in 2-D).  Same conventions as generate2d; true local strides."""
    Nx, Ny, Nz = N
    px, py, pz = grid if grid is not None else split_grid_3d(size)
    assert px * py * pz == size
    z = rank // (px * py)
    y = (rank - z * px * py) // px
    x = rank - z * px * py - y * px
    st = [max(x * Nx // px - overlap, 0), max(y * Ny // py - overlap, 0), max(z * Nz // pz - overlap, 0)]
    en = [min((x + 1) * Nx // px + overlap, Nx), min((y + 1) * Ny // py + overlap, Ny), min((z + 1) * Nz // pz + overlap, Nz)]
    w, h, t = en[0] - st[0], en[1] - st[1], en[2] - st[2]
    ndof = w * h * t
    hx, hy, hz = 10.0 / Nx, 10.0 / Ny, 10.0 / Nz
    rs = np.random.RandomState(seed + rank)
    f = np.asfortranarray(rs.uniform(0.0, 1.0, size=(max(mu, 1), ndof)).T)
    dist = np.minimum(np.minimum(_ramp(st[0] != 0, en[0] != Nx, st[0], en[0], overlap)[None, None, :],
                                 _ramp(st[1] != 0, en[1] != Ny, st[1], en[1], overlap)[None, :, None]),
                      _ramp(st[2] != 0, en[2] != Nz, st[2], en[2], overlap)[:, None, None])
    d = np.minimum(dist / overlap, 1.0).reshape(-1)
    loc = np.arange(ndof).reshape(t, h, w)
    o, mapping = [], []
    dims = (w, h, t)
    pos = (x, y, z)
    pg = (px, py, pz)
    for dz in (-1, 0, 1):
        for dyy in (-1, 0, 1):
            for dxx in (-1, 0, 1):
                dd = (dxx, dyy, dz)
                if dd == (0, 0, 0):
                    continue
                nb = [pos[a] + dd[a] for a in range(3)]
                if not all(0 <= nb[a] < pg[a] for a in range(3)):
                    continue
                sl = []
                for a in range(3):
                    if dd[a] == 0:
                        sl.append(slice(0, dims[a]))
                    elif dd[a] < 0:
                        sl.append(slice(0, 2 * overlap))
                    else:
                        sl.append(slice(dims[a] - 2 * overlap, dims[a]))
                o.append(nb[2] * px * py + nb[1] * px + nb[0])
                mapping.append(loc[sl[2], sl[1], sl[0]].reshape(-1).astype(np.int32))
    # 7-point stencil, vectorised
    kk = loc
    cx, cy, cz = -1 / (hx * hx), -1 / (hy * hy), -1 / (hz * hz)
    diag = 2 / (hx * hx) + 2 / (hy * hy) + 2 / (hz * hz)
    R, C, V = [kk.reshape(-1)], [kk.reshape(-1)], [np.full(ndof, diag)]
    VN = [np.full(ndof, diag)]
    dN = np.zeros((t, h, w))
    # Neumann (natural) condition on artificial faces: drop the missing
    # neighbour from the diagonal as well
    if st[0] != 0:
        dN[:, :, 0] += cx
    if en[0] != Nx:
        dN[:, :, -1] += cx
    if st[1] != 0:
        dN[:, 0, :] += cy
    if en[1] != Ny:
        dN[:, -1, :] += cy
    if st[2] != 0:
        dN[0, :, :] += cz
    if en[2] != Nz:
        dN[-1, :, :] += cz
    VN[0] = VN[0] + dN.reshape(-1)
    for (a, b, c) in ((kk[:, :, 1:], kk[:, :, :-1], cx), (kk[:, 1:, :], kk[:, :-1, :], cy), (kk[1:, :, :], kk[:-1, :, :], cz)):
        a, b = a.reshape(-1), b.reshape(-1)
        R += [a, b]
        C += [b, a]
        V += [np.full(a.size, c), np.full(a.size, c)]
        VN += [np.full(a.size, c), np.full(a.size, c)]
    R, C = np.concatenate(R), np.concatenate(C)
    Mat = sp.csr_matrix((np.concatenate(V), (R, C)), shape=(ndof, ndof))
    Mat.sort_indices()
    MatN = None
    if neumann:
        MatN = sp.csr_matrix((np.concatenate(VN), (R, C)), shape=(ndof, ndof))
        MatN.sort_indices()
    if sym:
        Mat = sp.tril(Mat, format="csr")
        Mat.sort_indices()
    return dict(o=o, mapping=mapping, ndof=ndof, Mat=Mat, MatNeumann=MatN, d=d, f=f, sym=sym,
                box=tuple(zip(st, en)), grid=(px, py, pz), dims=dims)


def generate_world(size, dim=2, **kw):
    gen = generate2d if dim == 2 else generate3d
    return [gen(r, size, **kw) for r in range(size)]


# ----------------------------------------------------------------------------- 3-D linear elasticity (BASELINE config 4)
def _hex8_stiffness(h, E=1.0, nu=0.3):
    """24x24 stiffness of a trilinear (Q1) hexahedron of size hx x hy x hz, isotropic material,
    2x2x2 Gauss quadrature; dofs ordered node-major (node = i + 2j + 4k), component fastest."""
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    C = np.zeros((6, 6))
    C[:3, :3] = lam
    C[np.arange(3), np.arange(3)] += 2 * mu
    C[np.arange(3, 6), np.arange(3, 6)] = mu
    g = 1 / np.sqrt(3.0)
    K = np.zeros((24, 24))
    corners = np.array([[i, j, k] for k in (0, 1) for j in (0, 1) for i in (0, 1)], dtype=float) * 2 - 1
    for xi in (-g, g):
        for eta in (-g, g):
            for zeta in (-g, g):
                p = np.array([xi, eta, zeta])
                dN = np.zeros((8, 3))
                for a in range(8):
                    s = corners[a]
                    for c in range(3):
                        o = [q for q in range(3) if q != c]
                        dN[a, c] = 0.125 * s[c] * (1 + s[o[0]] * p[o[0]]) * (1 + s[o[1]] * p[o[1]]) * 2.0 / h[c]
                B = np.zeros((6, 24))
                for a in range(8):
                    B[0, 3 * a] = dN[a, 0]
                    B[1, 3 * a + 1] = dN[a, 1]
                    B[2, 3 * a + 2] = dN[a, 2]
                    B[3, 3 * a] = dN[a, 1]
                    B[3, 3 * a + 1] = dN[a, 0]
                    B[4, 3 * a + 1] = dN[a, 2]
                    B[4, 3 * a + 2] = dN[a, 1]
                    B[5, 3 * a] = dN[a, 2]
                    B[5, 3 * a + 2] = dN[a, 0]
                K += B.T @ C @ B * (h[0] * h[1] * h[2] / 8.0)
    return K


def _assemble_elasticity(lo, hi, Nn, h, Ke, penalty):
    """stiffness of the elements inside the node box [lo, hi) (local node numbering, x fastest,
    3 interleaved dofs per node); nodes on the global face x = 0 are clamped by penalisation
    (HPDDM's convention: diagonal = 1e30, picked up by Subdomain::boundaryConditions)."""
    w, hh, t = [hi[a] - lo[a] for a in range(3)]
    nid = np.arange(w * hh * t).reshape(t, hh, w)
    conn = np.stack([nid[k:t - 1 + k, j:hh - 1 + j, i:w - 1 + i].reshape(-1) for k in (0, 1) for j in (0, 1) for i in (0, 1)], axis=1)
    dofs = (3 * conn[:, :, None] + np.arange(3)[None, None, :]).reshape(conn.shape[0], 24)
    R = np.repeat(dofs, 24, axis=1).reshape(-1)
    C = np.tile(dofs, (1, 24)).reshape(-1)
    V = np.tile(Ke.reshape(-1), conn.shape[0])
    n = 3 * w * hh * t
    A = sp.csr_matrix((V, (R, C)), shape=(n, n))
    A.sort_indices()
    if penalty and lo[0] == 0:
        clamp = (3 * nid[:, :, 0].reshape(-1)[:, None] + np.arange(3)[None, :]).reshape(-1)
        rows = np.repeat(np.arange(n), np.diff(A.indptr))
        on_diag = np.nonzero(A.indices == rows)[0]          # every row has a stored diagonal entry (element stiffness)
        A.data[on_diag[clamp]] = penalty
    return A


def generate_elasticity3d(rank, size, Nn=(9, 9, 7), overlap=1, mu=4, grid=None, neumann=False, seed=4321, penalty=1e30, assembly="global"):
    """Q1 hexahedral linear elasticity (E = 1, nu_P = 0.3) on [0,10]^3 with Nn nodes per direction and 3
    dofs per node (config 4 of BASELINE.json, element model of the reference's examples/petsc/ex56.c);
    same decomposition conventions as generate3d, on nodes.  The local matrix is the global matrix
    restricted to the subdomain's dofs: assembly="global" assembles the whole grid and restricts (test sizes only);
    assembly="local" assembles only the elements that touch the subdomain (its node box grown by one node layer) and
    restricts -- identical up to the summation order of the element contributions, and scalable to config-4 sizes."""
    px, py, pz = grid if grid is not None else split_grid_3d(size)
    z = rank // (px * py)
    y = (rank - z * px * py) // px
    x = rank - z * px * py - y * px
    pos, pg = (x, y, z), (px, py, pz)
    st = [max(pos[a] * Nn[a] // pg[a] - overlap, 0) for a in range(3)]
    en = [min((pos[a] + 1) * Nn[a] // pg[a] + overlap, Nn[a]) for a in range(3)]
    dims = tuple(en[a] - st[a] for a in range(3))
    h = [10.0 / (Nn[a] - 1) for a in range(3)]
    Ke = _hex8_stiffness(h)
    if assembly == "local":
        lo = [max(st[a] - 1, 0) for a in range(3)]
        hi = [min(en[a] + 1, Nn[a]) for a in range(3)]
        Aext = _assemble_elasticity(lo, hi, Nn, h, Ke, penalty)
        eid = np.arange((hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2])).reshape(hi[2] - lo[2], hi[1] - lo[1], hi[0] - lo[0])
        nodes = eid[st[2] - lo[2]:en[2] - lo[2], st[1] - lo[1]:en[1] - lo[1], st[0] - lo[0]:en[0] - lo[0]].reshape(-1)
        dofs = (3 * nodes[:, None] + np.arange(3)[None, :]).reshape(-1)
        Mat = sp.csr_matrix(Aext[dofs][:, dofs])
    else:
        Aglob = _assemble_elasticity([0, 0, 0], list(Nn), Nn, h, Ke, penalty)
        gid = np.arange(Nn[0] * Nn[1] * Nn[2]).reshape(Nn[2], Nn[1], Nn[0])
        nodes = gid[st[2]:en[2], st[1]:en[1], st[0]:en[0]].reshape(-1)
        dofs = (3 * nodes[:, None] + np.arange(3)[None, :]).reshape(-1)
        Mat = sp.csr_matrix(Aglob[dofs][:, dofs])
    Mat.sort_indices()
    ndof = dofs.size
    dist = np.minimum(np.minimum(_ramp(st[0] != 0, en[0] != Nn[0], st[0], en[0], overlap)[None, None, :],
                                 _ramp(st[1] != 0, en[1] != Nn[1], st[1], en[1], overlap)[None, :, None]),
                      _ramp(st[2] != 0, en[2] != Nn[2], st[2], en[2], overlap)[:, None, None])
    d = np.repeat(np.minimum(dist / overlap, 1.0).reshape(-1), 3)
    loc = np.arange(dims[0] * dims[1] * dims[2]).reshape(dims[2], dims[1], dims[0])
    o, mapping = [], []
    for dz in (-1, 0, 1):
        for dyy in (-1, 0, 1):
            for dxx in (-1, 0, 1):
                dd = (dxx, dyy, dz)
                if dd == (0, 0, 0):
                    continue
                nb = [pos[a] + dd[a] for a in range(3)]
                if not all(0 <= nb[a] < pg[a] for a in range(3)):
                    continue
                sl = [slice(0, dims[a]) if dd[a] == 0 else (slice(0, 2 * overlap) if dd[a] < 0 else slice(dims[a] - 2 * overlap, dims[a])) for a in range(3)]
                ln = loc[sl[2], sl[1], sl[0]].reshape(-1)
                o.append(nb[2] * px * py + nb[1] * px + nb[0])
                mapping.append((3 * ln[:, None] + np.arange(3)[None, :]).reshape(-1).astype(np.int32))
    rs = np.random.RandomState(seed + rank)
    f = np.asfortranarray(rs.uniform(0.0, 1.0, size=(max(mu, 1), ndof)).T)
    MatN = _assemble_elasticity(st, en, Nn, h, Ke, penalty) if neumann else None
    return dict(o=o, mapping=mapping, ndof=ndof, Mat=Mat, MatNeumann=MatN, d=d, f=f, sym=False, box=tuple(zip(st, en)), grid=(px, py, pz),
                dims=(dims[0], dims[1], dims[2], 3))


def rigid_body_modes(part, Nn):
    """The 6 rigid-body modes (3 translations, 3 infinitesimal rotations about the subdomain's centre) of a
    generate_elasticity3d() subdomain, columns normalised: the kernel of its Neumann matrix and the natural first
    vectors of a GenEO-shaped coarse space for config 4 (a driver would pass them to setVectors)."""
    (x0, x1), (y0, y1), (z0, z1) = part["box"]
    h = [10.0 / (Nn[a] - 1) for a in range(3)]
    zz, yy, xx = np.meshgrid(np.arange(z0, z1) * h[2], np.arange(y0, y1) * h[1], np.arange(x0, x1) * h[0], indexing="ij")
    X = np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], axis=1)
    X = X - X.mean(axis=0)
    n = X.shape[0]
    Z = np.zeros((3 * n, 6), order="F")
    for c in range(3):
        Z[c::3, c] = 1.0                              # translations
    Z[0::3, 3], Z[1::3, 3] = -X[:, 1], X[:, 0]        # rotation about z
    Z[1::3, 4], Z[2::3, 4] = -X[:, 2], X[:, 1]        # rotation about x
    Z[2::3, 5], Z[0::3, 5] = -X[:, 0], X[:, 2]        # rotation about y
    return np.asfortranarray(Z / np.linalg.norm(Z, axis=0, keepdims=True))


# ----------------------------------------------------------------------------- 3-D Helmholtz, complex scalars (BASELINE config 5)
def generate_helmholtz3d(rank, size, N=(16, 16, 16), overlap=1, mu=1, grid=None, k=2.0, nu=4, seed=5678):
    """-Laplace(u) - k^2 u on [0,10]^3 with the first-order absorbing condition du/dn - i k u = 0 on the outer
    boundary, 7-point cell-centred differences, K = complex double (config 5 of BASELINE.json; the reference has no
    in-tree Helmholtz generator -- SURVEY.md section 8d -- this is the driver-side model of one).  Same decomposition
    conventions as generate3d.  Returns, besides the generate3d keys:
      Mat       local matrix = restriction of the global operator to the subdomain (what Subdomain::initialize and
                GMV use); complex symmetric, not Hermitian
      MatRobin  the matrix ORAS factorises (Schwarz::callNumfact(A) -> Prcndtnr::OG, schwarz.hpp:351-365): the same
                operator with the impedance condition du/dn - i k u = 0 also on the artificial (interior) faces
      Z         nu plane waves exp(i k dir.x) restricted to the subdomain, column-normalised: the coarse vectors a
                driver would pass to setVectors (the reference expects user-supplied vectors / an (A, B) pencil here)
    """
    base = generate3d(rank, size, N=N, overlap=overlap, mu=mu, grid=grid, seed=seed)
    Nx, Ny, Nz = N
    (x0, x1), (y0, y1), (z0, z1) = base["box"]
    w, h, t = base["dims"]
    ndof = base["ndof"]
    hs = (10.0 / Nx, 10.0 / Ny, 10.0 / Nz)
    lo, hi, NN = (x0, y0, z0), (x1, y1, z1), (Nx, Ny, Nz)
    # ghost-cell elimination of  (u_g - u_b) / h = i k u_b  ->  the missing neighbour contributes c * (1 + i k h) to the diagonal
    ext = np.zeros((t, h, w), dtype=np.complex128)   # outer boundary faces
    art = np.zeros((t, h, w), dtype=np.complex128)   # artificial faces (ORAS transmission condition)
    for a in range(3):
        c = -1.0 / (hs[a] * hs[a])
        val = c * (1.0 + 1j * k * hs[a])
        first = [slice(None)] * 3
        last = [slice(None)] * 3
        first[2 - a] = 0
        last[2 - a] = -1
        (ext if lo[a] == 0 else art)[tuple(first)] += val
        (ext if hi[a] == NN[a] else art)[tuple(last)] += val
    A0 = sp.csr_matrix(base["Mat"]).astype(np.complex128)
    Mat = sp.csr_matrix(A0 + sp.diags(ext.reshape(-1) - k * k))
    Mat.sort_indices()
    MatRobin = sp.csr_matrix(Mat + sp.diags(art.reshape(-1)))
    MatRobin.sort_indices()
    rs = np.random.RandomState(seed + 17 * rank)
    f = np.asfortranarray((rs.uniform(0.0, 1.0, size=(max(mu, 1), ndof)) + 1j * rs.uniform(-0.5, 0.5, size=(max(mu, 1), ndof))).T)
    xs = (np.arange(x0, x1) + 0.5) * hs[0]
    ys = (np.arange(y0, y1) + 0.5) * hs[1]
    zs = (np.arange(z0, z1) + 0.5) * hs[2]
    dirs = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (-1, 0, 0), (0, -1, 0), (0, 0, -1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]
    Z = np.empty((ndof, nu), dtype=np.complex128, order="F")
    for c in range(nu):
        dv = np.array(dirs[c % len(dirs)], dtype=float)
        dv /= np.linalg.norm(dv)
        kk = k * (1 + c // len(dirs))
        v = np.exp(1j * kk * (dv[2] * zs[:, None, None] + dv[1] * ys[None, :, None] + dv[0] * xs[None, None, :])).reshape(-1)
        Z[:, c] = v / np.linalg.norm(v)
    return dict(base, Mat=Mat, MatRobin=MatRobin, MatNeumann=None, f=f, Z=Z, k=k, sym=False)
