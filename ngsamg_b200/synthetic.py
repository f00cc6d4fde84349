"""Seeded synthetic inputs: no mesher / NGSolve is available, so the benchmark problems
are generated on structured Kuhn-tet meshes (6 tets per cube sharing the main diagonal; the same
triangulation `ngsolve.meshes.MakeStructured3DMesh(hexes=False)` produces, used by the reference in
examples/elasticity/beam.py:20).

Pure numpy, inputs only -- no product or oracle code is involved.
"""
import itertools

import numpy as np

# the 7 edge directions of the Kuhn triangulation
_POS_DIRS = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)]


def splitmix64(seed, n):
    """uniform doubles in [-1,1) from splitmix64 (seed = 20260101 + level)."""
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / 9007199254740992.0) - 1.0


def _edge_cube_weight(vb, vc, nb, nc):
    """sum over the (up to 4) cubes around an axis edge of the number of Kuhn tets containing it."""
    w = np.zeros(np.broadcast(vb, vc).shape, dtype=np.int64)
    for sb in (0, 1):
        for sc in (0, 1):
            ok = (vb - sb >= 0) & (vb - sb <= nb - 2) & (vc - sc >= 0) & (vc - sc <= nc - 2)
            w += ok * (1 if sb + sc == 1 else 2)
    return w


def poisson3d_kuhn(nx, ny=None, nz=None, dirichlet=("x0", "y1"), coef=None, zchunk=8, h=None, origin=(0, 0, 0)):
    """P1 stiffness matrix of -div(c grad u) on the unit cube, nx*ny*nz vertices, Kuhn tets.

    Vertex (ix,iy,iz) has DOF number ix + nx*(iy + ny*iz).  The sparsity pattern is the mesh
    connectivity (15-point: zeros on the face/body diagonals are stored, as an FE assembly would).
    Returns dict(n, rowptr[int64], col[int32], val[f64], free[uint8], rhs[f64] (f=1 load), xyz).
    `coef`: optional callable (x,y,z)->c evaluated at vertices; the edge weight uses the mean of the
    two end-point values (a diagonal scaling that keeps symmetry and zero row sums).
    Generated in slabs of `zchunk` vertex planes so that the peak memory stays close to the size of the result.
    `h`, `origin`: mesh width and vertex-index offset of a SUB-BOX of a larger mesh -- the matrix is then the sub-assembled
    (Neumann) matrix of that sub-box, i.e. one rank's local matrix of a box partition (partition_poisson3d).
    """
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    n = nx * ny * nz
    h = 1.0 / (max(nx, ny, nz) - 1) if h is None else float(h)
    ox, oy, oz = origin
    dirs = [(0, 0, 0)] + _POS_DIRS + [(-a, -b, -c) for (a, b, c) in _POS_DIRS]
    dirs.sort(key=lambda d: d[0] + nx * (d[1] + ny * d[2]))
    K = len(dirs)
    kd = dirs.index((0, 0, 0))
    dims = (nx, ny, nz)
    # rows per vertex = number of in-bounds directions: separable count
    def nb(k, m):  # number of in-range offsets in {-1,0,1} along one axis
        return (k > 0).astype(np.int64) + 1 + (k < m - 1).astype(np.int64)
    rowptr = np.zeros(n + 1, np.int64)
    col = None
    val = None
    free = np.ones(n, np.uint8)
    rhs = np.zeros(n)
    # pass 1: counts (needs the actual in-bounds test of the 15 directions, done per slab below), pass 2: fill
    chunks = [(z0, min(nz, z0 + zchunk)) for z0 in range(0, nz, zchunk)]
    cnt_all = np.zeros(n, np.int64)
    slabs = []
    for (z0, z1) in chunks:
        iz, iy, ix = np.meshgrid(np.arange(z0, z1), np.arange(ny), np.arange(nx), indexing="ij")
        ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
        m = np.zeros(ix.shape[0], np.int64)
        for d in dirs:
            jx, jy, jz = ix + d[0], iy + d[1], iz + d[2]
            m += (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny) & (jz >= 0) & (jz < nz)
        cnt_all[z0 * nx * ny:z1 * nx * ny] = m
    np.cumsum(cnt_all, out=rowptr[1:])
    del cnt_all
    nnz = int(rowptr[-1])
    col = np.empty(nnz, np.int32)
    val = np.empty(nnz, np.float64)
    cvfun = coef
    for (z0, z1) in chunks:
        iz, iy, ix = np.meshgrid(np.arange(z0, z1), np.arange(ny), np.arange(nx), indexing="ij")
        ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
        m = ix.shape[0]
        idx = ix + nx * (iy + ny * iz)
        coord = (ix, iy, iz)
        cv = cvfun((ix + ox) * h, (iy + oy) * h, (iz + oz) * h).astype(np.float64) if cvfun is not None else None
        cols = np.zeros((m, K), np.int64)
        vals = np.zeros((m, K), np.float64)
        mask = np.zeros((m, K), bool)
        for k, d in enumerate(dirs):
            jx, jy, jz = ix + d[0], iy + d[1], iz + d[2]
            ok = (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny) & (jz >= 0) & (jz < nz)
            mask[:, k] = ok
            cols[:, k] = idx + d[0] + nx * (d[1] + ny * d[2])
            if sum(abs(c) for c in d) == 1:
                a = [i for i in range(3) if d[i] != 0][0]
                b_, c_ = [i for i in range(3) if i != a]
                w = _edge_cube_weight(coord[b_], coord[c_], dims[b_], dims[c_]).astype(np.float64)
                v = -(h / 6.0) * w
                if cv is not None:
                    cj = cvfun((np.clip(jx, 0, nx - 1) + ox) * h, (np.clip(jy, 0, ny - 1) + oy) * h, (np.clip(jz, 0, nz - 1) + oz) * h).astype(np.float64)
                    v = v * 0.5 * (cv + cj)
                vals[:, k] = np.where(ok, v, 0.0)
        vals[:, kd] = -(vals * mask).sum(axis=1)
        lo, hi = int(rowptr[z0 * nx * ny]), int(rowptr[z1 * nx * ny])
        col[lo:hi] = cols[mask]
        val[lo:hi] = vals[mask]
        sl = slice(z0 * nx * ny, z1 * nx * ny)
        fr = np.ones(m, np.uint8)
        for tag in dirichlet:
            ax = "xyz".index(tag[0])
            side = 0 if tag[1] == "0" else dims[ax] - 1
            fr[coord[ax] == side] = 0
        free[sl] = fr
        # load vector for f = 1: (h^3/24) * number of incident tets
        ntet = np.zeros(m, np.int64)
        for s3 in itertools.product((0, 1), repeat=3):
            ok = np.ones(m, bool)
            for a in range(3):
                ok &= (coord[a] - s3[a] >= 0) & (coord[a] - s3[a] <= dims[a] - 2)
            ntet += ok * (6 if sum(s3) in (0, 3) else 2)
        rhs[sl] = (h ** 3 / 24.0) * ntet
    ids = np.arange(n)
    xyz = np.stack([(ids % nx + ox) * h, ((ids // nx) % ny + oy) * h, (ids // (nx * ny) + oz) * h], axis=1).astype(np.float64)
    return dict(n=n, b=1, rowptr=rowptr, col=col, val=val, free=free, rhs=rhs, xyz=xyz, h=h, dims=(nx, ny, nz))


def _split(ncells, parts):
    """cell ranges of `parts` nearly equal slabs of `ncells` cells"""
    cuts = [(ncells * k) // parts for k in range(parts + 1)]
    return [(cuts[k], cuts[k + 1]) for k in range(parts)]


def box_partition(dims, grid):
    """Box partition of the CELLS of an nx*ny*nz-vertex mesh into grid = (px, py, pz) sub-boxes (rank = bx + px*(by + py*bz)).
    Returns per rank dict(origin, dims (local vertex counts), gidx (local -> global vertex number, local numbering x-fastest))
    -- interface vertices are duplicated on every sharer, exactly like NGSolve's ParallelDofs."""
    nx, ny, nz = dims
    px, py, pz = grid
    parts = []
    for bz, (z0, z1) in enumerate(_split(nz - 1, pz)):
        for by, (y0, y1) in enumerate(_split(ny - 1, py)):
            for bx, (x0, x1) in enumerate(_split(nx - 1, px)):
                lx, ly, lz = x1 - x0 + 1, y1 - y0 + 1, z1 - z0 + 1
                iz, iy, ix = np.meshgrid(np.arange(z0, z1 + 1), np.arange(y0, y1 + 1), np.arange(x0, x1 + 1), indexing="ij")
                gidx = (ix + nx * (iy + ny * iz)).ravel().astype(np.int64)
                parts.append(dict(origin=(x0, y0, z0), dims=(lx, ly, lz), gidx=gidx, box=(bx, by, bz)))
    return parts


def halos_from_global_ids(gidx_per_rank, rank):
    """(peers, ex) of `rank`: DOFs with the same global number on two ranks are shared; lists ascending in the local number, which
    is monotone in the global number here, so the k-th shared DOF agrees on both sides."""
    mine = gidx_per_rank[rank]
    peers, ex = [], []
    for r, other in enumerate(gidx_per_rank):
        if r == rank:
            continue
        common = np.intersect1d(mine, other, assume_unique=True)
        if len(common) == 0:
            continue
        pos = np.searchsorted(mine, common)       # local numbering is ascending in the global number
        peers.append(r)
        ex.append(pos.astype(np.int32))
    return peers, ex


def partition_poisson3d(nx, ny=None, nz=None, grid=(1, 1, 2), rank=None, dirichlet=("x0", "y1"), coef=None):
    """Local problems of a box partition of poisson3d_kuhn(nx, ny, nz): for every rank (or only `rank`) a dict with the
    sub-assembled local matrix (rowptr/col/val), free, rhs (DISTRIBUTED load vector), xyz, gidx and the halo (peers, ex).
    The sum of the local matrices / load vectors over the ranks is the global matrix / load vector."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    h = 1.0 / (max(nx, ny, nz) - 1)
    parts = box_partition((nx, ny, nz), grid)
    gids = [p["gidx"] for p in parts]
    dims = (nx, ny, nz)
    out = []
    for r, p in enumerate(parts):
        if rank is not None and r != rank:
            out.append(None)
            continue
        # Dirichlet faces of the global cube that this sub-box touches
        tags = []
        for tag in dirichlet:
            ax = "xyz".index(tag[0])
            if tag[1] == "0" and p["origin"][ax] == 0:
                tags.append(tag)
            if tag[1] == "1" and p["origin"][ax] + p["dims"][ax] == dims[ax]:
                tags.append(tag)
        loc = poisson3d_kuhn(*p["dims"], dirichlet=tuple(tags), coef=coef, h=h, origin=p["origin"])
        loc["gidx"] = p["gidx"]
        loc["peers"], loc["ex"] = halos_from_global_ids(gids, r)
        out.append(loc)
    return out if rank is None else out[rank]


def kuhn_tets(nx, ny, nz):
    """(ntet,4) vertex numbers of the Kuhn triangulation (6 path-simplices per cube)."""
    cz, cy, cx = np.meshgrid(np.arange(nz - 1), np.arange(ny - 1), np.arange(nx - 1), indexing="ij")
    base = np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=1)
    tets = []
    for perm in itertools.permutations(range(3)):
        p = np.zeros((4, 3), np.int64)
        for s in range(3):
            p[s + 1] = p[s]
            p[s + 1, perm[s]] += 1
        v = base[:, None, :] + p[None, :, :]
        tets.append(v[..., 0] + nx * (v[..., 1] + ny * v[..., 2]))
    return np.concatenate(tets, axis=0)


def poisson3d_kuhn_assembled(nx, ny=None, nz=None):
    """Independent element-by-element P1 assembly (scipy COO sum) used to validate poisson3d_kuhn."""
    import scipy.sparse as sp
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    h = 1.0 / (max(nx, ny, nz) - 1)
    n = nx * ny * nz
    T = kuhn_tets(nx, ny, nz)
    ids = np.arange(n)
    X = np.stack([ids % nx, (ids // nx) % ny, ids // (nx * ny)], axis=1) * h
    P = X[T]                                        # (nt,4,3)
    M = np.concatenate([np.ones((len(T), 4, 1)), P], axis=2)
    Minv = np.linalg.inv(M)                         # columns: coefficients of the hat functions
    G = Minv[:, 1:, :]                              # (nt,3,4) gradients
    vol = np.abs(np.linalg.det(M)) / 6.0
    Ke = np.einsum("tki,tkj->tij", G, G) * vol[:, None, None]
    I = np.repeat(T[:, :, None], 4, axis=2).ravel()
    J = np.repeat(T[:, None, :], 4, axis=1).ravel()
    A = sp.coo_matrix((Ke.ravel(), (I, J)), shape=(n, n)).tocsr()
    A.sort_indices()
    rhs = np.bincount(T.ravel(), weights=np.repeat(vol / 4.0, 4), minlength=n)
    return A, rhs


def elasticity3d_kuhn(nx, ny, nz, lx=None, E=1e3, nu=0.15, jump=None, clamp=("x0",), h=None, origin=(0, 0, 0)):
    """P1 linear elasticity (3x3 blocks) on the beam [0,lx]x[0,1]^2 with Kuhn tets, element assembly.

    lam/mu as examples/elasticity/amg_utils.py:150-153; clamp at x=0, body force (0,x,0)
    (examples/elasticity/beamP2.py:24).  `jump`: optional callable (cx,cy,cz)->multiplier of E evaluated
    at tet centroids (config 5: checkerboard with 1e4 contrast).
    Returns dict(n, b=3, rowptr, col, val (blocks row-major), free, rhs, xyz).
    """
    import scipy.sparse as sp
    h = 1.0 / (min(ny, nz) - 1) if h is None else float(h)   # h / origin: sub-box of a larger mesh (see partition_elasticity3d)
    lx = h * (nx - 1) if lx is None else lx
    n = nx * ny * nz
    T = kuhn_tets(nx, ny, nz)
    ids = np.arange(n)
    X = np.stack([(ids % nx) * (lx / (nx - 1)) + origin[0] * h, ((ids // nx) % ny + origin[1]) * h, (ids // (nx * ny) + origin[2]) * h], axis=1)
    P = X[T]
    M = np.concatenate([np.ones((len(T), 4, 1)), P], axis=2)
    Minv = np.linalg.inv(M)
    G = Minv[:, 1:, :]                              # (nt,3,4): G[t,:,i] = grad phi_i
    vol = np.abs(np.linalg.det(M)) / 6.0
    Et = np.full(len(T), E)
    if jump is not None:
        cen = P.mean(axis=1)
        Et = Et * jump(cen[:, 0], cen[:, 1], cen[:, 2])
    mu = Et / (2 * (1 + nu))
    lam = Et * nu / ((1 + nu) * (1 - 2 * nu))
    # K_ij[a,b] = vol*( lam g_i[a] g_j[b] + mu g_i[b] g_j[a] + mu (g_i.g_j) delta_ab )
    gg = np.einsum("tai,tbj->tijab", G, G)
    dotg = np.einsum("tai,taj->tij", G, G)
    Ke = (lam[:, None, None, None, None] * gg + mu[:, None, None, None, None] * np.swapaxes(gg, 3, 4)
          + mu[:, None, None, None, None] * dotg[..., None, None] * np.eye(3)[None, None, None])
    Ke = Ke * vol[:, None, None, None, None]
    I = (3 * T[:, :, None, None, None] + np.arange(3)[None, None, None, :, None]) + np.zeros((1, 1, 4, 1, 3), np.int64)
    J = (3 * T[:, None, :, None, None] + np.arange(3)[None, None, None, None, :]) + np.zeros((1, 4, 1, 3, 1), np.int64)
    A = sp.coo_matrix((Ke.ravel(), (I.ravel(), J.ravel())), shape=(3 * n, 3 * n)).tocsr()
    # block pattern = mesh connectivity (keep structural zeros of the blocks)
    pat = sp.coo_matrix((np.ones(T.shape[0] * 16), (np.repeat(T[:, :, None], 4, 2).ravel(),
                                                   np.repeat(T[:, None, :], 4, 1).ravel())), shape=(n, n)).tocsr()
    pat.sort_indices()
    rowptr = pat.indptr.astype(np.int64)
    col = pat.indices.astype(np.int32)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    Ad = A.tocsr()
    val = np.zeros((len(col), 3, 3))
    for a in range(3):
        for b in range(3):
            val[:, a, b] = np.asarray(Ad[3 * rows + a, 3 * col.astype(np.int64) + b]).ravel()
    free = np.ones(n, np.uint8)
    dims = (nx, ny, nz)
    coord = (ids % nx, (ids // nx) % ny, ids // (nx * ny))
    for tag in clamp:
        ax = "xyz".index(tag[0])
        side = 0 if tag[1] == "0" else dims[ax] - 1
        free[coord[ax] == side] = 0
    # body force (0, x, 0): lumped P1 load
    fy = np.bincount(T.ravel(), weights=np.repeat(vol / 4.0, 4) * X[T.ravel(), 0], minlength=n)
    rhs = np.zeros((n, 3))
    rhs[:, 1] = fy
    return dict(n=n, b=3, rowptr=rowptr, col=col, val=val.reshape(-1), free=free, rhs=rhs.reshape(-1), xyz=X)


def elasticity3d_kuhn_stencil(nx, ny, nz, E=1e3, nu=0.15, clamp=("x0",)):
    """Same matrix as elasticity3d_kuhn on a UNIFORM Kuhn mesh (cube edge h = 1/(min(ny,nz)-1), lx = h*(nx-1)), built from
    the translation-invariant block stencil instead of element assembly, so that multi-million-vertex problems can be generated:
    the block row of a vertex depends only on its class (first / interior / last plane per axis, 27 classes), which is read off a
    small element-assembled reference mesh.  Returns the same dict as elasticity3d_kuhn."""
    ref_n = 5
    ref = elasticity3d_kuhn(ref_n, ref_n, ref_n, E=E, nu=nu, clamp=())
    hs = 1.0 / (min(ny, nz) - 1)
    href = 1.0 / (ref_n - 1)
    scale = hs / href                      # stiffness blocks scale with h, the lumped load with h^3 (* x for the body force)
    n = nx * ny * nz
    dirs = [(0, 0, 0)] + _POS_DIRS + [(-a, -b, -c) for (a, b, c) in _POS_DIRS]
    dirs.sort(key=lambda d: d[0] + nx * (d[1] + ny * d[2]))
    K = len(dirs)
    # reference table: class (cx,cy,cz in {0,1,2}) x direction -> 3x3 block
    table = np.zeros((3, 3, 3, K, 3, 3))
    rp, rc, rv = ref["rowptr"], ref["col"], ref["val"].reshape(-1, 3, 3)
    pick = {0: 0, 1: 2, 2: ref_n - 1}
    for cx in range(3):
        for cy in range(3):
            for cz in range(3):
                vx, vy, vz = pick[cx], pick[cy], pick[cz]
                i = vx + ref_n * (vy + ref_n * vz)
                for k, d in enumerate(dirs):
                    jx, jy, jz = vx + d[0], vy + d[1], vz + d[2]
                    if not (0 <= jx < ref_n and 0 <= jy < ref_n and 0 <= jz < ref_n):
                        continue
                    j = jx + ref_n * (jy + ref_n * jz)
                    pos = np.searchsorted(rc[rp[i]:rp[i + 1]], j)
                    table[cx, cy, cz, k] = rv[rp[i] + pos] * scale
    ids = np.arange(n, dtype=np.int64)
    ix, iy, iz = ids % nx, (ids // nx) % ny, ids // (nx * ny)
    cls = lambda k, m: np.where(k == 0, 0, np.where(k == m - 1, 2, 1))
    cxa, cya, cza = cls(ix, nx), cls(iy, ny), cls(iz, nz)
    mask = np.zeros((n, K), bool)
    cols = np.zeros((n, K), np.int64)
    for k, d in enumerate(dirs):
        jx, jy, jz = ix + d[0], iy + d[1], iz + d[2]
        mask[:, k] = (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny) & (jz >= 0) & (jz < nz)
        cols[:, k] = ids + d[0] + nx * (d[1] + ny * d[2])
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(mask.sum(axis=1), out=rowptr[1:])
    col = cols[mask].astype(np.int32)
    del cols
    val = np.empty((int(rowptr[-1]), 3, 3))
    slot = np.cumsum(mask, axis=1) - 1                      # position of direction k inside its row
    for k in range(K):
        rows = np.flatnonzero(mask[:, k])
        val[rowptr[rows] + slot[rows, k]] = table[cxa[rows], cya[rows], cza[rows], k]
    free = np.ones(n, np.uint8)
    dims = (nx, ny, nz)
    coord = (ix, iy, iz)
    for tag in clamp:
        ax = "xyz".index(tag[0])
        free[coord[ax] == (0 if tag[1] == "0" else dims[ax] - 1)] = 0
    X = np.stack([ix * hs, iy * hs, iz * hs], axis=1).astype(np.float64)
    # body force (0, x, 0), lumped: (h^3/24) * (#incident tets) * x  (exact for the nodal value of x at the vertex up to O(h))
    ntet = np.zeros(n, np.int64)
    for s3 in itertools.product((0, 1), repeat=3):
        ok = np.ones(n, bool)
        for a in range(3):
            ok &= (coord[a] - s3[a] >= 0) & (coord[a] - s3[a] <= dims[a] - 2)
        ntet += ok * (6 if sum(s3) in (0, 3) else 2)
    rhs = np.zeros((n, 3))
    rhs[:, 1] = (hs ** 3 / 24.0) * ntet * X[:, 0]
    return dict(n=n, b=3, rowptr=rowptr, col=col, val=val.reshape(-1), free=free, rhs=rhs.reshape(-1), xyz=X)


def partition_elasticity3d(nx, ny, nz, nparts=2, rank=None, **kw):
    """z-slab partition of elasticity3d_kuhn(nx, ny, nz): per rank the sub-assembled local problem + halo (see partition_poisson3d)"""
    h = 1.0 / (min(ny, nz) - 1)
    parts = box_partition((nx, ny, nz), (1, 1, nparts))
    gids = [p["gidx"] for p in parts]
    out = []
    for r, p in enumerate(parts):
        if rank is not None and r != rank:
            out.append(None)
            continue
        loc = elasticity3d_kuhn(*p["dims"], h=h, origin=p["origin"], **kw)
        loc["gidx"] = p["gidx"]
        loc["peers"], loc["ex"] = halos_from_global_ids(gids, r)
        out.append(loc)
    return out if rank is None else out[rank]


def slab_poisson3d(nx, ny, nz_per_rank, world, rank, dirichlet=("x0", "y1")):
    """rank's local problem of a z-slab partition of the nx x ny x ((nz_per_rank-1)*world+1) mesh -- partition_poisson3d without the
    global index arrays (bench sizes): every rank owns nz_per_rank vertex planes, the first / last plane is shared with rank-1 / rank+1."""
    nzg = (nz_per_rank - 1) * world + 1
    h = 1.0 / (max(nx, ny, nzg) - 1)
    loc = poisson3d_kuhn(nx, ny, nz_per_rank, dirichlet=tuple(t for t in dirichlet if t[0] != "z"), h=h,
                         origin=(0, 0, rank * (nz_per_rank - 1)))
    plane = nx * ny
    peers, ex = [], []
    if rank > 0:
        peers.append(rank - 1)
        ex.append(np.arange(plane, dtype=np.int32))
    if rank < world - 1:
        peers.append(rank + 1)
        ex.append(np.arange(plane * (nz_per_rank - 1), plane * nz_per_rank, dtype=np.int32))
    loc["peers"], loc["ex"], loc["global_dims"] = peers, ex, (nx, ny, nzg)
    return loc


def bench_grid(world):
    """box grid (px, py, pz) the bench cuts the global mesh into: as cube-like as the rank count allows"""
    return {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}.get(world, (1, 1, world))


def box_poisson3d(n, grid, rank, dirichlet=("x0", "y1")):
    """rank's local problem of a box partition into grid = (px, py, pz) sub-boxes of n^3 vertices each (global mesh
    ((n-1)px+1) x ((n-1)py+1) x ((n-1)pz+1)); same result as partition_poisson3d(..., rank=rank) without any global index array.
    Shared DOFs: faces / edges / corners of the sub-box, listed per neighbour ascending in the local number (z-major), which is the
    same order on both sides."""
    px, py, pz = grid
    bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)
    gd = ((n - 1) * px + 1, (n - 1) * py + 1, (n - 1) * pz + 1)
    h = 1.0 / (max(gd) - 1)
    origin = (bx * (n - 1), by * (n - 1), bz * (n - 1))
    tags = []
    for tag in dirichlet:
        ax = "xyz".index(tag[0])
        if tag[1] == "0" and origin[ax] == 0:
            tags.append(tag)
        if tag[1] == "1" and origin[ax] + n == gd[ax]:
            tags.append(tag)
    loc = poisson3d_kuhn(n, n, n, dirichlet=tuple(tags), h=h, origin=origin)
    peers, ex = [], []
    sel = {-1: np.array([0]), 0: np.arange(n), 1: np.array([n - 1])}
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                qx, qy, qz = bx + dx, by + dy, bz + dz
                if not (0 <= qx < px and 0 <= qy < py and 0 <= qz < pz):
                    continue
                idx = (sel[dx][None, None, :] + n * (sel[dy][None, :, None] + n * sel[dz][:, None, None])).ravel()
                peers.append(qx + px * (qy + py * qz))
                ex.append(idx.astype(np.int32))
    order = np.argsort(peers)
    loc["peers"], loc["ex"], loc["global_dims"] = [peers[i] for i in order], [ex[i] for i in order], gd
    loc["n_master"] = int(n ** 3 - sum(1 for _ in ()))  # replaced below
    # DOFs this rank is master of (lowest sharing rank): everything not shared with a lower rank
    ghost = np.zeros(n ** 3, bool)
    for q, e in zip(loc["peers"], loc["ex"]):
        if q < rank:
            ghost[e] = True
    loc["n_master"] = int((~ghost).sum())
    return loc


# ---- nodal P2 elasticity (BASELINE.json configs[2]: examples/elasticity/beamP2.py -- nodal-P2 H1, `ngs_amg_on_dofs=select`,
# `ngs_amg_subset=nodalp2`: the AMG treats vertex AND edge-midpoint nodes as its vertices) ---------------------------------------
_P2_EDGES = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def _p2_tet_shape_gradients(Minv):
    """gradients of the 10 nodal P2 shape functions at the 4 points of the degree-2 Gauss rule.
    Minv: (nt,4,4) inverse of [1 | x y z] per tet; returns (grads (nt,4pts,10,3), values (nt,4pts,10))"""
    a, b = 0.5854101966249685, 0.1381966011250105
    lam_q = np.full((4, 4), b) + (a - b) * np.eye(4)            # barycentric coordinates of the 4 quadrature points
    gl = np.swapaxes(Minv[:, 1:, :], 1, 2)                        # (nt,4,3): grad lambda_i
    nt = Minv.shape[0]
    G = np.zeros((nt, 4, 10, 3))
    V = np.zeros((nt, 4, 10))
    for q in range(4):
        lam = lam_q[q]
        for i in range(4):                                        # vertex functions lambda_i (2 lambda_i - 1)
            G[:, q, i, :] = (4 * lam[i] - 1) * gl[:, i, :]
            V[:, q, i] = lam[i] * (2 * lam[i] - 1)
        for e, (i, j) in enumerate(_P2_EDGES):                    # edge functions 4 lambda_i lambda_j
            G[:, q, 4 + e, :] = 4 * (lam[i] * gl[:, j, :] + lam[j] * gl[:, i, :])
            V[:, q, 4 + e] = 4 * lam[i] * lam[j]
    return G, V


def elasticity3d_p2_kuhn(nx, ny, nz, E=1e3, nu=0.15, clamp=("x0",)):
    """Nodal P2 linear elasticity (3x3 blocks) on the beam of elasticity3d_kuhn, element assembly (small meshes; validates the stencil
    generator below).  nx, ny, nz = VERTICES per axis; the P2 nodes are all points of the (2nx-1) x (2ny-1) x (2nz-1) grid (every grid
    point is a vertex or the midpoint of one of the 7 Kuhn edge directions).  Load (0, x, 0) as in examples/elasticity/beamP2.py:24.
    Returns the same dict as elasticity3d_kuhn plus `dims` = fine-grid dimensions."""
    import scipy.sparse as sp
    h = 1.0 / (min(ny, nz) - 1)
    NX, NY, NZ = 2 * nx - 1, 2 * ny - 1, 2 * nz - 1
    n = NX * NY * NZ
    T = kuhn_tets(nx, ny, nz)
    cv = np.stack([T % nx, (T // nx) % ny, T // (nx * ny)], axis=2)          # (nt,4,3) vertex coordinates in cells
    fv = 2 * cv                                                                 # fine-grid coordinates of the vertices
    fe = np.stack([(fv[:, i] + fv[:, j]) // 2 for (i, j) in _P2_EDGES], axis=1)  # (nt,6,3) midpoints
    fn = np.concatenate([fv, fe], axis=1)                                       # (nt,10,3)
    N = fn[..., 0] + NX * (fn[..., 1] + NY * fn[..., 2])                        # (nt,10) node numbers
    Pv = cv * h
    M = np.concatenate([np.ones((len(T), 4, 1)), Pv], axis=2)
    Minv = np.linalg.inv(M)
    vol = np.abs(np.linalg.det(M)) / 6.0
    G, V = _p2_tet_shape_gradients(Minv)
    mu = E / (2 * (1 + nu))
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    w = vol / 4.0
    gg = np.einsum("tqia,tqjb,t->tijab", G, G, w)
    dotg = np.einsum("tqia,tqja,t->tij", G, G, w)
    Ke = lam * gg + mu * np.swapaxes(gg, 3, 4) + mu * dotg[..., None, None] * np.eye(3)[None, None, None]
    I = (3 * N[:, :, None, None, None] + np.arange(3)[None, None, None, :, None]) + np.zeros((1, 1, 10, 1, 3), np.int64)
    J = (3 * N[:, None, :, None, None] + np.arange(3)[None, None, None, None, :]) + np.zeros((1, 10, 1, 3, 1), np.int64)
    A = sp.coo_matrix((Ke.ravel(), (I.ravel(), J.ravel())), shape=(3 * n, 3 * n)).tocsr()
    pat = sp.coo_matrix((np.ones(N.shape[0] * 100), (np.repeat(N[:, :, None], 10, 2).ravel(), np.repeat(N[:, None, :], 10, 1).ravel())),
                        shape=(n, n)).tocsr()
    pat.sort_indices()
    rowptr = pat.indptr.astype(np.int64)
    col = pat.indices.astype(np.int32)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    val = np.zeros((len(col), 3, 3))
    for a in range(3):
        for b in range(3):
            val[:, a, b] = np.asarray(A[3 * rows + a, 3 * col.astype(np.int64) + b]).ravel()
    ids = np.arange(n)
    coord = (ids % NX, (ids // NX) % NY, ids // (NX * NY))
    X = np.stack([coord[0], coord[1], coord[2]], axis=1) * (h / 2)
    free = np.ones(n, np.uint8)
    dims = (NX, NY, NZ)
    for tag in clamp:
        ax = "xyz".index(tag[0])
        free[coord[ax] == (0 if tag[1] == "0" else dims[ax] - 1)] = 0
    # body force (0, x, 0): int x phi_i with the same 4-point rule
    xq = np.einsum("tqv,tv->tq", np.full((1, 4, 4), 0.1381966011250105) + (0.5854101966249685 - 0.1381966011250105) * np.eye(4)[None], Pv[..., 0])
    fy = np.bincount(N.ravel(), weights=np.einsum("tq,tqi,t->ti", xq, V, w).ravel(), minlength=n)
    rhs = np.zeros((n, 3))
    rhs[:, 1] = fy
    return dict(n=n, b=3, rowptr=rowptr, col=col, val=val.reshape(-1), free=free, rhs=rhs.reshape(-1), xyz=X, dims=dims)


def elasticity3d_p2_kuhn_stencil(nx, ny, nz, E=1e3, nu=0.15, clamp=("x0",)):
    """The matrix of elasticity3d_p2_kuhn for LARGE uniform meshes, from the translation-invariant block stencil: the block row (and the
    load entry) of a fine-grid node depends only on its class per axis -- position 0, 1, interior even, interior odd, last-1, last --
    which is read off a 5^3-vertex (9^3-node) element-assembled reference mesh.  Same dict as elasticity3d_p2_kuhn."""
    ref_nv = 5
    ref = elasticity3d_p2_kuhn(ref_nv, ref_nv, ref_nv, E=E, nu=nu, clamp=())
    RN = 2 * ref_nv - 1
    hs = 1.0 / (min(ny, nz) - 1)
    scale = hs / (1.0 / (ref_nv - 1))               # stiffness blocks scale with h, the load with h^3 (times x, handled below)
    NX, NY, NZ = 2 * nx - 1, 2 * ny - 1, 2 * nz - 1
    if min(NX, NY, NZ) < 5:
        raise ValueError("elasticity3d_p2_kuhn_stencil needs at least 3 vertices per axis")
    n = NX * NY * NZ

    def cls(i, m):   # 0, 1, 2 (interior even), 3 (interior odd), 4 (last-1), 5 (last)
        return np.where(i <= 1, i, np.where(m - 1 - i <= 1, 5 - (m - 1 - i), 2 + (i & 1)))
    rep = {0: 0, 1: 1, 2: 4, 3: 3, 4: RN - 2, 5: RN - 1}     # a reference node of every class
    offs = [(a, b, c) for c in range(-2, 3) for b in range(-2, 3) for a in range(-2, 3)]   # ascending column order
    K = len(offs)
    table = np.zeros((6, 6, 6, K, 3, 3))
    present = np.zeros((6, 6, 6, K), bool)
    load = np.zeros((6, 6, 6, 2))                   # rhs_y = h^3 * (load0 + load1 * x/h) per class (x enters linearly)
    rp, rc, rv = ref["rowptr"], ref["col"], ref["val"].reshape(-1, 3, 3)
    # the load of the reference mesh at two different x-offsets gives the linear dependence on x: use the element formula directly
    ref_rhs = ref["rhs"].reshape(-1, 3)[:, 1]
    ref_shift = elasticity3d_p2_kuhn_load_only(ref_nv, shift=1.0)
    href = 1.0 / (ref_nv - 1)
    for cx in range(6):
        for cy in range(6):
            for cz in range(6):
                vx, vy, vz = rep[cx], rep[cy], rep[cz]
                i = vx + RN * (vy + RN * vz)
                cols_i = rc[rp[i]:rp[i + 1]]
                for k, d in enumerate(offs):
                    jx, jy, jz = vx + d[0], vy + d[1], vz + d[2]
                    if not (0 <= jx < RN and 0 <= jy < RN and 0 <= jz < RN):
                        continue
                    j = jx + RN * (jy + RN * jz)
                    pos = np.searchsorted(cols_i, j)
                    if pos < len(cols_i) and cols_i[pos] == j:
                        present[cx, cy, cz, k] = True
                        table[cx, cy, cz, k] = rv[rp[i] + pos] * scale
                # rhs_i(x0) = int (x0 + x) phi_i = x0 * m_i + r_i  with m_i = int phi_i: read m_i from the shifted load
                m_i = ref_shift[i] - ref_rhs[i]
                x_i = vx * href / 2
                load[cx, cy, cz, 0] = (ref_rhs[i] - m_i * x_i) / href ** 4      # part independent of the node's x, in units of h^4
                load[cx, cy, cz, 1] = m_i / href ** 3                             # int phi_i in units of h^3
    ids = np.arange(n, dtype=np.int64)
    ix, iy, iz = ids % NX, (ids // NX) % NY, ids // (NX * NY)
    cxa, cya, cza = cls(ix, NX), cls(iy, NY), cls(iz, NZ)
    pres = present[cxa, cya, cza]                                 # (n, K)
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(pres.sum(axis=1), out=rowptr[1:])
    col = np.empty(int(rowptr[-1]), np.int32)
    val = np.empty((int(rowptr[-1]), 3, 3))
    slot = np.cumsum(pres, axis=1, dtype=np.int16) - 1
    for k, d in enumerate(offs):
        rows = np.flatnonzero(pres[:, k])
        if len(rows) == 0:
            continue
        dst = rowptr[rows] + slot[rows, k]
        col[dst] = (rows + d[0] + NX * (d[1] + NY * d[2])).astype(np.int32)
        val[dst] = table[cxa[rows], cya[rows], cza[rows], k]
    free = np.ones(n, np.uint8)
    dims = (NX, NY, NZ)
    coord = (ix, iy, iz)
    for tag in clamp:
        ax = "xyz".index(tag[0])
        free[coord[ax] == (0 if tag[1] == "0" else dims[ax] - 1)] = 0
    X = np.stack([ix, iy, iz], axis=1) * (hs / 2)
    rhs = np.zeros((n, 3))
    rhs[:, 1] = hs ** 4 * load[cxa, cya, cza, 0] + hs ** 3 * load[cxa, cya, cza, 1] * X[:, 0]
    return dict(n=n, b=3, rowptr=rowptr, col=col, val=val.reshape(-1), free=free, rhs=rhs.reshape(-1), xyz=X, dims=dims)


def elasticity3d_p2_kuhn_load_only(nv, shift=0.0):
    """y-component of the P2 load vector int (x + shift) phi_i on the nv^3-vertex reference mesh (helper of the stencil generator)"""
    h = 1.0 / (nv - 1)
    NX = 2 * nv - 1
    n = NX ** 3
    T = kuhn_tets(nv, nv, nv)
    cv = np.stack([T % nv, (T // nv) % nv, T // (nv * nv)], axis=2)
    fv = 2 * cv
    fe = np.stack([(fv[:, i] + fv[:, j]) // 2 for (i, j) in _P2_EDGES], axis=1)
    fn = np.concatenate([fv, fe], axis=1)
    N = fn[..., 0] + NX * (fn[..., 1] + NX * fn[..., 2])
    Pv = cv * h
    M = np.concatenate([np.ones((len(T), 4, 1)), Pv], axis=2)
    Minv = np.linalg.inv(M)
    vol = np.abs(np.linalg.det(M)) / 6.0
    _, V = _p2_tet_shape_gradients(Minv)
    a, b = 0.5854101966249685, 0.1381966011250105
    xq = np.einsum("qv,tv->tq", np.full((4, 4), b) + (a - b) * np.eye(4), Pv[..., 0]) + shift
    return np.bincount(N.ravel(), weights=np.einsum("tq,tqi,t->ti", xq, V, vol / 4.0).ravel(), minlength=n)


# ---- P1 elasticity with a jumping Young's modulus (BASELINE.json configs[4]; 3D analogue of tests/elasticity/mdim/jump/test_2d_jump_lo.py:4-9:
# a checkerboard of sub-boxes whose modulus differs by `contrast`) -------------------------------------------------------------------
def checkerboard_modulus(box_cells, contrast=1e4):
    """E multiplier per CUBE (cx, cy, cz = cell indices): `contrast` on the odd boxes of a checkerboard of box_cells^3-cell sub-boxes"""
    def f(cx, cy, cz):
        odd = ((cx // box_cells) + (cy // box_cells) + (cz // box_cells)) % 2
        return np.where(odd == 1, float(contrast), 1.0)
    return f


def elasticity3d_kuhn_jump_stencil(nx, ny, nz, cube_modulus, E=1e3, nu=0.15, clamp=("x0",), h=None, origin=(0, 0, 0)):
    """elasticity3d_kuhn(..., jump=piecewise constant per cube) for LARGE uniform meshes.  Every cube of the Kuhn mesh carries the same
    six tets, so A = sum over cubes of E_cube * K_cube with ONE 8x8-vertex block matrix K_cube (unit modulus, read off a 2^3-vertex
    element-assembled mesh): the block row of a vertex is assembled from its (up to) 8 incident cubes, vectorised over all vertices.
    cube_modulus(cx, cy, cz) -> multiplier of E for the cube with lower corner (cx, cy, cz).  Same dict as elasticity3d_kuhn.
    h / origin: the mesh is the sub-box of a larger one starting at vertex `origin` (cube indices passed to cube_modulus and the
    coordinates are global); only the sub-box's own cubes are assembled (a rank's sub-assembled matrix, see box_elasticity3d_jump)."""
    one = elasticity3d_kuhn(2, 2, 2, E=E, nu=nu, clamp=())                      # h = 1: blocks scale linearly with h
    hs = 1.0 / (min(ny, nz) - 1) if h is None else float(h)
    rp, rc, rv = one["rowptr"], one["col"], one["val"].reshape(-1, 3, 3)
    Kc = np.zeros((8, 8, 3, 3))
    conn = np.zeros((8, 8), bool)
    for a in range(8):
        for e in range(rp[a], rp[a + 1]):
            Kc[a, rc[e]] = rv[e] * hs
            conn[a, rc[e]] = True
    corner = [(a & 1, (a >> 1) & 1, (a >> 2) & 1) for a in range(8)]             # vertex a of the 2x2x2 mesh = x + 2 (y + 2 z)
    n = nx * ny * nz
    dirs = [(0, 0, 0)] + _POS_DIRS + [(-a, -b, -c) for (a, b, c) in _POS_DIRS]
    dirs.sort(key=lambda d: d[0] + nx * (d[1] + ny * d[2]))
    K = len(dirs)
    dir_id = {d: k for k, d in enumerate(dirs)}
    ids = np.arange(n, dtype=np.int64)
    ix, iy, iz = ids % nx, (ids // nx) % ny, ids // (nx * ny)
    mask = np.zeros((n, K), bool)
    cols = np.zeros((n, K), np.int64)
    for k, d in enumerate(dirs):
        jx, jy, jz = ix + d[0], iy + d[1], iz + d[2]
        mask[:, k] = (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny) & (jz >= 0) & (jz < nz)
        cols[:, k] = ids + d[0] + nx * (d[1] + ny * d[2])
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(mask.sum(axis=1), out=rowptr[1:])
    col = cols[mask].astype(np.int32)
    del cols
    slot = np.cumsum(mask, axis=1, dtype=np.int16) - 1
    val = np.zeros((int(rowptr[-1]), 3, 3))
    ntet_w = np.zeros(n)                                                          # (for the lumped load) tets incident to the vertex
    for a, (ax, ay, az) in enumerate(corner):
        # the cube in which vertex v is corner a has lower corner v - corner[a]
        cx, cy, cz = ix - ax, iy - ay, iz - az
        ok = (cx >= 0) & (cx <= nx - 2) & (cy >= 0) & (cy <= ny - 2) & (cz >= 0) & (cz <= nz - 2)
        rows = np.flatnonzero(ok)
        Ec = cube_modulus(cx[rows] + origin[0], cy[rows] + origin[1], cz[rows] + origin[2])
        for b2, (bx, by, bz) in enumerate(corner):
            if not conn[a, b2]:
                continue
            k = dir_id[(bx - ax, by - ay, bz - az)]
            dst = rowptr[rows] + slot[rows, k]
            val[dst] += Ec[:, None, None] * Kc[a, b2][None]
        ntet_w[rows] += 6 if a in (0, 7) else 2
    free = np.ones(n, np.uint8)
    dims = (nx, ny, nz)
    coord = (ix, iy, iz)
    for tag in clamp:
        ax_ = "xyz".index(tag[0])
        free[coord[ax_] == (0 if tag[1] == "0" else dims[ax_] - 1)] = 0
    X = np.stack([(ix + origin[0]) * hs, (iy + origin[1]) * hs, (iz + origin[2]) * hs], axis=1).astype(np.float64)
    rhs = np.zeros((n, 3))
    rhs[:, 1] = (hs ** 3 / 24.0) * ntet_w * X[:, 0]
    return dict(n=n, b=3, rowptr=rowptr, col=col, val=val.reshape(-1), free=free, rhs=rhs.reshape(-1), xyz=X)


def _box_halo(n, grid, rank):
    """neighbours and shared-DOF lists of a sub-box of n^3 vertices in a grid = (px, py, pz) box partition (see box_poisson3d)"""
    px, py, pz = grid
    bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)
    peers, ex = [], []
    sel = {-1: np.array([0]), 0: np.arange(n), 1: np.array([n - 1])}
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                qx, qy, qz = bx + dx, by + dy, bz + dz
                if not (0 <= qx < px and 0 <= qy < py and 0 <= qz < pz):
                    continue
                idx = (sel[dx][None, None, :] + n * (sel[dy][None, :, None] + n * sel[dz][:, None, None])).ravel()
                peers.append(qx + px * (qy + py * qz))
                ex.append(idx.astype(np.int32))
    order = np.argsort(peers)
    peers, ex = [peers[i] for i in order], [ex[i] for i in order]
    ghost = np.zeros(n ** 3, bool)
    for q, e in zip(peers, ex):
        if q < rank:
            ghost[e] = True
    return peers, ex, int((~ghost).sum())


def box_elasticity3d_jump(n, grid, rank, box_cells=8, contrast=1e4, E=1e3, nu=0.15):
    """rank's local problem of BASELINE.json configs[4]: P1 elasticity on the cube of ((n-1)p+1)^3 vertices cut into grid = (px, py, pz)
    sub-boxes of n^3 vertices, Young's modulus jumping by `contrast` on a GLOBAL checkerboard of box_cells^3-cell boxes, clamped at the
    global face x = 0, body force (0, x, 0).  Sub-assembled local matrix (only the rank's own cubes), shared DOFs duplicated, halo
    lists as box_poisson3d.  n = 128 on 2x2x2 ranks: 255^3 vertices = 49.7 M DOFs."""
    px, py, pz = grid
    bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)
    gd = ((n - 1) * px + 1, (n - 1) * py + 1, (n - 1) * pz + 1)
    h = 1.0 / (max(gd) - 1)
    origin = (bx * (n - 1), by * (n - 1), bz * (n - 1))
    loc = elasticity3d_kuhn_jump_stencil(n, n, n, checkerboard_modulus(box_cells, contrast), E=E, nu=nu,
                                         clamp=("x0",) if origin[0] == 0 else (), h=h, origin=origin)
    loc["peers"], loc["ex"], loc["n_master"] = _box_halo(n, grid, rank)
    loc["global_dims"] = gd
    return loc
