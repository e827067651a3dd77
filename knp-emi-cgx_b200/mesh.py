"""Synthetic mesh fixtures and mesh-side tables for the KNP-EMI path (host side, setup time).

Replaces, as input providers: src/CGx/utils/generate_square_mesh.py:28-42 with the tagging rules of
src/CGx/utils/misc.py:99-195 (square) and :256-398 (cube), the XDMF read of
src/CGx/utils/mixed_dim_problem.py:634-681 (no HDF5 reader exists in this image: meshes are generated in
memory from the file name or from a ``synthetic_mesh`` block in the YAML file), and the '+' = intracellular
orientation of membrane facets (:708-729).  For continuous P1 fields the orientation only matters through
which cell supplies which trace, so a membrane facet is stored as its vertex list plus its tag.
"""
from dataclasses import dataclass, field
import numpy as np


@dataclass
class Mesh:
    gdim: int
    x: np.ndarray           # (Nv, gdim) float64, scaled
    cells: np.ndarray       # (Nc, gdim+1) int32
    cell_tags: np.ndarray   # (Nc,) int32
    intra_tags: tuple
    extra_tag: int
    mf_verts: np.ndarray    # (Nf, gdim) int32
    mf_tags: np.ndarray     # (Nf,) int32
    grid: tuple = None      # (N, ...) for structured fixtures (used by the block partitioner)
    # distributed extras (None on a single GPU)
    n_owned: int = None
    cell_owned: np.ndarray = None
    mf_owned: np.ndarray = None
    vert_global: np.ndarray = None
    bc_verts: np.ndarray = None     # ingested meshes: local vertices on facets tagged with the config's boundary_tags
    # P2 node mesh (p2_node_mesh): degree 2, x / cells / mf_verts list NODES (vertices, then edge midpoints)
    degree: int = 1
    n_vertices: int = None          # number of mesh vertices = first node id of an edge
    edges: np.ndarray = None        # (Ne, 2) vertex pairs (lower, higher) of the edge nodes, sorted lexicographically


# ------------------------------------------------------------------------------------------ P2 node mesh
def _local_edges(nv):
    return [(i, j) for i in range(nv) for j in range(i + 1, nv)]


def p2_shape(bary):
    """P2 Lagrange basis on a simplex at barycentric points bary (..., nv): vertex functions lam_a (2 lam_a - 1), then
    4 lam_i lam_j per edge in the order of _local_edges."""
    bary = np.asarray(bary, float)
    nv = bary.shape[-1]
    return np.concatenate([bary * (2.0 * bary - 1.0)] + [4.0 * bary[..., i:i + 1] * bary[..., j:j + 1] for i, j in _local_edges(nv)],
                          axis=-1)


def reference_mass(d, degree):
    """int N_a N_b over a d-simplex divided by its measure, exactly (monomial formula int prod lam_i^a_i = d! prod a_i! /
    (sum a_i + d)!); P1: (1 + delta_ab) / ((d + 1)(d + 2))."""
    from math import factorial
    nv = d + 1
    if degree == 1:
        return (1.0 + np.eye(nv)) / ((d + 1) * (d + 2))
    unit = lambda i: tuple(1 if k == i else 0 for k in range(nv))
    add = lambda a, b: tuple(x + y for x, y in zip(a, b))
    polys = [{add(unit(a), unit(a)): 2.0, unit(a): -1.0} for a in range(nv)]
    polys += [{add(unit(i), unit(j)): 4.0} for i, j in _local_edges(nv)]

    def integ(e):
        num = factorial(d)
        for a in e:
            num *= factorial(a)
        return num / factorial(sum(e) + d)
    M = np.zeros((len(polys), len(polys)))
    for a, pa in enumerate(polys):
        for b, pb in enumerate(polys):
            M[a, b] = sum(ca * cb * integ(add(ea, eb)) for ea, ca in pa.items() for eb, cb in pb.items())
    return M


def p2_node_mesh(m: "Mesh") -> "Mesh":
    """The nodes of the ("Lagrange", 2) space (fem_order = 2, utils/mixed_dim_problem.py:207-208, KNPEMIx_problem.py:38-42) as
    a mesh the rest of the host code can treat like the P1 one (dof = node): the mesh vertices keep their ids, every edge adds
    one node at its midpoint (edges sorted by (lower vertex, higher vertex)); a cell lists its gdim+1 vertices and then its
    edge nodes in the order (0,1),(0,2),[(0,3),](1,2),[(1,3),(2,3)] of its local vertices, a membrane facet its gdim vertices
    and then its edge nodes in the same order (knp_mesh_desc::degree = 2 expects exactly this).  Single GPU only."""
    import dataclasses
    if m.degree == 2:
        return m
    if m.n_owned is not None and m.n_owned != m.x.shape[0]:
        raise NotImplementedError("fem_order = 2 runs on one GPU (no partitioned P2 meshes)")
    d, nv = m.gdim, m.x.shape[0]
    cells = np.asarray(m.cells, np.int64)
    nc = cells.shape[0]
    pairs = _local_edges(d + 1)
    keys = np.concatenate([np.minimum(cells[:, i], cells[:, j]) * nv + np.maximum(cells[:, i], cells[:, j]) for i, j in pairs])
    uk, inv = np.unique(keys, return_inverse=True)
    edges = np.stack([uk // nv, uk % nv], 1)
    cell_nodes = np.concatenate([cells] + [nv + inv[k * nc:(k + 1) * nc, None] for k in range(len(pairs))], axis=1)
    fv = np.asarray(m.mf_verts, np.int64).reshape(-1, d)
    fcols = [fv]
    for i, j in _local_edges(d):
        k = np.minimum(fv[:, i], fv[:, j]) * nv + np.maximum(fv[:, i], fv[:, j])
        pos = np.searchsorted(uk, k)
        if pos.size and not np.array_equal(uk[np.minimum(pos, uk.size - 1)], k):
            raise RuntimeError("a membrane facet has an edge that is not an edge of the mesh")
        fcols.append(nv + pos[:, None])
    x = np.concatenate([m.x, 0.5 * (m.x[edges[:, 0]] + m.x[edges[:, 1]])], axis=0)
    bc = None
    if m.bc_verts is not None:          # the tagged facets themselves are not kept by the ingest, only their vertices
        raise NotImplementedError("dirichlet_bcs on ingested meshes with fem_order = 2")
    return dataclasses.replace(m, x=x, cells=cell_nodes.astype(np.int32), mf_verts=np.concatenate(fcols, axis=1).astype(np.int32),
                               degree=2, n_vertices=nv, edges=edges.astype(np.int32), bc_verts=bc, grid=None)


# ------------------------------------------------------------------------------------------ quadrature
def _gauss_jacobi_10(n):
    """Gauss-Jacobi nodes/weights for the weight (1 - t) on [-1, 1] (Golub-Welsch)."""
    al, be = 1.0, 0.0
    k = np.arange(n, dtype=float)
    a = (be ** 2 - al ** 2) / ((2 * k + al + be) * (2 * k + al + be + 2))
    kk = np.arange(1, n, dtype=float)
    b = 2.0 / (2 * kk + al + be) * np.sqrt(kk * (kk + al) * (kk + be) * (kk + al + be)
                                           / ((2 * kk + al + be - 1) * (2 * kk + al + be + 1)))
    J = np.diag(a) + np.diag(b, 1) + np.diag(b, -1)
    t, V = np.linalg.eigh(J)
    w = 2.0 * V[0] ** 2
    return t, w


def facet_quadrature(gdim, npts=6):
    """Degree >= 10 rule on the membrane facet (quadrature_degree 10, mixed_dim_problem.py:732-733):
    6-point Gauss-Legendre on edges (what basix uses); collapsed Gauss-Legendre x Gauss-Jacobi(1,0) with
    6 x 6 points on triangles (basix's 25-point Xiao-Gimbutas table is not available offline).
    Returns barycentric points (nq, gdim) and weights summing to 1."""
    tu, wu = np.polynomial.legendre.leggauss(npts)
    u, wu = 0.5 * (tu + 1.0), 0.5 * wu
    if gdim == 2:
        return np.stack([1.0 - u, u], 1), wu
    tv, wv = _gauss_jacobi_10(npts)
    v, wv = 0.5 * (tv + 1.0), 0.25 * wv
    U, V = np.meshgrid(u, v, indexing="ij")
    l1 = V.ravel()
    l2 = (U * (1.0 - V)).ravel()
    return np.stack([1.0 - l1 - l2, l1, l2], 1), (np.outer(wu, wv).ravel() * 2.0)


# ------------------------------------------------------------------------------------------ grids
def _square_cells(n):
    ix, iy = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64), indexing="xy")
    v0 = (iy * (n + 1) + ix).ravel()
    v1, v2, v3 = v0 + 1, v0 + n + 1, v0 + n + 2
    c = np.empty((v0.size, 2, 3), np.int32)
    c[:, 0, 0], c[:, 0, 1], c[:, 0, 2] = v0, v1, v3
    c[:, 1, 0], c[:, 1, 1], c[:, 1, 2] = v0, v2, v3
    return c.reshape(-1, 3)


def _cube_cells(n):
    iz, iy, ix = np.meshgrid(np.arange(n, dtype=np.int64), np.arange(n, dtype=np.int64),
                             np.arange(n, dtype=np.int64), indexing="ij")
    m = n + 1
    v0 = (iz * m * m + iy * m + ix).ravel()
    v1, v2, v3 = v0 + 1, v0 + m, v0 + m + 1
    v4, v5, v6, v7 = v0 + m * m, v1 + m * m, v2 + m * m, v3 + m * m
    tets = [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4), (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
    c = np.empty((v0.size, 6, 4), np.int32)
    for j, t in enumerate(tets):
        for a in range(4):
            c[:, j, a] = t[a]
    return c.reshape(-1, 4)


def _grid_coords(n, gdim):
    g = np.arange(n + 1) / n
    if gdim == 2:
        X, Y = np.meshgrid(g, g, indexing="xy")
        return np.stack([X.ravel(), Y.ravel()], 1)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)


def membrane_facets(cells, cell_tags, intra_tags, extra_tag, membrane_tag=None):
    """Facets shared by an intracellular and an extracellular cell.  membrane_tag=None -> the facet
    carries the tag of its intracellular cell (production convention, configs/5m/100c.yaml:27-30)."""
    cells = np.asarray(cells)
    nc, nv = cells.shape
    d = nv - 1
    is_in = np.isin(cell_tags, np.asarray(intra_tags))
    is_ex = cell_tags == extra_tag
    nvert = int(cells.max()) + 1
    vi = np.zeros(nvert, bool)
    ve = np.zeros(nvert, bool)
    vi[cells[is_in].ravel()] = True
    ve[cells[is_ex].ravel()] = True
    mv = vi & ve
    cand = np.flatnonzero((mv[cells].sum(1) >= d) & (is_in | is_ex))
    cc = cells[cand]
    loc = [tuple(j for j in range(nv) if j != i) for i in range(nv)]
    fac = np.concatenate([cc[:, l] for l in loc], 0)
    owner = np.tile(cand, nv)
    onm = np.all(mv[fac], axis=1)
    fac, owner = fac[onm], owner[onm]
    key = np.sort(fac, 1)
    order = np.lexsort(tuple(key[:, j] for j in range(d - 1, -1, -1)))
    ks, ow = key[order], owner[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    c0, c1 = ow[:-1][same], ow[1:][same]
    mixed = is_in[c0] != is_in[c1]
    c0, c1, fv = c0[mixed], c1[mixed], ks[:-1][same][mixed]
    ci = np.where(is_in[c0], c0, c1)
    tags = cell_tags[ci] if membrane_tag is None else np.full(ci.shape, membrane_tag)
    return fv.astype(np.int32), tags.astype(np.int32)


def unit_square_fixture(n=32, scale=1e-6):
    """The reference CI fixture: intra 1 = cells with all vertices in [0.25,0.75]^2, extra 2, membrane 4."""
    x = _grid_coords(n, 2)
    cells = _square_cells(n)
    inside = (x[:, 0] <= 0.75) & (x[:, 0] >= 0.25) & (x[:, 1] <= 0.75) & (x[:, 1] >= 0.25)
    tags = np.where(np.all(inside[cells], axis=1), 1, 2).astype(np.int32)
    fv, ft = membrane_facets(cells, tags, (1,), 2, membrane_tag=4)
    return Mesh(2, x * scale, cells, tags, (1,), 2, fv, ft, grid=(n, n))


def unit_cube_fixture(n=8, scale=1e-6):
    x = _grid_coords(n, 3)
    cells = _cube_cells(n)
    inside = np.all((x <= 0.75) & (x >= 0.25), axis=1)
    tags = np.where(np.all(inside[cells], axis=1), 1, 2).astype(np.int32)
    fv, ft = membrane_facets(cells, tags, (1,), 2, membrane_tag=4)
    return Mesh(3, x * scale, cells, tags, (1,), 2, fv, ft, grid=(n, n, n))


def _inside(offs, bs, fill, shape):
    """Which grid squares / cubes of an n/m-wide block belong to its biological cell.  offs: per axis (x, y[, z]) the offset
    of the square inside its block, broadcastable against each other.
      shape None          : the middle `fill` fraction of the block, a square / cube (BASELINE C3 / C4)
      shape {"plates": ..}: dense-tissue-like (BASELINE C5): inside the same middle region the cell is a stack of thin plates
                            normal to the last axis (`thickness` grid units thick, one every `pitch`) joined by a spine of
                            `spine` units along the first axis -- membrane facets / cells >= 0.2 and membrane vertices /
                            vertices >= 0.5 for thickness 1, pitch 2 (real tissue: 0.30 / 0.88)."""
    lo = int(round(bs * (1 - fill) / 2))
    hi = bs - lo
    box = None
    for o in offs:
        a = (o >= lo) & (o < hi)
        box = a if box is None else box & a
    if not shape:
        return box
    t, pitch, spine = int(shape.get("thickness", 1)), int(shape.get("pitch", 2)), int(shape.get("spine", 1))
    plate = ((offs[-1] - lo) % pitch) < t
    return box & (plate | ((offs[0] - lo) < spine))


def cell_array_mesh(gdim, n, m, scale=1e-6, fill=0.5, first_tag=2, extra_tag=1, shape=None):
    """Synthetic tissue block (BASELINE configs C3/C4/C5): an m^gdim array of biological cells.
    Cell (p,q[,r]) occupies the middle `fill` fraction of its n/m-wide block (as a square / cube, or as a stack of thin
    plates: see _inside); intracellular tags first_tag..first_tag+m^gdim-1, extracellular tag `extra_tag`, membrane tag =
    intracellular tag."""
    assert n % m == 0
    bs = n // m
    idx = np.arange(n)
    blk, off = idx // bs, idx % bs
    if gdim == 2:
        IN = _inside([off[None, :], off[:, None]], bs, fill, shape)      # [iy, ix]
        tag = first_tag + blk[:, None] * m + blk[None, :]      # q*m + p
        gt = np.where(IN, tag, extra_tag).ravel()
        cells = _square_cells(n)
        tags = np.repeat(gt, 2).astype(np.int32)
    else:
        IN = _inside([off[None, None, :], off[None, :, None], off[:, None, None]], bs, fill, shape)   # [iz, iy, ix]
        tag = first_tag + (blk[:, None, None] * m + blk[None, :, None]) * m + blk[None, None, :]
        gt = np.where(IN, tag, extra_tag).ravel()
        cells = _cube_cells(n)
        tags = np.repeat(gt, 6).astype(np.int32)
    x = _grid_coords(n, gdim)
    intra = tuple(range(first_tag, first_tag + m ** gdim))
    fv, ft = membrane_facets(cells, tags, intra, extra_tag)
    return Mesh(gdim, x * scale, cells, tags, intra, extra_tag, fv, ft, grid=(n,) * gdim)


def rank_grid(gdim, size):
    """Process grid of the block partition: the prime factors of `size`, largest first, go to the axis that currently has
    the fewest blocks (2 -> 2x1(x1), 4 -> 2x2(x1), 8 -> 4x2 / 2x2x2, 6 -> 3x2 ...)."""
    dims = [1] * gdim
    f, rest, primes = 2, size, []
    while rest > 1:
        while rest % f == 0:
            primes.append(f)
            rest //= f
        f += 1
    for q in sorted(primes, reverse=True):
        dims[int(np.argmin(dims))] *= q
    return tuple(dims)


class BlockOwner:
    """Vertex -> rank map of the block partition of an (n+1)^gdim structured vertex grid, evaluated by formula (no global
    array): axis a is cut into dims[a] slabs of near-equal vertex counts; rank = sum_a block_a * stride_a (x fastest)."""

    def __init__(self, gdim, n, size):
        self.gdim, self.n, self.size = gdim, n, size
        self.dims = rank_grid(gdim, size)
        self.cuts = [np.array([(b * (n + 1)) // d for b in range(d + 1)], np.int64) for d in self.dims]

    def box(self, rank):
        lo, hi, r = [], [], rank
        for a in range(self.gdim):
            b = r % self.dims[a]
            r //= self.dims[a]
            lo.append(int(self.cuts[a][b]))
            hi.append(int(self.cuts[a][b + 1]))
        return lo, hi

    def __call__(self, gid):
        gid = np.asarray(gid, np.int64)
        m = self.n + 1
        rank = np.zeros(gid.shape, np.int64)
        stride, rest = 1, gid
        for a in range(self.gdim):
            i = rest % m
            rest = rest // m
            rank += (np.searchsorted(self.cuts[a], i, side="right") - 1) * stride
            stride *= self.dims[a]
        return rank.astype(np.int32)


def cell_array_mesh_local(gdim, n, m, rank, size, scale=1e-6, fill=0.5, first_tag=2, extra_tag=1, shape=None):
    """This rank's part of cell_array_mesh(gdim, n, m, ...) under the block partition, generated WITHOUT building the global
    mesh: owned vertices first (ascending global id), then ghosts; every cell / membrane facet touching an owned vertex;
    ownership flags for functionals -- the same local mesh partition.partition_mesh(global mesh, owner=BlockOwner) returns.
    Returns (Mesh, info) with info = {owner_of, l2g, rank, size}."""
    assert n % m == 0
    own = BlockOwner(gdim, n, size)
    lo, hi = own.box(rank)
    # grid squares / cubes touching an owned vertex: one layer around the owned vertex box
    rng = [np.arange(max(lo[a] - 1, 0), min(hi[a], n), dtype=np.int64) for a in range(gdim)]
    mv = n + 1
    bs = n // m
    off = [r % bs for r in rng]
    blk = [r // bs for r in rng]
    if gdim == 2:
        IY, IX = np.meshgrid(rng[1], rng[0], indexing="ij")
        v0 = (IY * mv + IX).ravel()
        v1, v2, v3 = v0 + 1, v0 + mv, v0 + mv + 1
        cells = np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], 1).reshape(-1, 3)
        IN = _inside([off[0][None, :], off[1][:, None]], bs, fill, shape)
        tag = first_tag + blk[1][:, None] * m + blk[0][None, :]
        tags = np.repeat(np.where(IN, tag, extra_tag).ravel(), 2)
    else:
        IZ, IY, IX = np.meshgrid(rng[2], rng[1], rng[0], indexing="ij")
        v0 = (IZ * mv * mv + IY * mv + IX).ravel()
        v1, v2, v3 = v0 + 1, v0 + mv, v0 + mv + 1
        v4, v5, v6, v7 = v0 + mv * mv, v1 + mv * mv, v2 + mv * mv, v3 + mv * mv
        tets = [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4), (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
        cells = np.stack([np.stack(t, 1) for t in tets], 1).reshape(-1, 4)
        IN = _inside([off[0][None, None, :], off[1][None, :, None], off[2][:, None, None]], bs, fill, shape)
        tag = first_tag + (blk[2][:, None, None] * m + blk[1][None, :, None]) * m + blk[0][None, None, :]
        tags = np.repeat(np.where(IN, tag, extra_tag).ravel(), 6)
    cell_owner = own(cells)
    keep = (cell_owner == rank).any(axis=1)
    cells, tags, cell_owner = cells[keep], tags[keep].astype(np.int32), cell_owner[keep]
    used = np.unique(cells.ravel())
    mine = own(used) == rank
    l2g = np.concatenate([used[mine], used[~mine]])
    lcells = np.searchsorted(used, cells)                  # index into `used`, then into the owned-first order
    pos = np.empty(used.size, np.int64)
    pos[np.concatenate([np.flatnonzero(mine), np.flatnonzero(~mine)])] = np.arange(used.size)
    lcells = pos[lcells].astype(np.int32)
    # coordinates from the structured index
    idx, rest = [], l2g
    for a in range(gdim):
        idx.append(rest % mv)
        rest = rest // mv
    x = np.stack([i / n for i in idx], 1) * scale
    # a cell / facet is integrated by the rank that owns its lowest-numbered (global) vertex
    cell_owned = (np.take_along_axis(cell_owner, np.argmin(cells, axis=1)[:, None], 1)[:, 0] == rank).astype(np.uint8)
    intra = tuple(range(first_tag, first_tag + m ** gdim))
    fv, ft = membrane_facets(lcells, tags, intra, extra_tag)
    n_owned = int(mine.sum())
    if fv.size:
        f_has = (fv < n_owned).any(axis=1)
        fv, ft = fv[f_has], ft[f_has]
        fg = l2g[fv]
        # facets sorted like the global generator sorts them (by their sorted global vertex ids)
        key = np.sort(fg, 1)
        order = np.lexsort(tuple(key[:, j] for j in range(gdim - 1, -1, -1)))
        fv, ft, fg, key = fv[order], ft[order], fg[order], key[order]
        fv = np.searchsorted(used, key)
        fv = pos[fv].astype(np.int32)
        mf_owned = (own(key[:, 0]) == rank).astype(np.uint8)
    else:
        mf_owned = np.zeros(0, np.uint8)
    local = Mesh(gdim, x, lcells, tags, intra, extra_tag, fv.astype(np.int32).reshape(-1, gdim), ft.astype(np.int32),
                 grid=(n,) * gdim, n_owned=n_owned, cell_owned=cell_owned, mf_owned=mf_owned, vert_global=l2g)
    return local, dict(owner_of=own, l2g=l2g, rank=rank, size=size)


def boundary_vertices(m: Mesh):
    """Vertices of the exterior boundary (the facets mark_boundaries_square / _cube tag PARTIAL_OMEGA, misc.py:139-186):
    by structured index on the generated fixtures (works on a rank's slab through its global vertex ids), otherwise the
    vertices of the facets that belong to exactly one cell of a GLOBAL mesh."""
    if m.degree == 2:
        # nodes of the P2 space on the exterior boundary: the vertices AND the edge nodes of the facets that belong to one cell
        # (an edge between two boundary vertices can cross the interior, so the facets decide)
        from .xdmf import _entity_keys
        d, nv = m.gdim, m.n_vertices
        cv = m.cells[:, :d + 1]
        fac = np.concatenate([cv[:, [j for j in range(d + 1) if j != i]] for i in range(d + 1)], 0)
        key, srt, exact = _entity_keys(fac, nv)
        uk, first, cnt = np.unique(key, return_index=True, return_counts=True)
        bf = srt[first[cnt == 1]].astype(np.int64)                       # boundary facets (sorted vertex ids)
        ek = m.edges[:, 0].astype(np.int64) * nv + m.edges[:, 1]
        nodes = [bf.ravel()]
        for i, j in _local_edges(d):
            nodes.append(nv + np.searchsorted(ek, np.minimum(bf[:, i], bf[:, j]) * nv + np.maximum(bf[:, i], bf[:, j])))
        return np.unique(np.concatenate(nodes)).astype(np.int32)
    if m.grid is not None:
        n = m.grid[0]
        gid = np.arange(m.x.shape[0], dtype=np.int64) if m.vert_global is None else np.asarray(m.vert_global, np.int64)
        on = np.zeros(gid.shape, bool)
        rest = gid
        for _ in range(m.gdim):
            i = rest % (n + 1)
            rest = rest // (n + 1)
            on |= (i == 0) | (i == n)
        return np.flatnonzero(on).astype(np.int32)
    if m.vert_global is not None:
        raise RuntimeError("exterior facets of a partitioned unstructured mesh must be found before the partition")
    from .xdmf import _entity_keys
    nvc = m.cells.shape[1]
    fac = np.concatenate([m.cells[:, [j for j in range(nvc) if j != i]] for i in range(nvc)], 0)
    key, srt, exact = _entity_keys(fac, m.x.shape[0])
    uk, first, cnt = np.unique(key, return_index=True, return_counts=True)
    return np.unique(srt[first[cnt == 1]].astype(np.int64).ravel()).astype(np.int32)


def from_xdmf(mesh_file, facet_file, ct_name, ft_name, intra_tags, extra_tag, boundary_tags, scale):
    """Mesh ingest (utils/mixed_dim_problem.py:634-681 + the coordinate scaling of :681): cells and cell tags from
    `mesh_file`, facet tags from `facet_file`.  Membrane facets are the facets between an intracellular and an extracellular
    cell (sorted by their vertex tuples, like the generated fixtures) and carry the tag the facet file gives them (-1 when
    it gives none: such facets are in no dS(gamma_tags) integral); `bc_verts` are the vertices of the facets tagged with
    one of `boundary_tags`."""
    from .xdmf import read_xdmf_mesh, match_entities
    d = read_xdmf_mesh(mesh_file, facet_file, ct_name, ft_name)
    cells, tags, nv = d["cells"], d["cell_tags"], d["x"].shape[0]
    fv, _own = membrane_facets(cells, tags, intra_tags, extra_tag)
    ft = match_entities(d["facets"], d["facet_tags"], fv, nv, -1) if fv.size else np.zeros(0, np.int32)
    on_b = np.isin(d["facet_tags"], np.asarray(boundary_tags, np.int32)) if len(boundary_tags) else np.zeros(0, bool)
    bc = np.unique(d["facets"][on_b].ravel()).astype(np.int32) if on_b.any() else np.zeros(0, np.int32)
    return Mesh(d["gdim"], d["x"] * scale, cells, tags, tuple(intra_tags), extra_tag, fv.reshape(-1, d["gdim"]), ft, bc_verts=bc)


def export_xdmf(m: Mesh, mesh_path, facet_path, scale=1.0, boundary_tag=3, default_tag=None, fmt="HDF"):
    """Write a generated mesh the way utils/generate_square_mesh.py:37-42 writes its fixture: `mesh_path` holds the mesh and
    the cell tags (grid "ct"), `facet_path` the mesh and the facet tags (grid "ft": exterior facets `boundary_tag`,
    membrane facets their tag).  Coordinates are divided by `scale` (the file is in mesh units)."""
    from .xdmf import write_xdmf_mesh, _entity_keys
    nvc = m.cells.shape[1]
    fac = np.concatenate([m.cells[:, [j for j in range(nvc) if j != i]] for i in range(nvc)], 0)
    key, srt, _e = _entity_keys(fac, m.x.shape[0])
    _u, first, cnt = np.unique(key, return_index=True, return_counts=True)
    ext = srt[first[cnt == 1]].astype(np.int64)
    ent = np.concatenate([ext, np.asarray(m.mf_verts, np.int64)], 0)
    val = np.concatenate([np.full(ext.shape[0], boundary_tag, np.int32), np.asarray(m.mf_tags, np.int32)])
    x = m.x / scale
    write_xdmf_mesh(mesh_path, x, m.cells, {"ct": (m.cells, m.cell_tags)}, fmt)
    write_xdmf_mesh(facet_path, x, m.cells, {"ft": (ent, val)}, fmt)


def from_descriptor(desc, scale):
    """``synthetic_mesh`` YAML block -> Mesh.  kinds: square, cube (reference fixtures), cell_array."""
    kind = desc.get("kind", "square")
    n = int(desc.get("N", 32))
    if kind == "square":
        return unit_square_fixture(n, scale)
    if kind == "cube":
        return unit_cube_fixture(n, scale)
    if kind == "cell_array":
        return cell_array_mesh(int(desc.get("dim", 2)), n, int(desc.get("cells_per_dim", 8)), scale,
                               float(desc.get("fill", 0.5)), int(desc.get("first_tag", 2)),
                               int(desc.get("extra_tag", 1)), desc.get("shape"))
    raise ValueError(f"unknown synthetic_mesh kind {kind!r}")
