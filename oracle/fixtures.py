"""Oracle-side mesh fixtures (test infrastructure; see oracle/__init__.py).

Independent restatement of the reference's fixture generator so that the
product's own mesh generator can be cross-checked against it:

* unit square, "right" diagonals, tags as in
  /root/reference/src/CGx/utils/generate_square_mesh.py:28-42 and
  /root/reference/src/CGx/utils/misc.py:99-195 (cells whose vertices all lie in
  [0.25,0.75]^2 -> tag 1, else 2; interface facets -> 4).
* unit cube, 6 tetrahedra per grid cube around the main diagonal
  (misc.py:256-398 for the tagging).

Membrane facets are oriented '+' = intracellular as in
/root/reference/src/CGx/utils/mixed_dim_problem.py:708-729.
"""
from dataclasses import dataclass
import numpy as np


@dataclass
class OracleMesh:
    gdim: int
    x: np.ndarray          # (Nv, gdim) float64, already scaled
    cells: np.ndarray      # (Nc, gdim+1) int64
    cell_tags: np.ndarray  # (Nc,) int64
    mf_verts: np.ndarray   # (Nf, gdim) int64   membrane facets (vertex ids)
    mf_tags: np.ndarray    # (Nf,) int64
    mf_cells: np.ndarray   # (Nf, 2) int64      [intracellular cell, extracellular cell]


def _square_cells(n):
    ix, iy = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v0 = (iy * (n + 1) + ix).ravel()
    v1, v2, v3 = v0 + 1, v0 + n + 1, v0 + n + 2
    # diagonal v0-v3 ("right"), two triangles per grid square
    return np.stack([np.stack([v0, v1, v3], 1), np.stack([v0, v2, v3], 1)], 1).reshape(-1, 3)


def _cube_cells(n):
    iz, iy, ix = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    m = n + 1
    v0 = (iz * m * m + iy * m + ix).ravel()
    v1, v2, v3 = v0 + 1, v0 + m, v0 + m + 1
    v4, v5, v6, v7 = v0 + m * m, v1 + m * m, v2 + m * m, v3 + m * m
    tets = [(v0, v1, v3, v7), (v0, v1, v7, v5), (v0, v5, v7, v4),
            (v0, v3, v2, v7), (v0, v6, v4, v7), (v0, v2, v6, v7)]
    return np.stack([np.stack(t, 1) for t in tets], 1).reshape(-1, 4)


def membrane_facets(cells, cell_tags, intra_tags, membrane_tag_of):
    """All facets shared by an intracellular and an extracellular cell.

    membrane_tag_of(intra_cell_tag array) -> facet tag array."""
    nc, nv = cells.shape
    d = nv - 1
    loc = [tuple(j for j in range(nv) if j != i) for i in range(nv)]
    fac = np.concatenate([cells[:, l] for l in loc], 0)          # (nv*nc, d)
    owner = np.tile(np.arange(nc), nv)
    key = np.sort(fac, 1)
    order = np.lexsort(tuple(key[:, j] for j in range(d - 1, -1, -1)))
    ks = key[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    first = order[:-1][same]
    second = order[1:][same]
    c0, c1 = owner[first], owner[second]
    is_in = np.isin(cell_tags, np.asarray(intra_tags))
    mixed = is_in[c0] != is_in[c1]
    c0, c1, fv = c0[mixed], c1[mixed], key[first][mixed]
    ci = np.where(is_in[c0], c0, c1)
    ce = np.where(is_in[c0], c1, c0)
    o = np.lexsort(tuple(fv[:, j] for j in range(d - 1, -1, -1)))
    fv, ci, ce = fv[o], ci[o], ce[o]
    return fv, membrane_tag_of(cell_tags[ci]), np.stack([ci, ce], 1)


def unit_square(n, scale=1e-6):
    """The C1/C2 fixture: tags intra=1, extra=2, membrane=4."""
    g = np.arange(n + 1) / n
    X, Y = np.meshgrid(g, g, indexing="xy")
    x = np.stack([X.ravel(), Y.ravel()], 1)
    cells = _square_cells(n)
    inside_v = (x[:, 0] <= 0.75) & (x[:, 0] >= 0.25) & (x[:, 1] <= 0.75) & (x[:, 1] >= 0.25)
    tags = np.where(np.all(inside_v[cells], axis=1), 1, 2)
    fv, ft, fc = membrane_facets(cells, tags, [1], lambda t: np.full(t.shape, 4))
    return OracleMesh(2, x * scale, cells, tags, fv, ft, fc)


def unit_cube(n, scale=1e-6):
    """Cube analogue (intra = [0.25,0.75]^3 tag 1, extra 2, membrane 4)."""
    g = np.arange(n + 1) / n
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    x = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
    cells = _cube_cells(n)
    inside_v = np.all((x <= 0.75) & (x >= 0.25), axis=1)
    tags = np.where(np.all(inside_v[cells], axis=1), 1, 2)
    fv, ft, fc = membrane_facets(cells, tags, [1], lambda t: np.full(t.shape, 4))
    return OracleMesh(3, x * scale, cells, tags, fv, ft, fc)


def from_arrays(gdim, x, cells, cell_tags, intra_tags, membrane_tag_of=None):
    """Wrap arbitrary mesh arrays (used to feed the product's meshes to the oracle)."""
    if membrane_tag_of is None:
        membrane_tag_of = lambda t: t
    fv, ft, fc = membrane_facets(np.asarray(cells, np.int64), np.asarray(cell_tags, np.int64),
                                 intra_tags, membrane_tag_of)
    return OracleMesh(gdim, np.asarray(x, float), np.asarray(cells, np.int64),
                      np.asarray(cell_tags, np.int64), fv, ft, fc)
