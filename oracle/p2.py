"""P2 (``fem_order: 2``) restatement of the KNP-EMI forms (TEST INFRASTRUCTURE; see oracle/__init__.py).

The reference selects the element order with ``fem_order`` (utils/mixed_dim_problem.py:207-208) and builds every space as
``("Lagrange", fem_order)`` (KNPEMIx_problem.py:38-42); the forms (KNPEMIx_problem.py:454-655, preconditioner :657-744) are
the same for both orders.  No shipped config, test or golden vector of the reference uses order 2 and DOLFINx cannot run
in this image, so this restatement is **parity unpinned** against the reference: it is pinned instead against what the
forms mean -- exact integrals of the P2 basis (tests/test_oracle_p2.py: monomial integration formulas, patch tests,
agreement with the P1 restatement on fields that both spaces represent, convergence towards a fine P1 solution).

Convention (ours; DOLFINx's P2 numbering depends on its graph reordering and cannot be reproduced): the P2 "node mesh"
lists the mesh vertices first (same ids) and then one node per edge, edges sorted by (lower vertex, higher vertex).  A cell
carries its d+1 vertices followed by its edge nodes in the order (0,1),(0,2),[(0,3),](1,2),[(1,3),(2,3)] of its local
vertices; a membrane facet its d vertices followed by its edge nodes in the same lexicographic order.  With that the
oracle's block layout (field-major blocks, ascending node id restricted to the subdomain) carries over unchanged.
"""
import numpy as np
from scipy.special import roots_jacobi

from .knpemi import KNPEMIOracle


def local_edges(nv):
    return [(i, j) for i in range(nv) for j in range(i + 1, nv)]


def simplex_rule(d, n):
    """Collapsed-coordinate Gauss-Jacobi rule with n^d points on the d-simplex: barycentric points (n^d, d+1) and weights
    summing to 1; exact for polynomials of total degree <= 2n-1."""
    t1, w1 = np.polynomial.legendre.leggauss(n)
    t1, w1 = 0.5 * (t1 + 1.0), 0.5 * w1
    if d == 1:
        return np.stack([1.0 - t1, t1], 1), w1
    t2, w2 = roots_jacobi(n, 1.0, 0.0)
    t2, w2 = 0.5 * (t2 + 1.0), w2 / 4.0                     # weight (1 - t) on [0, 1]
    if d == 2:
        A, B = np.meshgrid(t1, t2, indexing="ij")
        W = np.outer(w1, w2).ravel()
        l2 = B.ravel()
        l1 = (A * (1.0 - B)).ravel()
        lam = np.stack([1.0 - l1 - l2, l1, l2], 1)
        return lam, W / W.sum()
    t3, w3 = roots_jacobi(n, 2.0, 0.0)
    t3, w3 = 0.5 * (t3 + 1.0), w3 / 8.0                     # weight (1 - t)^2 on [0, 1]
    A, B, Cc = np.meshgrid(t1, t2, t3, indexing="ij")
    W = (w1[:, None, None] * w2[None, :, None] * w3[None, None, :]).ravel()
    l3 = Cc.ravel()
    l2 = (B * (1.0 - Cc)).ravel()
    l1 = (A * (1.0 - B) * (1.0 - Cc)).ravel()
    lam = np.stack([1.0 - l1 - l2 - l3, l1, l2, l3], 1)
    return lam, W / W.sum()


def p2_basis(lam):
    """P2 Lagrange basis on a simplex at barycentric points lam (nq, nv): values (nq, nloc) and derivatives with respect to
    the barycentric coordinates (nq, nloc, nv); vertex functions lam_a (2 lam_a - 1) first, then 4 lam_i lam_j per edge."""
    nq, nv = lam.shape
    ed = local_edges(nv)
    N = np.zeros((nq, nv + len(ed)))
    dN = np.zeros((nq, nv + len(ed), nv))
    for a in range(nv):
        N[:, a] = lam[:, a] * (2.0 * lam[:, a] - 1.0)
        dN[:, a, a] = 4.0 * lam[:, a] - 1.0
    for e, (i, j) in enumerate(ed):
        N[:, nv + e] = 4.0 * lam[:, i] * lam[:, j]
        dN[:, nv + e, i] = 4.0 * lam[:, j]
        dN[:, nv + e, j] = 4.0 * lam[:, i]
    return N, dN


def p2_mesh(mesh):
    """The P2 node mesh of a simplicial mesh (any object with gdim, x, cells, cell_tags, mf_verts, mf_tags): a shallow copy
    whose x / cells / mf_verts list nodes as described in the module docstring, plus `n_vertices` and `edges` (ne, 2)."""
    import copy
    d = mesh.gdim
    nv = mesh.x.shape[0]
    cells = np.asarray(mesh.cells, np.int64)
    ce = [np.sort(cells[:, list(p)], axis=1) for p in local_edges(d + 1)]
    keys = np.concatenate([e[:, 0] * nv + e[:, 1] for e in ce])
    uk, inv = np.unique(keys, return_inverse=True)
    edges = np.stack([uk // nv, uk % nv], 1)
    nc = cells.shape[0]
    cell_nodes = np.concatenate([cells] + [nv + inv[i * nc:(i + 1) * nc, None] for i in range(len(ce))], axis=1)
    fv = np.asarray(mesh.mf_verts, np.int64).reshape(-1, d)
    fe = [np.sort(fv[:, list(p)], axis=1) for p in local_edges(d)]
    fcols = [fv]
    for e in fe:
        k = e[:, 0] * nv + e[:, 1]
        pos = np.searchsorted(uk, k)
        assert np.array_equal(uk[pos], k), "a membrane facet edge is not an edge of the mesh"
        fcols.append(nv + pos[:, None])
    m2 = copy.copy(mesh)
    m2.x = np.concatenate([mesh.x, 0.5 * (mesh.x[edges[:, 0]] + mesh.x[edges[:, 1]])], axis=0)
    m2.cells = cell_nodes.astype(mesh.cells.dtype)
    m2.mf_verts = np.concatenate(fcols, axis=1).astype(mesh.mf_verts.dtype)
    m2.n_vertices = nv
    m2.edges = edges
    m2.degree = 2
    return m2


class KNPEMIOracleP2(KNPEMIOracle):
    """KNPEMIOracle on the P2 node mesh of `mesh` (built here unless the mesh already is one)."""

    QUAD_N = 4            # points per direction of the cell rule (exact to degree 7; the integrands reach degree 4)

    def __init__(self, mesh, params, models):
        if getattr(mesh, "degree", 1) != 2:
            mesh = p2_mesh(mesh)
        d = mesh.gdim
        lam, w = simplex_rule(d, self.QUAD_N)
        N, dN = p2_basis(lam)
        self.Mref = np.einsum("q,qa,qb->ab", w, N, N)
        self.T2 = np.einsum("q,qam,qbn->ambn", w, dN, dN)
        self.T3 = np.einsum("q,qe,qam,qbn->eambn", w, N, dN, dN)
        super().__init__(mesh, params, models)

    def _trace_basis(self, qb):
        return p2_basis(qb)[0]

    def lumped_mass(self, M):
        """HRZ lumping (row sums of a P2 mass matrix vanish at the vertices in 2D and are negative in 3D): the diagonal,
        scaled by the same factor on every element so that the total mass is kept."""
        return M.diagonal() / np.trace(self.Mref)

    def _cell_geometry(self, cells):
        geo = super()._cell_geometry(cells)
        vol, g = geo["vol"], geo["g"]
        G = np.einsum("cmi,cni->cmn", g, g)
        return dict(vol=vol, g=g, G=G, M=vol[:, None, None] * self.Mref[None],
                    K=vol[:, None, None] * np.einsum("ambn,cmn->cab", self.T2, G))

    def _cK(self, s, k, coef):
        geo = self.geo[s]
        ck = self.c[s][k][self.cells_s[s]]
        return coef * geo["vol"][:, None, None] * np.einsum("eambn,ce,cmn->cab", self.T3, ck, geo["G"], optimize=True)

    def _tagged(self, tags):
        m = self.mesh
        cells = m.cells[np.isin(m.cell_tags, np.atleast_1d(tags))]
        return cells, super()._cell_geometry(cells)["vol"]

    def integral(self, u, tags, power=1):
        cells, vol = self._tagged(tags)
        if power == 0:
            return float(vol.sum())
        if power == 1:
            return float((vol * (u[cells] @ self.Mref.sum(axis=1))).sum())
        return self.l2_norm(u, tags) ** 2

    def l2_norm(self, u, tags):
        cells, vol = self._tagged(tags)
        uc = u[cells]
        return float(np.sqrt((vol * np.einsum("ca,ab,cb->c", uc, self.Mref, uc)).sum()))
