"""The P2 restatement (oracle/p2.py) against what the forms mean: exact integrals of the P2 basis, and the pinned P1
restatement on everything both spaces represent (the P1 space is a subspace of the P2 space, so the bilinear and linear
forms must agree on interpolated P1 functions when the frozen fields are P1 as well).  CPU only."""
from math import factorial

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fixtures
from oracle.knpemi import KNPEMIOracle, OracleParams
from oracle.p2 import KNPEMIOracleP2, local_edges, p2_basis, p2_mesh, simplex_rule


def _monomial_integral(alpha):
    """int over the reference d-simplex of prod lam_i^alpha_i, divided by the simplex volume: d! prod alpha_i! / (|alpha| + d)!"""
    d = len(alpha) - 1
    num = factorial(d)
    for a in alpha:
        num *= factorial(a)
    return num / factorial(sum(alpha) + d)


@pytest.mark.parametrize("d", [1, 2, 3])
def test_simplex_rule_integrates_monomials_exactly(d):
    lam, w = simplex_rule(d, 3)
    assert abs(w.sum() - 1.0) < 1e-14
    rng = np.random.default_rng(d)
    for _ in range(40):
        alpha = rng.integers(0, 3, d + 1)
        if alpha.sum() > 5:
            continue
        got = float((w * np.prod(lam ** alpha[None, :], axis=1)).sum())
        assert abs(got - _monomial_integral(alpha)) < 1e-14


@pytest.mark.parametrize("d", [2, 3])
def test_p2_basis_is_nodal_and_mass_matrix_is_the_textbook_one(d):
    nv = d + 1
    nodes = [np.eye(nv)[a] for a in range(nv)] + [0.5 * (np.eye(nv)[i] + np.eye(nv)[j]) for i, j in local_edges(nv)]
    N, dN = p2_basis(np.array(nodes))
    assert np.allclose(N, np.eye(len(nodes)), atol=1e-15)
    lam, w = simplex_rule(d, 4)
    Nq, dNq = p2_basis(lam)
    # partition of unity; sum_a dN_a/dlam_m = 4 sum(lam) - 1 = 3 for every m, i.e. a zero physical gradient (sum_m grad lam_m = 0)
    assert np.allclose(Nq.sum(axis=1), 1.0) and np.allclose(dNq.sum(axis=1), 3.0)
    M = np.einsum("q,qa,qb->ab", w, Nq, Nq)
    if d == 2:       # |T| / 180 * [[6, -1, ...], edges 32 / 16, vertex-opposite edge -4]
        assert np.allclose(np.diag(M), np.array([6, 6, 6, 32, 32, 32]) / 180.0)
        assert abs(M[0, 1] + 1 / 180.0) < 1e-15 and abs(M[0, 5] + 4 / 180.0) < 1e-15 and abs(M[0, 3]) < 1e-15
        assert np.allclose(M.sum(axis=1)[:3], 0.0, atol=1e-15)                            # why row-sum lumping fails for P2
    else:
        assert np.allclose(np.diag(M), np.array([6] * 4 + [32] * 6) / 420.0)
        assert np.allclose(M.sum(axis=1)[:4], -1.0 / 20.0)


def _p2_interpolation(o1, o2):
    """Block-diagonal matrix that maps P1 unknown vectors of o1 to the P2 unknown vectors of o2 (vertex values kept, edge
    nodes = mean of the end points)."""
    m2 = o2.mesh
    nv = m2.n_vertices
    blocks = []
    for s in range(2):
        S2, r1 = o2.S[s], o1.r[s]
        rows, cols, vals = [], [], []
        for i, node in enumerate(S2):
            if node < nv:
                rows.append(i), cols.append(r1[node]), vals.append(1.0)
            else:
                a, b = m2.edges[node - nv]
                rows += [i, i]
                cols += [r1[a], r1[b]]
                vals += [0.5, 0.5]
        assert min(cols) >= 0
        I = sp.csr_matrix((vals, (rows, cols)), shape=(S2.size, o1.ns[s]))
        blocks += [I] * 4
    return sp.block_diag(blocks).tocsr()


def _pair(gdim, models, **kw):
    mesh = fixtures.unit_square(8) if gdim == 2 else fixtures.unit_cube(4)
    p = OracleParams(**kw)
    o1 = KNPEMIOracle(mesh, p, models)
    o2 = KNPEMIOracleP2(mesh, p, models)
    rng = np.random.default_rng(7)
    nv = mesh.x.shape[0]
    m2 = o2.mesh
    lift = lambda v: np.concatenate([v, 0.5 * (v[m2.edges[:, 0]] + v[m2.edges[:, 1]])])
    for o in (o1, o2):
        o.t = 0.0
    for s in range(2):
        for k in range(3):
            v = o1.c[s][k] * (1.0 + 0.2 * rng.random(nv))
            o1.c[s][k], o2.c[s][k] = v, lift(v)
        v = o1.phi[s] + 0.01 * rng.random(nv)
        o1.phi[s], o2.phi[s] = v, lift(v)
    o1.phi_m, o2.phi_m = o1.phi[0] - o1.phi[1], o2.phi[0] - o2.phi[1]
    for j in range(3):
        v = o1.gates[j] * (1.0 + 0.1 * rng.random(nv))
        o1.gates[j], o2.gates[j] = v, lift(v)
    return o1, o2


@pytest.mark.parametrize("gdim", [2, 3])
@pytest.mark.parametrize("models", [[("HH", None), ("ATP", None), ("NeuronalCT", None)], [("Passive", None)]])
def test_p2_forms_agree_with_p1_forms_on_the_p1_subspace(gdim, models):
    o1, o2 = _pair(gdim, models, stimulus_region=(0, 0.2e-6, 0.6e-6))
    A1, b1 = o1.assemble(2.5e-5)
    A2, b2 = o2.assemble(2.5e-5)
    I = _p2_interpolation(o1, o2)
    G = (I.T @ A2 @ I).tocsr()
    scale = abs(A1).max()
    assert abs(G - A1).max() < 1e-12 * scale
    assert np.abs(I.T @ b2 - b1).max() < 1e-12 * np.abs(b1).max()
    P1, P2 = o1.assemble_P(), o2.assemble_P()
    assert abs((I.T @ P2 @ I).tocsr() - P1).max() < 1e-12 * abs(P1).max()
    # functionals of P1 fields
    for tags in ([1], [2], [1, 2]):
        u1, u2 = o1.c[1][1], o2.c[1][1]
        for power in (0, 1, 2):
            assert abs(o2.integral(u2, tags, power) - o1.integral(u1, tags, power)) < 1e-12 * abs(o1.integral(u1, tags, power))
    assert abs(o2.stimulus_area() - o1.stimulus_area()) < 1e-13 * o1.stimulus_area()


def test_p2_stiffness_reproduces_quadratic_energies():
    mesh = fixtures.unit_square(4, scale=1.0)
    o = KNPEMIOracleP2(mesh, OracleParams(), [("Passive", None)])
    x = o.mesh.x
    K = sp.lil_matrix((x.shape[0], x.shape[0]))
    M = sp.lil_matrix((x.shape[0], x.shape[0]))
    geo = o._cell_geometry(o.mesh.cells)
    for c, nodes in enumerate(o.mesh.cells):
        K[np.ix_(nodes, nodes)] += geo["K"][c]
        M[np.ix_(nodes, nodes)] += geo["M"][c]
    K, M = K.tocsr(), M.tocsr()
    u, v = x[:, 0] ** 2 + x[:, 0] * x[:, 1], x[:, 1] ** 2 - 3.0 * x[:, 0]
    # grad u = (2x + y, x), grad v = (-3, 2y): int over the unit square = -3 (1 + 1/2) + 2 * (1/2 * 1/2)
    assert abs(v @ (K @ u) - (-4.5 + 0.5)) < 1e-12
    # int u v over the unit square: (x^2 + x y)(y^2 - 3 x) = x^2 y^2 - 3 x^3 + x y^3 - 3 x^2 y
    assert abs(v @ (M @ u) - (1 / 9 - 3 / 4 + 1 / 8 - 1 / 2)) < 1e-12


def test_p2_time_loop_converges_towards_the_fine_p1_solution():
    """Three steps with the direct solver: the P2 solution on the 8 x 8 mesh is closer to the P1 solution on the 32 x 32 mesh
    than the P1 solution on the 8 x 8 mesh is (membrane potential norm over the intracellular cells)."""
    models = [("HH", None)]
    runs = {}
    for name, cls, n in (("p1_coarse", KNPEMIOracle, 8), ("p2_coarse", KNPEMIOracleP2, 8), ("p1_fine", KNPEMIOracle, 32)):
        o = cls(fixtures.unit_square(n), OracleParams(), models)
        o.run(3, "direct")
        runs[name] = (o.l2_norm(o.phi[0], [1]), o.l2_norm(o.phi[1], [2]))
    for i in range(2):
        e1 = abs(runs["p1_coarse"][i] - runs["p1_fine"][i])
        e2 = abs(runs["p2_coarse"][i] - runs["p1_fine"][i])
        assert e2 < e1, (runs, i)


def test_p2_schur_preconditioner_iteration_counts():
    from oracle.amg import SchurPC
    o = KNPEMIOracleP2(fixtures.unit_square(16), OracleParams(), [("HH", None)])
    pc = SchurPC(o)
    its = o.run(3, "gmres", rtol=1e-9, Pinv_factory=lambda P: pc)
    assert max(its) <= 60, its
