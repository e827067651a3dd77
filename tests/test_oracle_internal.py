"""Independent checks of the oracle's own arithmetic (the reference holds no matrices/vectors to pin it with):
exact element integrals via sympy, explicit-quadrature loop assembly on a tiny mesh, facet rule exactness,
nullspace property, closed-form gate update."""
import numpy as np
import pytest
import sympy as sy
from oracle.fixtures import unit_square, unit_cube, OracleMesh, membrane_facets
from oracle.knpemi import KNPEMIOracle, OracleParams
from oracle.quadrature import facet_rule, interval_rule, triangle_rule
from conftest import MODELS_TEST


def _exact_element(xv):
    """Exact P1 mass and stiffness matrices of a simplex by symbolic integration over the reference cell."""
    d = len(xv) - 1
    xi = sy.symbols(f"xi0:{d}")
    N = [1 - sum(xi)] + list(xi)
    J = sy.Matrix([[xv[a + 1][i] - xv[0][i] for a in range(d)] for i in range(d)])
    detJ = abs(J.det())
    Jit = J.inv().T
    grads = [Jit * sy.Matrix([sy.diff(n, x) for x in xi]) for n in N]

    def integ(f):
        for k in range(d - 1, -1, -1):
            f = sy.integrate(f, (xi[k], 0, 1 - sum(xi[:k])))
        return f
    M = [[integ(N[a] * N[b]) * detJ for b in range(d + 1)] for a in range(d + 1)]
    K = [[integ((grads[a].T * grads[b])[0]) * detJ for b in range(d + 1)] for a in range(d + 1)]
    return np.array(M, float), np.array(K, float)


@pytest.mark.parametrize("gdim", [2, 3])
def test_cell_geometry_against_sympy(gdim):
    rng = np.random.default_rng(3)
    xv = rng.random((gdim + 1, gdim)) + np.eye(gdim + 1, gdim)
    mesh = OracleMesh(gdim, xv, np.arange(gdim + 1)[None, :], np.array([1]), np.zeros((0, gdim), int),
                      np.zeros(0, int), np.zeros((0, 2), int))
    o = KNPEMIOracle.__new__(KNPEMIOracle)
    o.mesh = mesh
    geo = KNPEMIOracle._cell_geometry(o, mesh.cells)
    M, K = _exact_element([[sy.Rational(float(v)).limit_denominator(10**12) for v in row] for row in xv])
    np.testing.assert_allclose(geo["M"][0], M, rtol=1e-10)
    np.testing.assert_allclose(geo["K"][0], K, rtol=1e-9, atol=1e-12)


def test_facet_rules_exact_to_degree_10():
    b, w = interval_rule(6)
    for k in range(12):
        assert abs((w * b[:, 1] ** k).sum() - 1.0 / (k + 1)) < 1e-14
    b, w = triangle_rule(6)
    from math import factorial
    for i in range(11):
        for j in range(11 - i):
            exact = 2.0 * factorial(i) * factorial(j) / factorial(i + j + 2)     # mean value over the triangle
            assert abs((w * b[:, 1] ** i * b[:, 2] ** j).sum() - exact) < 1e-14, (i, j)


def _loop_assemble(o, t):
    """Slow restatement with explicit loops and a generic quadrature on cells (degree-4 exact Duffy rule),
    straight from the weak form KNPEMIx_problem.py:598-614,633-642; used on tiny meshes only."""
    p, m = o.p, o.mesh
    d = m.gdim
    A = np.zeros((o.n, o.n))
    b = np.zeros(o.n)
    qb, qw = triangle_rule(4) if d == 2 else (None, None)
    for s in range(2):
        for cell in o.cells_s[s]:
            x = m.x[cell]
            J = (x[1:] - x[0]).T
            vol = abs(np.linalg.det(J)) / 2
            G = np.vstack([-np.linalg.inv(J).sum(0), np.linalg.inv(J)])
            for q in range(qw.size):
                N = qb[q]
                wq = qw[q] * vol
                cprev = [N @ o.c[s][k][cell] for k in range(3)]
                for a in range(3):
                    for bb in range(3):
                        gg = G[a] @ G[bb]
                        for k in range(3):
                            rk, ck = o.row(s, k, cell[a]), o.row(s, k, cell[bb])
                            rp, cp = o.row(s, 3, cell[a]), o.row(s, 3, cell[bb])
                            A[rk, ck] += wq * (N[a] * N[bb] + p.dt * p.D[k] * gg)
                            A[rk, cp] += wq * p.dt * (p.D[k] * p.z[k] / p.psi) * cprev[k] * gg
                            A[rp, ck] += wq * p.dt * p.z[k] * p.D[k] * gg
                            A[rp, cp] += wq * p.dt * (p.D[k] * p.z[k] ** 2 / p.psi) * cprev[k] * gg
                    for k in range(3):
                        b[o.row(s, k, cell[a])] += wq * cprev[k] * N[a]
    # facets
    o._stim_area = o.stimulus_area()
    t_mod = np.mod(t + 1e-12, p.T_stim)
    fb, fw = facet_rule(d)
    for f, fv in enumerate(m.mf_verts):
        area = o.farea[f]
        for q in range(fw.size):
            N = fb[q]
            wq = fw[q] * area
            ci = [N @ o.c[0][k][fv] for k in range(3)]
            ce = [N @ o.c[1][k][fv] for k in range(3)]
            pm = N @ o.phi_m[fv]
            gq = [N @ o.gates[j][fv] for j in range(3)]
            xq = N @ m.x[fv]
            I = o.channel_currents(int(m.mf_tags[f]), [np.array(v) for v in ci], [np.array(v) for v in ce],
                                   np.array(pm), [np.array(v) for v in gq], xq, t_mod)
            Itot = sum(I)
            den = [sum(p.D[j] * p.z[j] ** 2 * c[j] for j in range(3)) for c in (ci, ce)]
            for a in range(d):
                for s, sign, cs in ((0, 1.0, ci), (1, -1.0, ce)):
                    rp = o.row(s, 3, fv[a])
                    b[rp] -= sign / p.F * wq * (p.dt * Itot - p.C_M * pm) * N[a]
                    for k in range(3):
                        al = p.D[k] * p.z[k] ** 2 * cs[k] / den[s]
                        rk = o.row(s, k, fv[a])
                        b[rk] -= sign / (p.F * p.z[k]) * wq * (p.dt * I[k] - al * p.C_M * pm) * N[a]
                        for bb in range(d):
                            A[rk, o.row(0, 3, fv[bb])] += sign * al * p.C_M / (p.F * p.z[k]) * wq * N[a] * N[bb]
                            A[rk, o.row(1, 3, fv[bb])] -= sign * al * p.C_M / (p.F * p.z[k]) * wq * N[a] * N[bb]
                    for bb in range(d):
                        A[rp, o.row(0, 3, fv[bb])] += sign * p.C_M / p.F * wq * N[a] * N[bb]
                        A[rp, o.row(1, 3, fv[bb])] -= sign * p.C_M / p.F * wq * N[a] * N[bb]
    return A, b


def test_vectorised_assembly_equals_loop_assembly():
    o = KNPEMIOracle(unit_square(4), OracleParams(), MODELS_TEST)
    rng = np.random.default_rng(0)
    for s in range(2):
        o.c[s] *= 1 + 0.05 * rng.random(o.c[s].shape)
    o.phi_m += 0.003 * rng.standard_normal(o.phi_m.shape)
    o.gates *= 1 + 0.1 * rng.random(o.gates.shape)
    A, b = o.assemble(3 * o.p.dt)
    Al, bl = _loop_assemble(o, 3 * o.p.dt)
    scale = np.abs(Al).max(axis=1, keepdims=True)
    assert np.abs(A.toarray() - Al).max() / np.abs(Al).max() < 1e-13
    assert (np.abs(A.toarray() - Al) / scale).max() < 1e-11
    assert np.abs(b - bl).max() / np.abs(bl).max() < 1e-12


@pytest.mark.parametrize("mesh", [unit_square(8), unit_cube(4)])
def test_nullspace_and_pattern(mesh):
    o = KNPEMIOracle(mesh, OracleParams(), MODELS_TEST)
    A, b = o.assemble(o.p.dt)
    ns = o.nullspace()
    assert np.abs(A @ ns).max() < 1e-22 and np.abs(A.T @ ns).max() < 1e-22          # KNPEMIx_solver.py:327
    assert abs(ns @ b) < 1e-13 * np.abs(b).sum()
    assert A.has_sorted_indices and (np.diff(A.indptr) > 0).all()
    P = o.assemble_P()
    blk = np.searchsorted(np.cumsum([o.ns[0]] * 4 + [o.ns[1]] * 4), np.arange(o.n), side="right")
    C = P.tocoo()
    assert (blk[C.row] == blk[C.col]).all()                                          # block diagonal


def test_gate_update_closed_form():
    o = KNPEMIOracle(unit_square(4), OracleParams(), MODELS_TEST)
    o.phi_m[:] = np.linspace(-0.09, 0.03, o.phi_m.size)
    g0 = o.gates.copy()
    o.gate_update()
    V = 1000 * (o.phi_m - o.p.phi_rest)
    an = 0.01e3 * (10 - V) / (np.exp((10 - V) / 10) - 1); bn = 0.125e3 * np.exp(-V / 80)
    yinf, tau = an / (an + bn), 1 / (an + bn)
    np.testing.assert_allclose(o.gates[0], yinf + (g0[0] - yinf) * np.exp(-o.p.dt / tau), rtol=1e-12)


def test_charge_conservation_row_operation_and_schur_pc():
    """The identity the product's Schur preconditioner rests on: (phi row) - sum_k z_k (ion row k) = -sum_k z_k M c_k,
    i.e. L A has a vanishing (phi, phi) block and a static mass-matrix (phi, c) block; and GMRES preconditioned with
    oracle/amg.py::SchurPC reaches the sparse-LU solution of a transient step in a few dozen iterations."""
    import scipy.sparse as sp
    from oracle.amg import SchurPC
    o = KNPEMIOracle(unit_square(16), OracleParams(), MODELS_TEST)
    rng = np.random.default_rng(2)
    for s in range(2):
        o.c[s] *= 1 + 0.05 * rng.random(o.c[s].shape)
    o.phi[0] += 0.004 * rng.standard_normal(o.phi[0].shape)
    o.phi_m = o.phi[0] - o.phi[1]
    pc = SchurPC(o)
    A, b = o.assemble(o.p.dt)
    n = o.n
    rows, cols, vals = [], [], []
    for s in range(2):
        q = np.arange(o.ns[s])
        for k in range(3):
            rows.append(o.base[s] + 3 * o.ns[s] + q); cols.append(o.base[s] + k * o.ns[s] + q)
            vals.append(np.full(o.ns[s], -o.p.z[k]))
    L = sp.identity(n) + sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))
    LA = (L @ A).tocsr()
    App = A[pc.ip][:, pc.ip]
    assert abs(LA[pc.ip][:, pc.ip]).max() < 1e-12 * abs(App).max()
    Bpc = LA[pc.ip][:, pc.ic].tocsr()
    Mz = sp.bmat([[sp.hstack([-o.p.z[k] * pc.M[0] for k in range(3)]), None],
                  [None, sp.hstack([-o.p.z[k] * pc.M[1] for k in range(3)])]]).tocsr()
    assert abs(Bpc - Mz).max() < 1e-9 * abs(Mz).max()       # cancellation against dt D_k K (1e5 larger) costs ~5 digits
    ns = o.nullspace()
    b = b - ns * (ns @ b)
    x, its = o.solve_gmres(A, b, o.pack(), ns, pc, 1e-10)
    xd = o.solve_direct(A, b, ns)
    assert its < 60
    for s in range(2):
        for f in range(3):
            sl = slice(o.base[s] + f * o.ns[s], o.base[s] + (f + 1) * o.ns[s])
            assert np.linalg.norm(x[sl] - xd[sl]) <= 1e-8 * np.linalg.norm(xd[sl])


def test_conservation_functionals_closed_form():
    """int 1, int u and the membrane area on the C1 fixture: the intracellular square [0.25, 0.75]^2 um^2."""
    o = KNPEMIOracle(unit_square(32), OracleParams(), MODELS_TEST)
    L = 1e-6
    assert abs(o.integral(o.c[0][0], 1, power=0) - 0.25 * L * L) < 1e-12 * L * L
    assert abs(o.integral(o.c[0][0], [1, 2], power=0) - L * L) < 1e-12 * L * L
    assert abs(o.integral(o.c[0][0], 1) - 12.0 * 0.25 * L * L) < 1e-10 * L * L          # Na_i = 12 mM
    x = o.mesh.x[:, 0]
    assert abs(o.integral(x, [1, 2]) - 0.5 * L ** 3) < 1e-12 * L ** 3                   # int x dx, exact for P1
    assert abs(o.membrane_area(4) - 4 * 0.5 * L) < 1e-12 * L


def test_pcg_restatement_solves_spd_system():
    """oracle PCG (checker of the device CG loop) against a dense solve, with and without a Jacobi preconditioner."""
    import scipy.sparse as sp
    from oracle.knpemi import KNPEMIOracle
    n = 200
    rng = np.random.default_rng(0)
    T = sp.diags([-1.0, 2.5, -1.0], [-1, 0, 1], shape=(n, n)).tocsr()
    A = (T + sp.diags(rng.random(n))).tocsr()
    b = rng.standard_normal(n)
    x_exact = np.linalg.solve(A.toarray(), b)
    for Binv in (lambda v: v, lambda v: v / A.diagonal()):
        x, its = KNPEMIOracle.solve_pcg(A, b, np.zeros(n), Binv, 1e-12)
        assert 0 < its < n
        assert np.abs(x - x_exact).max() <= 1e-10 * np.abs(x_exact).max()


@pytest.mark.parametrize("mode", ["dirichlet", "pinned"])
def test_essential_conditions_equal_the_reduced_system(mode):
    """oracle.apply_bcs restates assemble_matrix_block / assemble_vector_block with bcs (KNPEMIx_solver.py:113-116: rows and
    columns zeroed, unit diagonal, lifted right-hand side).  Independent guard: the solution equals the one of the system
    with the constrained unknowns eliminated by hand, the constrained values come out exactly, the matrix is regular (no
    nullspace is attached, :380,415), and GMRES with the Schur preconditioner -- Dirichlet dofs cut out of its hierarchies --
    reaches the direct solution."""
    import scipy.sparse.linalg as spla
    from oracle.amg import SchurPC
    om = unit_square(16)
    x, y = om.x[:, 0] / om.x[:, 0].max(), om.x[:, 1] / om.x[:, 1].max()
    bv = np.flatnonzero((x == 0) | (x == 1) | (y == 0) | (y == 1))
    kw = dict(dirichlet_bcs=True, boundary_verts=tuple(bv)) if mode == "dirichlet" else dict(pin_vertex=0)
    models = [("NeuronalCT", None), ("HH", None), ("ATP", None)]
    o, free_o = KNPEMIOracle(om, OracleParams(**kw), models), KNPEMIOracle(om, OracleParams(), models)
    idx, g = o.bc_dofs()
    assert idx.size == (8 * 0 + 4 * bv.size if mode == "dirichlet" else 1)      # boundary vertices are extracellular only
    A, b = o.assemble(o.p.dt)
    A0, b0 = free_o.assemble(o.p.dt)
    assert np.array_equal(A.indptr, A0.indptr) and np.array_equal(A.indices, A0.indices)      # entries stay in the pattern
    assert not o.nullspace().any()
    free = np.setdiff1d(np.arange(o.n), idx)
    ref = np.zeros(o.n)
    ref[idx] = g
    ref[free] = spla.splu(A0[free][:, free].tocsc()).solve(b0[free] - A0[free][:, idx] @ g)
    sol = o.solve_direct(A, b, o.nullspace())
    assert np.array_equal(sol[idx], g)
    assert np.abs(sol - ref).max() <= 1e-10 * np.abs(ref).max()
    oa, ob = KNPEMIOracle(om, OracleParams(**kw), models), KNPEMIOracle(om, OracleParams(**kw), models)
    pc = SchurPC(ob)
    xb = ob.pack()
    for i in range(2):
        _, _, xa, _ = oa.step("direct")
        _, _, xb, its = ob.step("gmres", pc, 1e-12, xb, first=(i == 0))
        assert its <= (12 if mode == "dirichlet" else 22)
        assert np.array_equal(xb[idx], g)
        assert np.abs(xa - xb).max() <= 1e-9 * np.abs(xa).max()
