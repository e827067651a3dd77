"""numpy/scipy restatement of CGx's per-timestep KNP-EMI path (TEST INFRASTRUCTURE).

Follows, function by function:

* ``assemble``            <- SolverKNPEMI.assemble  (KNPEMIx_solver.py:104-116) applied to the
                             forms of ProblemKNPEMI.setup_variational_form (KNPEMIx_problem.py:454-655)
* ``assemble_P``          <- setup_preconditioner / assemble_preconditioner
                             (KNPEMIx_problem.py:657-744, KNPEMIx_solver.py:118-135)
* ``channel_currents``    <- IonicModel._eval family + HodgkinHuxley._add_stimulus
                             (KNPEMIx_ionic_model.py:89-91,140-222,246-298,317-369,389-424,487-603)
* ``gate_update``         <- HodgkinHuxley.update_gating_variables (KNPEMIx_ionic_model.py:605-671)
* ``solve_direct``        <- KSP preonly + MUMPS with attached nullspace (KNPEMIx_solver.py:167-172,297-335,378-383,435)
* ``solve_gmres``         <- KSP gmres, left PC, preconditioned norm (KNPEMIx_solver.py:212-214,276-280,386-389,435)
* ``step`` / ``run``      <- SolverKNPEMI.solve time loop (KNPEMIx_solver.py:365-468)
* ``l2_norm``             <- tests/KNPEMI/electric_potential_norms_direct_solver.py:45-51

Unknown ordering (serial multiphenicsx convention): field-major blocks
[Na_i K_i Cl_i phi_i | Na_e K_e Cl_e phi_e]; inside a block ascending vertex id
restricted to the subdomain's vertex set.  Sparsity convention: *minimal*
pattern (all dof pairs of every cell for the dx blocks; facet couplings only
between the membrane facet's own vertices); explicit zeros are kept.
"""
from dataclasses import dataclass, field
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .quadrature import facet_rule

ION_NAMES = ("Na", "K", "Cl")


@dataclass
class OracleParams:
    dt: float = 2.5e-5
    T: float = 300.0
    F: float = 96485.0
    R: float = 8.314
    C_M: float = 0.02
    z: tuple = (1.0, 1.0, -1.0)
    D: tuple = (1.33e-9, 1.96e-9, 2.03e-9)
    phi_rest: float = -0.065
    # conductances (mixed_dim_problem.py:311-332)
    g_Na_bar: float = 1200.0
    g_K_bar: float = 360.0
    g_leak: tuple = (0.3, 0.1, 0.25)
    g_leak_g: tuple = (1.0, 16.96, 2.0)
    g_syn_bar: float = 1e-9
    a_syn: float = 5e-4
    T_stim: float = 1.0
    scale_stimulus: bool = True
    # initial conditions (configs/tests/*.yaml:27-38)
    phi_m_init: float = -0.070
    c_i_init: tuple = (12.0, 130.0, 5.0)
    c_e_init: tuple = (140.0, 4.0, 125.0)
    n_init: float = 0.276
    m_init: float = 0.0379
    h_init: float = 0.688
    # glia (optional)
    glia_tags: tuple = ()
    phi_m_g_init: float = -0.085
    c_i_g_init: tuple = (15.0, 100.0, 5.0)
    # tags
    intra_tags: tuple = (1,)
    extra_tag: int = 2
    membrane_tags: tuple = (4,)
    stimulus_tags: tuple = (4,)
    stimulus_region: tuple = None      # (direction, lo, hi), a tuple of such triples (`multiple` directions), or None
    ode_substeps: int = 25
    rush_larsen: bool = True
    # "ion_injection": K+ / Cl- source in the extracellular space around the mesh centre (KNPEMIx_problem.py:200-218)
    source_terms: str = None
    injection_current: float = 5e-9
    # essential boundary conditions (KNPEMIx_problem.py:96-198): dirichlet_bcs = every field at its initial value on the
    # vertices of the facets tagged boundary_tags (given here as the vertex list); pin_vertex = phi_e = 0 at one vertex
    dirichlet_bcs: bool = False
    boundary_verts: tuple = ()
    pin_vertex: int = None

    @property
    def psi(self):
        return self.R * self.T / self.F


class KNPEMIOracle:
    """State + operators for one problem instance.

    ``models`` is an ordered list of (name, tags) with name in
    {"NeuronalCT","HH","ATP","Passive","GlialCT","KirNa"}; tags=None -> all membrane tags
    (IonicModel.__init__, KNPEMIx_ionic_model.py:13-34).
    """

    def __init__(self, mesh, params: OracleParams, models):
        self.mesh, self.p = mesh, params
        p = params
        self.models = [(nm, tuple(p.membrane_tags) if tg is None else tuple(tg)) for nm, tg in models]
        d = mesh.gdim
        nv = mesh.x.shape[0]
        is_in = np.isin(mesh.cell_tags, np.asarray(p.intra_tags))
        is_ex = mesh.cell_tags == p.extra_tag
        self.cells_s = [mesh.cells[is_in], mesh.cells[is_ex]]
        self.celltags_s = [mesh.cell_tags[is_in], mesh.cell_tags[is_ex]]
        # restricted dof sets (DofMapRestriction, KNPEMIx_problem.py:85-89)
        self.S = [np.unique(c.ravel()) for c in self.cells_s]
        self.r = []
        for S in self.S:
            r = np.full(nv, -1, np.int64)
            r[S] = np.arange(S.size)
            self.r.append(r)
        self.ns = [S.size for S in self.S]
        self.n = 4 * (self.ns[0] + self.ns[1])
        self.base = [0, 4 * self.ns[0]]
        # geometry of cells per subdomain
        self.geo = [self._cell_geometry(c) for c in self.cells_s]
        # membrane facets
        fv = mesh.mf_verts
        xf = mesh.x[fv]                                        # (nf, d, gdim)
        if d == 2:
            self.farea = np.linalg.norm(xf[:, 1] - xf[:, 0], axis=1)
        else:
            self.farea = 0.5 * np.linalg.norm(np.cross(xf[:, 1] - xf[:, 0], xf[:, 2] - xf[:, 0]), axis=1)
        self.qb, self.qw = facet_rule(d)
        self.qN = self._trace_basis(self.qb)                   # trace basis at the facet quadrature points (P1: the barycentrics)
        self.mverts = np.unique(fv.ravel())
        # fields on the whole mesh, like the reference's wh / phi_m_prev / n,m,h
        self.c = [np.zeros((3, nv)), np.zeros((3, nv))]
        self.phi = [np.zeros(nv), np.zeros(nv)]
        self.phi_m = np.zeros(nv)
        self.gates = np.zeros((3, nv))
        self.t = 0.0
        self.set_initial_conditions()
        self._pattern = None
        self.f_e = [np.zeros(nv) for _ in range(3)]             # source functions f_e of the ions (ion['f_e'], :771-800)
        if p.source_terms == "ion_injection":
            self.setup_ion_injection()

    # ------------------------------------------------------------------ setup
    def row(self, s, f, verts):
        return self.base[s] + f * self.ns[s] + self.r[s][verts]

    def bc_dofs(self):
        """ProblemKNPEMI.setup_boundary_conditions (KNPEMIx_problem.py:96-198), non-MMS branches: with dirichlet_bcs every
        field takes its initial value (k_init, phi_m_init inside, 0 outside) on the boundary dofs of its restriction
        (:139-161); otherwise pin_ecs_potential pins phi_e = 0 at one vertex off the membrane (:163-194).
        Returns (unknown indices ascending, values)."""
        p = self.p
        idx, val = [], []
        if p.dirichlet_bcs:
            bv = np.unique(np.asarray(p.boundary_verts, np.int64))
            for s in range(2):
                v = bv[self.r[s][bv] >= 0]
                init = p.c_i_init if s == 0 else p.c_e_init
                for k in range(3):
                    idx.append(self.row(s, k, v))
                    val.append(np.full(v.size, init[k]))
                idx.append(self.row(s, 3, v))
                val.append(np.full(v.size, p.phi_m_init if s == 0 else 0.0))
        elif p.pin_vertex is not None:
            assert self.r[1][p.pin_vertex] >= 0 and p.pin_vertex not in self.mverts
            idx.append(np.array([self.row(1, 3, p.pin_vertex)]))
            val.append(np.zeros(1))
        if not idx:
            return np.zeros(0, np.int64), np.zeros(0)
        idx, val = np.concatenate(idx), np.concatenate(val)
        o = np.argsort(idx)
        return idx[o], val[o]

    def apply_bcs(self, A, b=None, diag=1.0):
        """assemble_matrix_block(A, a, bcs) + assemble_vector_block(b, L, a, bcs) (KNPEMIx_solver.py:113-116): rows and columns
        of the constrained dofs are zeroed (the entries stay in the pattern), the diagonal is set to `diag`; the right-hand
        side is lifted with the unconstrained columns, b -= A[:, bc] g, and set to g on the constrained rows."""
        idx, g = self.bc_dofs()
        if idx.size == 0:
            return A, b
        A = A.tocsr().copy()
        isbc = np.zeros(self.n, bool)
        isbc[idx] = True
        if b is not None:
            gf = np.zeros(self.n)
            gf[idx] = g
            b = b - A @ gf
            b[idx] = g
        rows = np.repeat(np.arange(self.n), np.diff(A.indptr))
        A.data[isbc[rows] | isbc[A.indices]] = 0.0
        A.data[isbc[rows] & (rows == A.indices)] = diag
        return A, b

    def _trace_basis(self, qb):
        """Basis functions of the facet's dofs at barycentric points qb (nq, d): P1 -> the barycentric coordinates."""
        return qb

    def lumped_mass(self, M):
        """Diagonal stand-in for a mass matrix (M_sigma of the Schur preconditioner): P1 -> row sums."""
        return np.asarray(M.sum(axis=1)).ravel()

    def _cK(self, s, k, coef):
        """Element matrices coef * int c_k grad N_a . grad N_b dx of subdomain s (the c_k-weighted stiffness terms of
        KNPEMIx_problem.py:599-610): for P1 the cell mean of c_k times the stiffness matrix, exactly."""
        cbar = self.c[s][k][self.cells_s[s]].mean(axis=1)
        return coef * cbar[:, None, None] * self.geo[s]["K"]

    def _cell_geometry(self, cells):
        d = self.mesh.gdim
        x = self.mesh.x[cells[:, :d + 1]]                       # (nc, d+1, gdim): the cell's vertices come first
        J = np.transpose(x[:, 1:] - x[:, :1], (0, 2, 1))        # columns = edge vectors
        det = np.linalg.det(J)
        vol = np.abs(det) / (2.0 if d == 2 else 6.0)
        Jinv = np.linalg.inv(J)                                 # rows = grads of N_1..N_d
        g = np.concatenate([-Jinv.sum(1, keepdims=True), Jinv], 1)   # (nc, d+1, gdim)
        K = vol[:, None, None] * np.einsum("cai,cbi->cab", g, g)
        M = vol[:, None, None] / ((d + 1) * (d + 2)) * (1.0 + np.eye(d + 1))[None]
        return dict(vol=vol, K=K, M=M, g=g)

    def set_initial_conditions(self):
        """ProblemKNPEMI.set_initial_conditions (KNPEMIx_problem.py:326-353,386-447) and
        HodgkinHuxley._init (KNPEMIx_ionic_model.py:466-485)."""
        p = self.p
        if not p.glia_tags:
            self.phi_m[:] = p.phi_m_init
            self.phi[0][:] = p.phi_m_init
            self.phi[1][:] = 0.0
            for k in range(3):
                self.c[0][k, :] = p.c_i_init[k]
                self.c[1][k, :] = p.c_e_init[k]
        else:
            tags = self.mesh.cell_tags
            glia = np.isin(tags, np.asarray(p.glia_tags))
            neur = np.isin(tags, np.asarray(p.intra_tags)) & ~glia
            nd = np.unique(self.mesh.cells[neur].ravel())
            gd = np.unique(self.mesh.cells[glia].ravel())
            self.phi_m[nd] = p.phi_m_init
            self.phi_m[gd] = p.phi_m_g_init
            self.phi[0][nd] = p.phi_m_init
            self.phi[0][gd] = p.phi_m_g_init
            self.phi[1][:] = 0.0
            for k in range(3):
                self.c[0][k, nd] = p.c_i_init[k]
                self.c[0][k, gd] = p.c_i_g_init[k]
                self.c[1][k, :] = p.c_e_init[k]
        self.gates[0, :] = p.n_init
        self.gates[1, :] = p.m_init
        self.gates[2, :] = p.h_init
        self.t = 0.0

    # ------------------------------------------------------------ membrane physics
    def gate_update(self):
        """HodgkinHuxley.update_gating_variables (KNPEMIx_ionic_model.py:605-671)."""
        p = self.p
        dt_ode = p.dt / p.ode_substeps
        V = 1000.0 * (self.phi_m - p.phi_rest)
        with np.errstate(divide="ignore", invalid="ignore"):
            a_n = 0.01e3 * (10.0 - V) / (np.exp((10.0 - V) / 10.0) - 1.0)
            b_n = 0.125e3 * np.exp(-V / 80.0)
            a_m = 0.1e3 * (25.0 - V) / (np.exp((25.0 - V) / 10.0) - 1.0)
            b_m = 4.0e3 * np.exp(-V / 18.0)
            a_h = 0.07e3 * np.exp(-V / 20.0)
            b_h = 1.0e3 / (np.exp((30.0 - V) / 10.0) + 1.0)
        al = (a_n, a_m, a_h)
        be = (b_n, b_m, b_h)
        if p.rush_larsen:
            for j in range(3):
                tau = 1.0 / (al[j] + be[j])
                yinf = al[j] * tau
                yexp = np.exp(-dt_ode / tau)
                y = self.gates[j]
                for _ in range(p.ode_substeps):
                    y = yinf + (y - yinf) * yexp
                self.gates[j] = y
        else:
            for j in range(3):
                aa, bb = al[j] * dt_ode, be[j] * dt_ode
                y = self.gates[j]
                for _ in range(p.ode_substeps):
                    y = y + (aa * (1.0 - y) - bb * y)
                self.gates[j] = y

    def stimulus_mask(self, xq):
        p = self.p
        if p.stimulus_region is None:
            return np.ones(xq.shape[:-1])
        regions = p.stimulus_region if isinstance(p.stimulus_region[0], (tuple, list)) else [p.stimulus_region]
        mask = np.ones(xq.shape[:-1])
        for direction, lo, hi in regions:                  # product of the per-direction masks (ionic_model.py:574-587)
            c = xq[..., direction]
            mask = mask * ((c > lo) & (c < hi)).astype(float)
        return mask

    def stimulus_area(self):
        """p.stimulus_area = assemble(mask * dS(stimulus_tags)) (KNPEMIx_ionic_model.py:591-601)."""
        m = self.mesh
        sel = np.isin(m.mf_tags, np.asarray(self.p.stimulus_tags))
        xq = np.einsum("qa,fai->fqi", self.qb, m.x[m.mf_verts[sel][:, :m.gdim]])
        return float(np.sum(self.farea[sel, None] * self.qw[None, :] * self.stimulus_mask(xq)))

    def channel_currents(self, tag, ci, ce, phim, gq, xq, t_mod):
        """I_ch,k (k = Na,K,Cl) at quadrature points of facets carrying membrane tag `tag`.
        All inputs are already interpolated to the quadrature points (fields first, then
        the nonlinear functions: that is what FFCx generates for these UFL expressions)."""
        p = self.p
        psi = p.psi
        E = [(psi / p.z[k]) * np.log(ce[k] / ci[k]) for k in range(3)]     # KNPEMIx_problem.py:516
        I = [np.zeros_like(phim) for _ in range(3)]
        for name, tags in self.models:
            if tag not in tags:
                continue
            if name == "Passive":                                         # ionic_model.py:89-91
                for k in range(3):
                    I[k] = I[k] + phim
            elif name == "NeuronalCT":                                    # ionic_model.py:342-369
                I_KCC2 = 0.0068 * np.log((ci[1] * ci[2]) / (ce[1] * ce[2]))
                I_NKCC1 = 0.0023 * 0.0 * np.log((ce[0] * ce[1] * ce[2] ** 2) / (ci[0] * ci[1] * ci[2] ** 2))
                I[0] = I[0] + (-I_NKCC1)
                I[1] = I[1] + (-I_NKCC1 + I_KCC2)
                I[2] = I[2] + (I_NKCC1 - I_KCC2)
            elif name == "GlialCT":                                       # ionic_model.py:239-298
                I_KCC1 = (7e-2 * psi) * np.log((ci[1] * ci[2]) / (ce[1] * ce[2]))
                I_NKCC1 = (2e-2 * psi) * 0.0 * np.log((ce[0] * ce[1] * ce[2] ** 2) / (ci[0] * ci[1] * ci[2] ** 2))
                I[0] = I[0] + (-I_NKCC1)
                I[1] = I[1] + (-I_NKCC1 + I_KCC1)
                I[2] = I[2] + (2 * I_NKCC1 - I_KCC1)
            elif name == "ATP":                                           # ionic_model.py:385-422
                par1 = 1.0 + 1.5 / ce[1]
                par2 = 1.0 + 10.0 / ci[0]
                I_ATP = 0.25 / (par1 ** 2 * par2 ** 3)
                I[0] = I[0] + 3 * I_ATP
                I[1] = I[1] + (-2 * I_ATP)
                I[2] = I[2] + 0.0
            elif name == "HH":                                            # ionic_model.py:487-515
                n, m, h = gq
                g = [p.g_leak[0] + p.g_Na_bar * m ** 3 * h,
                     p.g_leak[1] + p.g_K_bar * n ** 4,
                     p.g_leak[2] + 0.0 * n]
                Ik = [g[k] * (phim - E[k]) for k in range(3)]
                if tag in p.stimulus_tags:                                # KNPEMIx_problem.py:531-549
                    stim = self.stimulus_mask(xq) * p.g_syn_bar * np.exp(-t_mod / p.a_syn) * (phim - E[0])
                    if p.scale_stimulus:
                        stim = stim * (1.0 / self._stim_area)
                    Ik[0] = Ik[0] + stim
                for k in range(3):
                    I[k] = I[k] + Ik[k]
            elif name == "KirNa":                                         # ionic_model.py:117-222
                E_K_init = psi * np.log(p.c_e_init[1] / p.c_i_g_init[1])
                rho = 1.1 * 1.12e-6
                pump = (1.0 / (1.0 + (10.0 / ci[0]) ** 1.5)) * (1.0 / (1.0 + 1.5 / ce[1])) * rho
                A_ = 1 + np.exp(0.433)
                B_ = 1 + np.exp(-(0.1186 + E_K_init) / 0.0441)
                C_ = 1 + np.exp(((phim - E[1]) + 0.0185) / 0.0425)
                D_ = 1 + np.exp(-(0.1186 + phim) / 0.0441)
                f_kir = np.sqrt(ce[1] / p.c_e_init[1]) * A_ * B_ / (C_ * D_)
                I[0] = I[0] + (1.0 * p.g_leak_g[0] * (phim - E[0]) + 3 * p.z[0] * p.F * pump)
                I[1] = I[1] + (f_kir * p.g_leak_g[1] * (phim - E[1]) + (-2 * p.z[1] * p.F * pump))
                I[2] = I[2] + (1.0 * p.g_leak_g[2] * (phim - E[2]) + 0.0)
            else:
                raise ValueError(name)
        return I

    # ------------------------------------------------------------------ assembly
    def _facet_quadrature_fields(self):
        m = self.mesh
        fv = m.mf_verts
        qb = self.qN
        ci = [np.einsum("qa,fa->fq", qb, self.c[0][k][fv]) for k in range(3)]
        ce = [np.einsum("qa,fa->fq", qb, self.c[1][k][fv]) for k in range(3)]
        phim = np.einsum("qa,fa->fq", qb, self.phi_m[fv])
        gq = [np.einsum("qa,fa->fq", qb, self.gates[j][fv]) for j in range(3)]
        xq = np.einsum("qa,fai->fqi", self.qb, m.x[fv[:, :m.gdim]])
        return ci, ce, phim, gq, xq

    def facet_tensors(self, t_mod):
        """Per-facet element tensors: GA[s][k] (nf,d,d) = G[alpha_{k,s}], G1 (nf,d,d),
        bc[s][k] (nf,d) and bphi (nf,d) as in Appendix A of SURVEY.md / KNPEMIx_problem.py:594-642."""
        p, m = self.p, self.mesh
        d = m.gdim
        nf = m.mf_verts.shape[0]
        ci, ce, phim, gq, xq = self._facet_quadrature_fields()
        wq = self.farea[:, None] * self.qw[None, :]                      # (nf, nq)
        NN = np.einsum("qa,qb->qab", self.qN, self.qN)
        cs = [ci, ce]
        alpha = []
        for s in range(2):
            den = sum(p.D[j] * p.z[j] ** 2 * cs[s][j] for j in range(3))
            alpha.append([p.D[k] * p.z[k] ** 2 * cs[s][k] / den for k in range(3)])
        I = [np.zeros((nf, self.qw.size)) for _ in range(3)]
        for tag in np.unique(m.mf_tags):
            sel = m.mf_tags == tag
            Ik = self.channel_currents(int(tag), [a[sel] for a in ci], [a[sel] for a in ce],
                                       phim[sel], [a[sel] for a in gq], xq[sel], t_mod)
            for k in range(3):
                I[k][sel] = Ik[k]
        Itot = (I[0] + I[1]) + I[2]
        GA = [[np.einsum("fq,fq,qab->fab", wq, alpha[s][k], NN) for k in range(3)] for s in range(2)]
        G1 = np.einsum("fq,qab->fab", wq, NN)
        bc = [[np.einsum("fq,fq,qa->fa", wq, (p.dt * I[k] - alpha[s][k] * p.C_M * phim), self.qN) / (p.F * p.z[k])
               for k in range(3)] for s in range(2)]
        bphi = np.einsum("fq,fq,qa->fa", wq, (p.dt * Itot - p.C_M * phim), self.qN) / p.F
        return GA, G1, bc, bphi

    def setup_ion_injection(self):
        """Injection site = cells with every vertex inside the cube of half-width (x_max - x_min)/10 around the centre of
        the bounding box (utils/mixed_dim_problem.py:496-540,806-811); f_e of K and Cl = I / (F vol) on all vertices of those
        cells (KNPEMIx_problem.py:200-218).  The form integrates f_e over the extracellular cells only (:614)."""
        m, p = self.mesh, self.p
        lo, hi = m.x.min(axis=0), m.x.max(axis=0)
        centre, delta, tol = 0.5 * (lo + hi), (hi[0] - lo[0]) / 10.0, 1e-14
        inside = np.all((m.x >= centre - delta - tol) & (m.x <= centre + delta + tol), axis=1)
        self.injection_cells = np.flatnonzero(inside[m.cells].all(axis=1))
        vol = self._cell_geometry(m.cells[self.injection_cells])["vol"]
        self.injection_volume = float(vol.sum())
        src = (p.injection_current / p.F) / self.injection_volume
        verts = np.unique(m.cells[self.injection_cells].ravel())
        self.f_e[1][verts] = src
        self.f_e[2][verts] = src

    def assemble(self, t):
        """Returns (A csr with sorted indices and explicit zeros kept, b)."""
        p, m = self.p, self.mesh
        d = m.gdim
        psi = p.psi
        self._stim_area = self.stimulus_area() if p.scale_stimulus else 1.0
        t_mod = np.mod(t + 1e-12, p.T_stim)                               # ionic_model.py:673-674
        rows, cols, vals = [], [], []
        b = np.zeros(self.n)

        def add(R, C, V):
            rows.append(np.broadcast_to(R, V.shape).ravel())
            cols.append(np.broadcast_to(C, V.shape).ravel())
            vals.append(V.ravel())

        for s in range(2):
            cells = self.cells_s[s]
            K, M = self.geo[s]["K"], self.geo[s]["M"]
            cv = [self.c[s][k][cells] for k in range(3)]                  # (nc, dofs per cell)
            Rphi = self.row(s, 3, cells)
            Kphi = np.zeros_like(K)
            for k in range(3):
                Rk = self.row(s, k, cells)
                add(Rk[:, :, None], Rk[:, None, :], M + p.dt * p.D[k] * K)
                add(Rk[:, :, None], Rphi[:, None, :], self._cK(s, k, p.dt * p.D[k] * p.z[k] / psi))
                add(Rphi[:, :, None], Rk[:, None, :], (p.dt * p.z[k] * p.D[k]) * K)
                Kphi = Kphi + self._cK(s, k, p.dt * p.D[k] * p.z[k] ** 2 / psi)
                np.add.at(b, Rk.ravel(), np.einsum("cab,cb->ca", M, cv[k]).ravel())
                if s == 1 and self.f_e[k].any():                          # L += dt f_e v dx_e  (KNPEMIx_problem.py:614)
                    np.add.at(b, Rk.ravel(), p.dt * np.einsum("cab,cb->ca", M, self.f_e[k][cells]).ravel())
            add(Rphi[:, :, None], Rphi[:, None, :], Kphi)

        GA, G1, bc, bphi = self.facet_tensors(t_mod)
        fv = m.mf_verts
        Rp = [self.row(0, 3, fv), self.row(1, 3, fv)]
        sign = [1.0, -1.0]
        for s in range(2):
            for k in range(3):
                Rk = self.row(s, k, fv)
                coef = p.C_M / (p.F * p.z[k])
                add(Rk[:, :, None], Rp[0][:, None, :], sign[s] * coef * GA[s][k])
                add(Rk[:, :, None], Rp[1][:, None, :], -sign[s] * coef * GA[s][k])
                np.add.at(b, Rk.ravel(), (-sign[s] * bc[s][k]).ravel())
            add(Rp[s][:, :, None], Rp[0][:, None, :], sign[s] * (p.C_M / p.F) * G1)
            add(Rp[s][:, :, None], Rp[1][:, None, :], -sign[s] * (p.C_M / p.F) * G1)
            np.add.at(b, Rp[s].ravel(), (-sign[s] * bphi).ravel())

        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(self.n, self.n)).tocsr()
        A.sum_duplicates()
        A.sort_indices()
        return self.apply_bcs(A, b)

    def assemble_P(self, membrane_sign=-1.0, D_scale=1.0, bc_diag=1.0):
        """Block-Jacobi preconditioner form (KNPEMIx_problem.py:717-738), from the current fields.
        membrane_sign=+1 / D_scale=0 give the two auxiliary matrices of the product's own Schur preconditioner
        (oracle/amg.py::SchurPC): the phi blocks with the sign the membrane term has in `a`, and the mass matrices."""
        p, m = self.p, self.mesh
        psi = p.psi
        rows, cols, vals = [], [], []

        def add(R, C, V):
            rows.append(np.broadcast_to(R, V.shape).ravel())
            cols.append(np.broadcast_to(C, V.shape).ravel())
            vals.append(V.ravel())

        for s in range(2):
            cells = self.cells_s[s]
            K, M = self.geo[s]["K"], self.geo[s]["M"]
            Rphi = self.row(s, 3, cells)
            Kphi = np.zeros_like(K)
            for k in range(3):
                Rk = self.row(s, k, cells)
                add(Rk[:, :, None], Rk[:, None, :], M + (D_scale * p.dt * p.D[k]) * K)
                Kphi = Kphi + self._cK(s, k, D_scale * p.dt * p.D[k] * p.z[k] ** 2 / psi)
            add(Rphi[:, :, None], Rphi[:, None, :], Kphi)
        NN = np.einsum("qa,qb->qab", self.qN, self.qN)
        G1 = np.einsum("fq,qab->fab", self.farea[:, None] * self.qw[None, :], NN)
        fv = m.mf_verts
        for s in range(2):
            Rp = self.row(s, 3, fv)
            add(Rp[:, :, None], Rp[:, None, :], membrane_sign * (p.C_M / p.F) * G1)
        P = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(self.n, self.n)).tocsr()
        P.sum_duplicates()
        P.sort_indices()
        return self.apply_bcs(P, None, bc_diag)[0]      # assemble_matrix_block(p.P, bcs=p.bcs) (KNPEMIx_solver.py:125-126)

    # ------------------------------------------------------------------- vectors
    def nullspace(self):
        """create_and_set_nullspace (KNPEMIx_solver.py:297-335): normalised indicator of phi rows."""
        ns = np.zeros(self.n)
        if self.p.dirichlet_bcs or self.p.pin_vertex is not None:
            return ns                       # no nullspace is attached with essential conditions (KNPEMIx_solver.py:380,415)
        ns[self.base[0] + 3 * self.ns[0]: self.base[0] + 4 * self.ns[0]] = 1.0
        ns[self.base[1] + 3 * self.ns[1]: self.base[1] + 4 * self.ns[1]] = 1.0
        return ns / np.linalg.norm(ns)

    def pack(self):
        """Restricted block vector from the full-mesh fields (BlockVecSubVectorWrapper, solver.py:203-209)."""
        x = np.empty(self.n)
        for s in range(2):
            for k in range(3):
                x[self.base[s] + k * self.ns[s]: self.base[s] + (k + 1) * self.ns[s]] = self.c[s][k][self.S[s]]
            x[self.base[s] + 3 * self.ns[s]: self.base[s] + 4 * self.ns[s]] = self.phi[s][self.S[s]]
        return x

    def unpack(self, x):
        """KNPEMIx_solver.py:451-468: restricted entries only; then phi_m = phi_i - phi_e on all dofs."""
        for s in range(2):
            for k in range(3):
                self.c[s][k][self.S[s]] = x[self.base[s] + k * self.ns[s]: self.base[s] + (k + 1) * self.ns[s]]
            self.phi[s][self.S[s]] = x[self.base[s] + 3 * self.ns[s]: self.base[s] + 4 * self.ns[s]]
        self.phi_m = self.phi[0] - self.phi[1]

    # -------------------------------------------------------------------- solves
    @staticmethod
    def equilibrate(A):
        """Symmetric-ish row/column scaling by sqrt|diag| (the raw system has cond_1 ~ 7e17)."""
        dgl = np.abs(A.diagonal())
        s = 1.0 / np.sqrt(dgl)
        return sp.diags(s) @ A @ sp.diags(s), s

    def solve_direct(self, A, b, ns, refine=2):
        """Bordered sparse LU [[A, ns],[ns^T, 0]] -> the solution with ns^T x = 0, which is what
        PREONLY+MUMPS(ICNTL24) followed by KSP's nullspace removal returns."""
        As, s = self.equilibrate(A)
        if not ns.any():                    # essential conditions: the matrix is regular, plain LU
            lu = spla.splu(As.tocsc())
            y = lu.solve(s * b)
            for _ in range(refine):
                y = y + lu.solve(s * b - As @ y)
            return s * y
        nss = ns / s
        nss = nss / np.linalg.norm(nss)
        n = self.n
        Kb = sp.bmat([[As, sp.csr_matrix(nss[:, None])], [sp.csr_matrix(nss[None, :]), None]]).tocsc()
        lu = spla.splu(Kb)
        rhs = np.concatenate([s * b, [0.0]])
        y = lu.solve(rhs)
        for _ in range(refine):
            y = y + lu.solve(rhs - Kb @ y)
        x = s * y[:n]
        return x - ns * (ns @ x)

    def solve_gmres(self, A, b, x0, ns, Pinv, rtol, restart=30, maxit=5000):
        """Left-preconditioned GMRES(restart) with classical Gram-Schmidt (refinement if needed),
        convergence ||B r|| <= rtol ||B b||, nullspace removed after every PC apply
        (PETSc KSP defaults).  Returns (x, iterations)."""
        def B(v):
            w = Pinv(v)
            return w - ns * (ns @ w)
        x = x0.copy()
        bnorm = np.linalg.norm(B(b))
        its = 0
        while True:
            r = B(b - A @ x)
            beta = np.linalg.norm(r)
            if beta <= rtol * bnorm or its >= maxit:
                return x, its
            V = np.zeros((restart + 1, self.n))
            H = np.zeros((restart + 1, restart))
            V[0] = r / beta
            g = np.zeros(restart + 1)
            g[0] = beta
            cs, sn = np.zeros(restart), np.zeros(restart)
            j_done = 0
            for j in range(restart):
                # Gram-Schmidt on d = (B A - I) v_j instead of w = B A v_j: same Krylov space, H(:, j) = e_j + V^T d, same
                # new direction, but no cancellation when B A is close to the identity (the product does the same,
                # csrc/solver.cu::gmres_solve)
                w = B(A @ V[j]) - V[j]
                h = V[: j + 1] @ w
                before = w @ w
                w = w - V[: j + 1].T @ h
                nrm2 = before - h @ h
                if not (nrm2 > 0.01 * before):         # second pass only after a cancellation by more than 10 (see the product)
                    h2 = V[: j + 1] @ w
                    w2 = w @ w
                    w = w - V[: j + 1].T @ h2
                    h = h + h2
                    nrm2 = w2 - h2 @ h2
                H[: j + 1, j] = h
                H[j, j] += 1.0
                H[j + 1, j] = np.sqrt(max(nrm2, 0.0))
                if H[j + 1, j] > 0:
                    V[j + 1] = w / H[j + 1, j]
                for i in range(j):
                    tmp = cs[i] * H[i, j] + sn[i] * H[i + 1, j]
                    H[i + 1, j] = -sn[i] * H[i, j] + cs[i] * H[i + 1, j]
                    H[i, j] = tmp
                den = np.hypot(H[j, j], H[j + 1, j])
                cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
                H[j, j] = den
                H[j + 1, j] = 0.0
                g[j + 1] = -sn[j] * g[j]
                g[j] = cs[j] * g[j]
                its += 1
                j_done = j + 1
                if abs(g[j + 1]) <= rtol * bnorm or its >= maxit:
                    break
            y = np.linalg.solve(np.triu(H[:j_done, :j_done]), g[:j_done])
            x = x + V[:j_done].T @ y
            if abs(g[j_done]) <= rtol * bnorm or its >= maxit:
                return x, its

    @staticmethod
    def solve_pcg(A, b, x0, Binv, rtol, maxit=5000):
        """Preconditioned conjugate gradients with PETSc's KSPCG conventions for the options the reference passes through
        (KNPEMIx_solver.py:212,276-280): preconditioned norm ||B r|| relative to ||B b||, nonzero initial guess.
        Returns (x, iterations)."""
        x = x0.copy()
        tol = rtol * np.linalg.norm(Binv(b))
        r = b - A @ x
        z = Binv(r)
        if np.linalg.norm(z) <= tol:
            return x, 0
        p, rz = z.copy(), r @ z
        for it in range(1, maxit + 1):
            q = A @ p
            alpha = rz / (p @ q)
            x += alpha * p
            r -= alpha * q
            z = Binv(r)
            rz_new = r @ z
            if np.linalg.norm(z) <= tol:
                return x, it
            p = z + (rz_new / rz) * p
            rz = rz_new
        return x, maxit

    # ----------------------------------------------------------------- time loop
    def step(self, solver="direct", Pinv=None, rtol=1e-9, x_prev=None, first=False):
        """One pass of the SolverKNPEMI.solve loop body (KNPEMIx_solver.py:365-468)."""
        p = self.p
        self.t += p.dt
        if any(nm == "HH" for nm, _ in self.models):
            self.gate_update()
        A, b = self.assemble(self.t)
        ns = self.nullspace()
        if first:
            b = b - ns * (ns @ b)                                       # nullspace.remove(b), step 1 only
        if solver == "direct":
            x, its = self.solve_direct(A, b, ns), 0
        else:
            idx, g = self.bc_dofs()
            x0 = x_prev.copy()
            x0[idx] = g                                                 # the initial guess carries the boundary values
            x, its = self.solve_gmres(A, b, x0, ns, Pinv, rtol)
            x[idx] = g                                                  # ... and so does the solution, exactly
        self.unpack(x)
        return A, b, x, its

    def run(self, steps, solver="direct", rtol=1e-9, Pinv_factory=None):
        its_all = []
        x = self.pack()                                                  # ICs = initial guess (solver.py:179-209)
        Pinv = None
        if solver != "direct":
            P = self.assemble_P()
            Pinv = Pinv_factory(P) if Pinv_factory else (lambda v, lu=spla.splu(P.tocsc()): lu.solve(v))
        for i in range(1, steps + 1):
            A, b, x, its = self.step(solver, Pinv, rtol, x, first=(i == 1))
            its_all.append(its)
        return its_all

    # -------------------------------------------------------------- functionals
    def integral(self, u, tags, power=1):
        """int u^power dx(tags) for a P1 field on all vertices: power 1 = the ion amounts of
        ProblemKNPEMI.print_conservation (KNPEMIx_problem.py:807-843), power 0 = the measure of the tagged cells."""
        m = self.mesh
        d = m.gdim
        sel = np.isin(m.cell_tags, np.atleast_1d(tags))
        cells = m.cells[sel]
        x = m.x[cells]
        vol = np.abs(np.linalg.det(x[:, 1:] - x[:, :1])) / (2.0 if d == 2 else 6.0)
        if power == 0:
            return float(vol.sum())
        if power == 1:
            return float((vol * u[cells].mean(axis=1)).sum())
        return self.l2_norm(u, tags) ** 2

    def membrane_area(self, tag):
        return float(self.farea[self.mesh.mf_tags == tag].sum())

    def stimulus_current(self, t):
        """int stim_expr dS(stimulus_tags) (KNPEMIx_solver.py:578-610 with the expression of
        HodgkinHuxley._add_stimulus, KNPEMIx_ionic_model.py:517-603), from the current fields."""
        p, m = self.p, self.mesh
        self._stim_area = self.stimulus_area() if p.scale_stimulus else 1.0
        t_mod = np.mod(t + 1e-12, p.T_stim)
        ci, ce, phim, gq, xq = self._facet_quadrature_fields()
        sel = np.isin(m.mf_tags, np.asarray(p.stimulus_tags))
        E_Na = (p.psi / p.z[0]) * np.log(ce[0] / ci[0])
        stim = self.stimulus_mask(xq) * p.g_syn_bar * np.exp(-t_mod / p.a_syn) * (phim - E_Na)
        if p.scale_stimulus:
            stim = stim / self._stim_area
        return float(np.sum((self.farea[:, None] * self.qw[None, :] * stim)[sel]))

    def l2_norm(self, u, tags):
        """sqrt(int u^2 dx(tags)) for a P1 field given on all vertices."""
        m = self.mesh
        d = m.gdim
        sel = np.isin(m.cell_tags, np.atleast_1d(tags))
        cells = m.cells[sel]
        x = m.x[cells]
        J = x[:, 1:] - x[:, :1]
        vol = np.abs(np.linalg.det(J)) / (2.0 if d == 2 else 6.0)
        uc = u[cells]
        ssum = uc.sum(1)
        integ = vol / ((d + 1) * (d + 2)) * ((uc ** 2).sum(1) + ssum ** 2)
        return float(np.sqrt(integ.sum()))
