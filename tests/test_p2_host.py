"""P2 element path (fem_order = 2) without a GPU: the product's node mesh, dof maps and CSR pattern against the oracle's
(bit-exact), and one assembly by knp_p2_emulate_host -- the functions the P2 kernels run per thread, called in a loop on the
CPU -- against the oracle's matrices and vectors (1e-12 relative to the row's largest entry)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.fixtures import unit_square, unit_cube, from_arrays
from oracle.knpemi import OracleParams
from oracle.p2 import KNPEMIOracleP2, p2_mesh
from conftest import MODELS_TEST, params_struct, perturb


def _mesh(kb, name):
    if name == "square8":
        return unit_square(8), OracleParams(stimulus_region=(0, 0.2e-6, 0.6e-6))
    if name == "cube4":
        return unit_cube(4), OracleParams(stimulus_region=((0, 0.2e-6, 0.8e-6), (2, 0.0, 0.6e-6)))
    if name == "plates3d":      # BASELINE C5 in miniature: plate-stack cells, every intracellular vertex on the membrane
        mm = kb.mesh.cell_array_mesh(3, 8, 2, fill=0.75, shape={"plates": True, "thickness": 1, "pitch": 2, "spine": 1})
    else:
        d, n, m = (2, 12, 3) if name == "cells2d" else (3, 6, 2)
        mm = kb.mesh.cell_array_mesh(d, n, m)
    d = mm.gdim
    it = tuple(mm.intra_tags)
    return (from_arrays(d, mm.x, mm.cells, mm.cell_tags, mm.intra_tags),
            OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,)))


def _product_node_mesh(kb, om, p):
    m = kb.mesh.Mesh(om.gdim, om.x, om.cells.astype(np.int32), om.cell_tags.astype(np.int32), tuple(p.intra_tags), p.extra_tag,
                     om.mf_verts.astype(np.int32), om.mf_tags.astype(np.int32))
    return kb.mesh.p2_node_mesh(m)


def rel_rows(A_ref, vals):
    scale = np.maximum.reduceat(np.abs(A_ref.data), A_ref.indptr[:-1])
    return (np.abs(vals - A_ref.data) / np.repeat(scale, np.diff(A_ref.indptr))).max()


GLIA = [("KirNa", None), ("GlialCT", None), ("Passive", None)]


@pytest.mark.parametrize("name,models", [("square8", MODELS_TEST), ("cube4", MODELS_TEST), ("cells2d", MODELS_TEST),
                                         ("cells3d", MODELS_TEST), ("plates3d", MODELS_TEST), ("square8", GLIA), ("cube4", GLIA)])
def test_p2_node_mesh_pattern_and_emulated_assembly(kb, name, models):
    om, p = _mesh(kb, name)
    m2 = _product_node_mesh(kb, om, p)
    o2m = p2_mesh(om)
    assert np.array_equal(m2.x, o2m.x) and np.array_equal(m2.cells, o2m.cells) and np.array_equal(m2.mf_verts, o2m.mf_verts)
    o = perturb(KNPEMIOracleP2(o2m, p, models), seed=3)
    t = 3 * p.dt
    A, b = o.assemble(t)
    P = o.assemble_P()
    qb, qw = kb.mesh.facet_quadrature(om.gdim)
    args = (om.gdim, m2.x, m2.cells, m2.cell_tags, p.intra_tags, p.extra_tag, m2.mf_verts, m2.mf_tags, qb, qw)
    ip, ix, vi, ve = kb.lib.pattern_host(*args, degree=2)
    assert np.array_equal(ip, A.indptr) and np.array_equal(ix, A.indices)
    assert np.array_equal(vi, o.S[0]) and np.array_equal(ve, o.S[1])
    Pm, table = params_struct(kb, p, models, stim_area=o.stimulus_area())
    tm = [kb.lib.TagModels(tg, fl, int(st)) for tg, fl, st in table]
    ip2, ix2, vals, bv = kb.lib.p2_emulate_host(Pm, tm, t, 0, o.pack(), o.gates[:, o.mverts], *args)
    assert np.array_equal(ip2, A.indptr) and np.array_equal(ix2, A.indices)
    assert rel_rows(A, vals) < 1e-12
    for s in range(2):
        for f in range(4):
            sl = slice(o.base[s] + f * o.ns[s], o.base[s] + (f + 1) * o.ns[s])
            assert np.abs(bv[sl] - b[sl]).max() <= 1e-12 * np.abs(b[sl]).max()
    pv = kb.lib.p2_emulate_host(Pm, tm, t, 1, o.pack(), None, *args)[:P.nnz]
    assert rel_rows(P, pv) < 1e-12
    # the two auxiliary matrices of the Schur preconditioner come from the same rows with modified constants
    Pm2, _ = params_struct(kb, p, models, stim_area=1.0)
    Pm2.C_M = -p.C_M
    pv2 = kb.lib.p2_emulate_host(Pm2, tm, t, 1, o.pack(), None, *args)[:P.nnz]
    assert rel_rows(o.assemble_P(membrane_sign=+1.0), pv2) < 1e-12


def test_p2_boundary_nodes_shape_functions_and_reference_mass(kb):
    """Boundary nodes of the P2 space (vertices and edge nodes of the exterior facets) and P2 point evaluation."""
    om = unit_square(8)
    m = kb.mesh.Mesh(2, om.x, om.cells.astype(np.int32), om.cell_tags.astype(np.int32), (1,), 2, om.mf_verts.astype(np.int32),
                     om.mf_tags.astype(np.int32))
    m2 = kb.mesh.p2_node_mesh(m)
    bn = kb.mesh.boundary_vertices(m2)
    x = m2.x / 1e-6
    on = (np.abs(x[:, 0]) < 1e-12) | (np.abs(x[:, 0] - 1) < 1e-12) | (np.abs(x[:, 1]) < 1e-12) | (np.abs(x[:, 1] - 1) < 1e-12)
    assert np.array_equal(bn, np.flatnonzero(on))
    # a quadratic field is reproduced exactly by the P2 shape functions at an arbitrary point of a cell
    c = m2.cells[37]
    bary = np.array([0.2, 0.5, 0.3])
    pt = bary @ m2.x[c[:3]]
    f = lambda y: 1.0 + y[..., 0] * 3e6 + (y[..., 0] * 1e6) ** 2 - 2.0 * (y[..., 0] * 1e6) * (y[..., 1] * 1e6)
    assert abs(kb.mesh.p2_shape(bary) @ f(m2.x[c]) - f(pt)) < 1e-12
    # reference mass matrices (exact monomial integrals) against the oracle's quadrature
    for d in (2, 3):
        o = KNPEMIOracleP2(unit_square(2) if d == 2 else unit_cube(2), OracleParams(), [("Passive", None)])
        assert np.abs(kb.mesh.reference_mass(d, 2) - o.Mref).max() < 1e-15



@pytest.mark.parametrize("case", ["ion_2d", "potential_2d", "ion_3d", "potential_3d"])
def test_host_amg_setup_on_p2_blocks_matches_oracle(kb, case):
    """The hierarchy setup the P2 contexts use (amg_setup.cpp on the host) on the ion / potential blocks of the P2 Schur
    preconditioner against oracle/amg.py level by level (P2 stiffness matrices have positive off-diagonal entries)."""
    from oracle.amg import SAAMG, SchurPC
    om, p = (unit_square(16), OracleParams()) if "2d" in case else (unit_cube(6), OracleParams())
    o = perturb(KNPEMIOracleP2(om, p, MODELS_TEST), seed=4)
    pc = SchurPC(o, exact=True)
    Pt = o.assemble_P(membrane_sign=+1.0).tocsr()
    idx = pc.ic if "ion" in case else pc.ip
    A = Pt[idx][:, idx].tocsr()
    ref = SAAMG(A, coarse_size=100)
    levels = kb.lib.amg_setup_host(A, theta=0.08, coarse_size=100)
    ref_ops = [lv["A"] for lv in ref.levels] + [ref.Ac]
    assert [a.shape[0] for a in levels] == [a.shape[0] for a in ref_ops]
    assert len(levels) >= 2
    for a, r in zip(levels, ref_ops):
        dd = (a - r).tocoo()
        assert dd.nnz == 0 or np.abs(dd.data).max() <= 1e-10 * np.abs(r.data).max()
    # HRZ-lumped M_sigma is positive on every dof (row sums are not)
    for s in range(2):
        assert (pc.msig[s] > 0).all()


P2_YAML = """
problem_type: "KNP-EMI"
dt: 2.5e-5
time_steps: 2
fem_order: 2
physical_constants: {T: 300, F: 96485, R: 8.314}
C_M: 0.02
mesh_file: "./input/geometries/%s.xdmf"
cell_tag_file: "./input/geometries/%s.xdmf"
facet_tag_file: "./input/geometries/%s_facets.xdmf"
ics_tags: [1]
ecs_tags: [2]
boundary_tags: [3]
membrane_tags: [4]
mesh_conversion_factor: 1e-6
source_terms: "ion_injection"
initial_conditions:
  {phi_m: -0.070, Na_i: 12, Na_e: 140, K_i: 130, K_e: 4, Cl_i: 5, Cl_e: 125, n: 0.276, m: 0.0379, h: 0.688}
solver:
  direct: False
  ksp_settings: {ksp_rtol: 1.0e-9, ksp_type: gmres, pc_type: hypre, norm_type: preconditioned, non_zero_init_guess: True}
  output: {save_xdmf: False, save_cpoints: False, save_pngs: False, save_dat: False}
"""


TISSUE_YAML = """
problem_type: "KNP-EMI"
dt: 2.5e-5
time_steps: 2
fem_order: %d
physical_constants: {T: 300, F: 96485, R: 8.314}
C_M: 0.02
synthetic_mesh: {kind: cell_array, dim: %d, N: %d, cells_per_dim: 2, fill: 0.5, first_tag: 2, extra_tag: 1}
ics_tags: !range [2, %d]
ecs_tags: [1]
membrane_tags: !range [2, %d]
mesh_conversion_factor: 1e-6
source_terms: "ion_injection"
initial_conditions:
  {phi_m: -0.070, Na_i: 12, Na_e: 140, K_i: 130, K_e: 4, Cl_i: 5, Cl_e: 125, n: 0.276, m: 0.0379, h: 0.688}
solver:
  direct: False
  ksp_settings: {ksp_rtol: 1.0e-9, ksp_type: gmres, pc_type: hypre, norm_type: preconditioned, non_zero_init_guess: True}
  output: {save_xdmf: False, save_cpoints: False, save_pngs: False, save_dat: False}
"""


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("dim,n", [(2, 20), (3, 10)])
def test_host_path_and_ion_injection_source(kb, tmp_path, order, dim, n):
    """fem_order 1 / 2 through ProblemKNPEMI on the CPU for a tissue block whose centre (the injection site) lies in the
    extracellular space: node mesh, restrictions and the ion-injection entries of the right-hand side against the oracle;
    creating the device context without a GPU fails loudly (no CPU fallback)."""
    ncell = 2 ** dim
    cfg = tmp_path / "tissue.yaml"
    cfg.write_text(TISSUE_YAML % (order, dim, n, 2 + ncell, 2 + ncell))
    p = kb.ProblemKNPEMI(str(cfg), verbose=False)
    mm = kb.mesh.cell_array_mesh(dim, n, 2)
    it = tuple(mm.intra_tags)
    om = from_arrays(dim, mm.x, mm.cells, mm.cell_tags, mm.intra_tags)
    from oracle.knpemi import KNPEMIOracle
    cls = KNPEMIOracleP2 if order == 2 else KNPEMIOracle
    kw = dict(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,), c_i_init=(12.0, 130.0, 5.0), c_e_init=(140.0, 4.0, 125.0))
    o = cls(om, OracleParams(source_terms="ion_injection", **kw), MODELS_TEST)
    assert p.mesh.degree == order and np.array_equal(p.mesh.cells, o.mesh.cells) and np.array_equal(p.mesh.mf_verts, o.mesh.mf_verts)
    assert np.array_equal(p.dofs_intra, o.S[0]) and np.array_equal(p.dofs_extra, o.S[1])
    assert abs(p.injection_volume - o.injection_volume) <= 1e-14 * o.injection_volume
    # source entries: b with the source minus b without it
    rows, vals = p._source_entries([o.S[0].astype(np.int32), o.S[1].astype(np.int32)])
    _, b1 = o.assemble(2.5e-5)
    _, b0 = cls(om, OracleParams(**kw), MODELS_TEST).assemble(2.5e-5)
    ref = b1 - b0
    assert np.abs(ref).max() > 0.0 and rows.size > 0
    got = np.zeros(o.n)
    got[rows] = vals
    # ref is a difference of right-hand sides that are 1e6 times larger than the source: cancellation limits it to ~1e-9
    assert np.abs(got - ref).max() <= min(1e-6 * np.abs(ref).max(), 1e-14 * np.abs(b1).max())
    try:
        import torch
        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    if not gpu:
        p.set_initial_conditions()
        p.init_ionic_models([kb.NeuronalCotransporters(p), kb.HodgkinHuxley(p), kb.ATPPump(p)])
        with pytest.raises(kb.lib.KnpError):
            p.setup_variational_form()


@pytest.mark.parametrize("order", [1, 2])
def test_dirichlet_entries_match_the_oracle(kb, tmp_path, order):
    """dirichlet_bcs through ProblemKNPEMI on the CPU: the constrained columns and values handed to knp_set_dirichlet against
    the oracle's bc_dofs (P1: boundary vertices; P2: vertices and edge nodes of the exterior facets)."""
    cfg = tmp_path / "bc.yaml"
    txt = (P2_YAML % ("square16", "square16", "square16")).replace('source_terms: "ion_injection"', "dirichlet_bcs: True")
    cfg.write_text(txt.replace("fem_order: 2", f"fem_order: {order}"))
    p = kb.ProblemKNPEMI(str(cfg), verbose=False)
    p.set_initial_conditions()            # the boundary values are the initial values of the config
    om = unit_square(16)
    mesh = p2_mesh(om) if order == 2 else om
    bn = kb.mesh.boundary_vertices(p.mesh)
    from oracle.knpemi import KNPEMIOracle
    cls = KNPEMIOracleP2 if order == 2 else KNPEMIOracle
    prm = OracleParams(dirichlet_bcs=True, boundary_verts=tuple(int(v) for v in bn), c_i_init=(12.0, 130.0, 5.0),
                       c_e_init=(140.0, 4.0, 125.0))
    o = cls(mesh, prm, MODELS_TEST)
    idx, g = o.bc_dofs()
    cols, vals = p._bc_entries([o.S[0].astype(np.int32), o.S[1].astype(np.int32)], o.mverts)
    order_ = np.argsort(cols)
    assert np.array_equal(cols[order_], idx) and np.array_equal(vals[order_], g)
    assert idx.size == 4 * bn.size          # the exterior boundary belongs to the extracellular space only


@pytest.mark.parametrize("order", [1, 2])
def test_probe_tables_reproduce_polynomial_fields(kb, tmp_path, order):
    """point_evaluation through ProblemKNPEMI on the CPU: the sparse functionals handed to knp_probe_setup evaluate a field of
    the space exactly (linear for P1, quadratic for P2) at points inside cells and on the membrane."""
    cfg = tmp_path / "probe.yaml"
    txt = (P2_YAML % ("square16", "square16", "square16")).replace('source_terms: "ion_injection"', """point_evaluation:
  ics_points: [[0.43, 0.52], [0.3, 0.7]]
  ecs_points: [[0.11, 0.93], [0.8, 0.13]]
  gamma_points: [[0.25, 0.4], [0.6, 0.75]]""")
    cfg.write_text(txt.replace("fem_order: 2", f"fem_order: {order}"))
    p = kb.ProblemKNPEMI(str(cfg), verbose=False)
    om = unit_square(16)
    o = (KNPEMIOracleP2 if order == 2 else __import__("oracle.knpemi", fromlist=["KNPEMIOracle"]).KNPEMIOracle)(
        om, OracleParams(), MODELS_TEST)
    nv = [o.S[0].astype(np.int32), o.S[1].astype(np.int32)]
    ptr, cols, wts = p._probe_tables(nv)
    x = p.mesh.x / 1e-6
    f = (lambda y: 2.0 - y[:, 0] + 3.0 * y[:, 1] + (y[:, 0] ** 2 - 0.5 * y[:, 0] * y[:, 1] if order == 2 else 0.0))
    # state vector in the column layout: every field carries f (phi_e carries 2 f so that phi_m = phi_i - phi_e = -f)
    u = np.concatenate([f(x[nv[0]])] * 4 + [f(x[nv[1]])] * 3 + [2.0 * f(x[nv[1]])])
    vals = np.array([np.dot(wts[ptr[i]:ptr[i + 1]], u[cols[ptr[i]:ptr[i + 1]]]) for i in range(len(ptr) - 1)])
    pts = np.array([[0.43, 0.52], [0.3, 0.7]]), np.array([[0.11, 0.93], [0.8, 0.13]]), np.array([[0.25, 0.4], [0.6, 0.75]])
    ref = np.concatenate([np.repeat(f(pts[0]), 4), np.repeat(f(pts[1]), 4) * np.tile([1, 1, 1, 2.0], 2), -f(pts[2])])
    assert np.abs(vals - ref).max() < 1e-12


def test_upload_wrappers_hand_the_tables_to_the_context(kb, tmp_path):
    """_upload_source / _upload_bcs / _setup_probes with a recording stand-in for the device context (the pure table builders
    are checked above; this guards the glue that the GPU tests exercise)."""
    class Recorder:
        def __init__(self):
            self.calls = {}

        def __getattr__(self, name):
            return lambda *a: self.calls.__setitem__(name, a)

    txt = (TISSUE_YAML % (1, 2, 20, 6, 6)) + """
dirichlet_bcs: True
boundary_tags: [1]
point_evaluation:
  ics_points: [[0.25, 0.26]]
  ecs_points: [[0.51, 0.52]]
"""
    cfg = tmp_path / "glue.yaml"
    cfg.write_text(txt)
    p = kb.ProblemKNPEMI(str(cfg), verbose=False)
    p.set_initial_conditions()
    from oracle.knpemi import KNPEMIOracle
    mm = kb.mesh.cell_array_mesh(2, 20, 2)
    it = tuple(mm.intra_tags)
    o = KNPEMIOracle(from_arrays(2, mm.x, mm.cells, mm.cell_tags, mm.intra_tags),
                     OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,)), MODELS_TEST)
    p._ctx, p._node_vert, p._mverts = Recorder(), [o.S[0].astype(np.int32), o.S[1].astype(np.int32)], o.mverts
    p._upload_source()
    p._upload_bcs()
    p._setup_probes()
    rows, vals = p._ctx.calls["set_source"]
    assert rows.size == vals.size > 0
    cols, g = p._ctx.calls["set_dirichlet"]
    assert cols.size == g.size == 4 * kb.mesh.boundary_vertices(p.mesh).size
    ptr, pc, pw = p._ctx.calls["probe_setup"]
    assert len(ptr) == 1 + 8 and len(pc) == len(pw) == 8 * 3
