"""CPU-only checks of the host side: YAML schema / error behaviour of the reference, mesh fixtures against the
oracle's independent generator, quadrature tables, the membrane-model table, and the C-ABI library itself
(it must load and export every symbol declared in include/knpemi_b200.h; it must refuse to run without a GPU)."""
import os
import re
import numpy as np
import pytest
from oracle.fixtures import unit_square, unit_cube, from_arrays
from oracle.quadrature import facet_rule
from conftest import has_gpu

BASE = """
problem_type: "KNP-EMI"
dt: 2.5e-5
time_steps: 2
physical_constants: {T: 300, F: 96485, R: 8.314}
C_M: 0.02
cell_tag_file: "./input/geometries/square8.xdmf"
facet_tag_file: "./input/geometries/square8_facets.xdmf"
ics_tags: [1]
ecs_tags: [2]
boundary_tags: [3]
membrane_tags: [4]
mesh_conversion_factor: 1e-6
initial_conditions: {phi_m: -0.070, Na_i: 12, Na_e: 140, K_i: 130, K_e: 4, Cl_i: 5, Cl_e: 125, n: 0.276, m: 0.0379, h: 0.688}
stimulus: {conductance: {g_syn_bar: 1.0e-9}, a_syn: 5.0e-4, T_stim: 1.0, scale: True}
solver: {direct: True, output: {save_xdmf: False}}
"""


def write(tmp_path, text, name="cfg.yaml"):
    f = tmp_path / name
    f.write_text(text)
    return str(f)


def drop(text, key):
    return "\n".join(l for l in text.splitlines() if not l.startswith(key))


def test_config_schema_and_errors(kb, tmp_path):
    p = kb.ProblemKNPEMI(write(tmp_path, BASE), verbose=False)
    assert p.time_steps == 2 and p.dt.value == 2.5e-5 and p.N_ions == 3
    assert p.intra_tags == (1,) and p.extra_tag == (2,) and p.gamma_tags == (4,) and p.stimulus_tags == (4,)
    assert abs(p.psi.value - 8.314 * 300 / 96485) < 1e-18
    # conductance defaults when a stimulus dict is present (mixed_dim_problem.py:311-318)
    assert (p.g_Na_leak.value, p.g_K_leak.value, p.g_Cl_leak.value) == (0.3, 0.1, 0.25)
    assert p.solver_config["direct"] is True
    # required keys raise RuntimeError like the reference (mixed_dim_problem.py:98-166)
    for key, msg in [("solver", "solver configuration"), ("dt", "dt"), ("time_steps", "time_steps"), ("ics_tags", "ics_tags")]:
        with pytest.raises(RuntimeError, match=msg):
            kb.ProblemKNPEMI(write(tmp_path, drop(BASE, key)), verbose=False)
    with pytest.raises(RuntimeError, match="cell_tag_file"):
        kb.ProblemKNPEMI(write(tmp_path, drop(drop(BASE, "cell_tag_file"), "facet_tag_file")), verbose=False)
    with pytest.raises(RuntimeError, match="scale"):
        kb.ProblemKNPEMI(write(tmp_path, BASE.replace(", scale: True", "")), verbose=False)
    # T instead of time_steps: int(T/dt) (mixed_dim_problem.py:157)
    p2 = kb.ProblemKNPEMI(write(tmp_path, BASE.replace("time_steps: 2", "T: 1.0e-4")), verbose=False)
    assert p2.time_steps == int(1.0e-4 / 2.5e-5)
    # no stimulus block -> other defaults (mixed_dim_problem.py:320-332)
    p3 = kb.ProblemKNPEMI(write(tmp_path, drop(BASE, "stimulus")), verbose=False)
    assert (p3.g_Na_leak.value, p3.g_K_leak.value, p3.g_syn_bar.value, p3.scale_stimulus) == (1.0, 4.0, 40.0, False)


def test_range_tag_and_tag_order_check(kb, tmp_path):
    txt = BASE.replace("ics_tags: [1]", "ics_tags: !range [2, 6]").replace("ecs_tags: [2]", "ecs_tags: [1]") \
              .replace("membrane_tags: [4]", "membrane_tags: !range [2, 6]")
    txt = drop(drop(txt, "cell_tag_file"), "facet_tag_file") + \
        "\nsynthetic_mesh: {kind: cell_array, dim: 2, N: 16, cells_per_dim: 2, first_tag: 2, extra_tag: 1}\n"
    p = kb.ProblemKNPEMI(write(tmp_path, txt), verbose=False)
    assert p.intra_tags == (2, 3, 4, 5) and p.gamma_tags == (2, 3, 4, 5)
    assert sorted(np.unique(p.mesh.mf_tags)) == [2, 3, 4, 5]
    bad = txt.replace("ecs_tags: [1]", "ecs_tags: [3]")
    with pytest.raises(RuntimeError, match="all smaller or all larger"):
        kb.ProblemKNPEMI(write(tmp_path, bad), verbose=False)


def test_model_table_and_mismatch(kb, tmp_path):
    p = kb.ProblemKNPEMI(write(tmp_path, BASE), verbose=False)
    HH, ATP, NCT = kb.HodgkinHuxley(p), kb.ATPPump(p), kb.NeuronalCotransporters(p)
    assert HH.tags == (4,) and HH.time_steps_ODE == 25 and abs(HH.dt_ode - 1e-6) < 1e-20
    p.set_initial_conditions()
    p.init_ionic_models([NCT, HH, ATP])
    assert p.gating_variables
    assert p._tag_table() == [(4, kb.lib.MODEL_HH | kb.lib.MODEL_ATP | kb.lib.MODEL_NEURONAL_CT, True)]
    assert np.all(p.n.x.array == 0.276) and np.all(p.wh[0][3].x.array == -0.070) and np.all(p.wh[1][1].x.array == 4)
    with pytest.raises(RuntimeError, match="Mismatch between membrane tags"):
        p.init_ionic_models([kb.PassiveModel(p, tags=(7,))])


def test_mesh_fixtures_match_oracle_generator(kb):
    for n in (4, 32):
        m, o = kb.mesh.unit_square_fixture(n), unit_square(n)
        assert np.array_equal(m.cells, o.cells) and np.array_equal(m.x, o.x) and np.array_equal(m.cell_tags, o.cell_tags)
        assert np.array_equal(m.mf_verts, o.mf_verts) and np.all(m.mf_tags == 4)
    m, o = kb.mesh.unit_cube_fixture(6), unit_cube(6)
    assert np.array_equal(m.cells, o.cells) and np.array_equal(m.mf_verts, o.mf_verts)
    # C1 sizes quoted by the survey
    m = kb.mesh.unit_square_fixture(32)
    assert m.x.shape[0] == 1089 and m.cells.shape[0] == 2048 and (m.cell_tags == 1).sum() == 512 and m.mf_verts.shape[0] == 64
    # tissue block: facets agree with the oracle's generic facet matcher
    m = kb.mesh.cell_array_mesh(2, 32, 4)
    o = from_arrays(2, m.x, m.cells, m.cell_tags, m.intra_tags)
    assert np.array_equal(m.mf_verts, o.mf_verts) and np.array_equal(m.mf_tags, o.mf_tags)
    m3 = kb.mesh.cell_array_mesh(3, 8, 2)
    o3 = from_arrays(3, m3.x, m3.cells, m3.cell_tags, m3.intra_tags)
    assert np.array_equal(m3.mf_verts, o3.mf_verts) and np.array_equal(m3.mf_tags, o3.mf_tags)
    assert m3.mf_verts.shape[0] == 8 * 6 * 2 * 2 * 2      # 8 cubes x 6 faces x (2x2 squares) x 2 triangles


def test_quadrature_tables(kb):
    for d in (2, 3):
        b, w = kb.mesh.facet_quadrature(d)
        ob, ow = facet_rule(d)
        np.testing.assert_allclose(b, ob, atol=1e-15)
        np.testing.assert_allclose(w, ow, atol=1e-15)
        assert abs(w.sum() - 1) < 1e-14 and np.allclose(b.sum(1), 1)


def test_library_exports_every_declared_symbol(kb):
    lib = kb.lib.load()
    header = open(os.path.join(os.path.dirname(os.path.dirname(kb.__file__)), "include", "knpemi_b200.h")).read()
    declared = set(re.findall(r"\b(knp_[A-Za-z0-9_]+)\s*\(", header))
    assert declared == set(kb.lib.SYMBOLS), declared ^ set(kb.lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.knp_version() >= 100


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(kb, tmp_path):
    p = kb.ProblemKNPEMI(write(tmp_path, BASE), verbose=False)
    p.set_initial_conditions()
    p.init_ionic_models([kb.PassiveModel(p)])
    with pytest.raises(kb.lib.KnpError, match="no CPU fallback"):
        p.setup_variational_form()


def sp_diag(A):
    import scipy.sparse as sp
    return sp.diags(A.diagonal())


@pytest.mark.parametrize("case", ["schur_ion_2d", "schur_phi_2d", "schur_ion_3d", "jacobi_P_2d", "schur_ion_2d_dirichlet",
                                  "schur_phi_2d_dirichlet", "schur_phi_3d_dirichlet"])
def test_native_amg_setup_matches_oracle_level_by_level(kb, case):
    """amg_setup.cpp (host code of libknpemi_b200.so, no GPU needed) against oracle/amg.py: same MIS(2) aggregates,
    filtered prolongator smoothing, adaptive strength threshold and Galerkin products -> the level operators agree.
    *_dirichlet: blocks with essential boundary rows (identity rows); those dofs leave the coarse space on both sides, so
    the first coarse level is as small as that of the interior problem."""
    from oracle.amg import SAAMG, SchurPC
    from oracle.fixtures import from_arrays
    from oracle.knpemi import KNPEMIOracle, OracleParams
    from conftest import MODELS_TEST
    d = 3 if "3d" in case else 2
    mm = kb.mesh.cell_array_mesh(d, 24 if d == 2 else 12, 3 if d == 2 else 2)
    it = tuple(mm.intra_tags)
    bc = dict(dirichlet_bcs=True, boundary_verts=tuple(kb.mesh.boundary_vertices(mm))) if "dirichlet" in case else {}
    o = KNPEMIOracle(from_arrays(d, mm.x, mm.cells, mm.cell_tags, mm.intra_tags),
                     OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,), **bc), MODELS_TEST)
    rng = np.random.default_rng(4)
    for s in range(2):
        o.c[s] *= 1 + 0.05 * rng.random(o.c[s].shape)
    if case == "jacobi_P_2d":
        A = o.assemble_P().tocsr()
    else:
        pc = SchurPC(o, exact=True)
        Pt = o.assemble_P(membrane_sign=+1.0).tocsr()
        idx = pc.ic if "ion" in case else pc.ip
        A = Pt[idx][:, idx].tocsr()
    ref = SAAMG(A, coarse_size=100)
    levels = kb.lib.amg_setup_host(A, theta=0.08, coarse_size=100)
    ref_ops = [lv["A"] for lv in ref.levels] + [ref.Ac]
    assert [a.shape[0] for a in levels] == [a.shape[0] for a in ref_ops]
    assert len(levels) >= 2
    if bc:
        nb = int((np.diff(A.indptr) > 0).sum() - (abs(A - sp_diag(A)).sum(axis=1) > 0).sum())      # identity rows
        assert nb > 0 and levels[1].shape[0] < (A.shape[0] - nb) / 2
    for a, r in zip(levels, ref_ops):
        dd = (a - r).tocoo()
        assert dd.nnz == 0 or np.abs(dd.data).max() <= 1e-10 * np.abs(r.data).max()


@pytest.mark.parametrize("name", ["square32", "square7", "cube6", "cells2d", "cells3d"])
def test_native_csr_pattern_and_dofmaps_bit_exact_on_host(kb, name):
    """The structure builder of libknpemi_b200.so (topology.cpp; host code, no GPU) against the oracle: restricted dof
    maps and the CSR pattern of A bit for bit (north star: 'bit-exact CSR structure and DOF maps')."""
    from oracle.fixtures import from_arrays, unit_cube, unit_square
    from oracle.knpemi import KNPEMIOracle, OracleParams
    from conftest import MODELS_TEST
    if name.startswith("cells"):
        d = 2 if name == "cells2d" else 3
        mm = kb.mesh.cell_array_mesh(d, 24 if d == 2 else 8, 3 if d == 2 else 2)
        om = from_arrays(d, mm.x, mm.cells, mm.cell_tags, mm.intra_tags)
        it = tuple(mm.intra_tags)
        p = OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,))
    else:
        om = {"square32": lambda: unit_square(32), "square7": lambda: unit_square(7), "cube6": lambda: unit_cube(6)}[name]()
        p = OracleParams()
    o = KNPEMIOracle(om, p, MODELS_TEST)
    A, _ = o.assemble(p.dt)
    qb, qw = kb.mesh.facet_quadrature(om.gdim)
    ip, ix, vi, ve = kb.lib.pattern_host(om.gdim, om.x, om.cells, om.cell_tags, p.intra_tags, p.extra_tag, om.mf_verts,
                                         om.mf_tags, qb, qw)
    assert np.array_equal(ip, A.indptr) and np.array_equal(ix, A.indices)
    assert np.array_equal(vi, o.S[0]) and np.array_equal(ve, o.S[1])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the C++/OpenMP CPU port of the oracle, all host cores, unscaled) runs without a GPU and
    prints ONE JSON line with the keys the driver reads; here on a reduced mesh (--size 64) to keep the CPU tier short."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                        "--size", "64"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is False and d["unit"] == "ms"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["extrapolated"] is False
    assert d["cpu_baseline"]["value"] == d["value"] and "N=64" in d["config"]["workload"]
    assert "workload" in d["config"] and d["value"] > 0


def test_native_spmv_rowblocks_cover_every_row_once(kb):
    """Row blocks of the TMA-staged SpMV (linalg.cu::build_rowblocks): every row in exactly one block, 16-byte aligned
    TMA sources (first row and staged start multiples of 4), stage capacity and row limit respected."""
    rng = np.random.default_rng(3)
    for lens in (rng.integers(7, 29, 5000), np.full(3000, 7), rng.integers(1, 120, 2000), np.array([3]), np.zeros(0, int)):
        indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        blk = kb.lib.rowblocks_host(indptr)
        n = indptr.size - 1
        if n == 0:
            assert blk is None
            continue
        assert blk is not None
        r0, nr, a4, cnt = blk.T
        assert r0[0] == 0 and np.array_equal(r0[1:], (r0 + nr)[:-1]) and r0[-1] + nr[-1] == n      # contiguous cover
        assert (r0 % 4 == 0).all() and (a4 % 4 == 0).all()
        assert np.array_equal(a4, indptr[r0] & ~3) and np.array_equal(cnt, indptr[r0 + nr] - a4)
        assert (nr <= 256).all() and (cnt <= 1280).all() and (nr > 0).all()
    # a group of four rows that cannot fit one stage -> no blocks (the CSR-vector kernel takes over)
    assert kb.lib.rowblocks_host(np.array([0, 2000, 4000], np.int32)) is None


def test_solver_option_mapping(kb, tmp_path):
    """SolverKNPEMI._opts: how the reference's solver settings (KNPEMIx_solver.py:25-51,164-291) map onto the C-ABI
    knp_solve_opts -- `pc_type: hypre` -> the Schur preconditioner (pc 3) unless amg_form says block_jacobi, gamg -> the
    SA-AMG cycle on the reference's P (pc 2); unsupported choices raise instead of silently doing something else."""
    it = BASE.replace("solver: {direct: True, output: {save_xdmf: False}}",
                      "solver: {direct: False, ksp_settings: {ksp_rtol: 1.0e-9, ksp_type: gmres, pc_type: hypre, "
                      "norm_type: preconditioned, non_zero_init_guess: True}, output: {save_xdmf: False}}")
    p = kb.ProblemKNPEMI(write(tmp_path, it), verbose=False)
    p.solver_config["view_ksp"] = False
    s = kb.SolverKNPEMI(p, solver_config=p.solver_config)
    o = s._opts()
    assert (o.pc, o.rtol, o.restart, o.max_it, o.project_nullspace, o.zero_mean_solution, o.refine) == (3, 1e-9, 30, 5000, 1, 0, 0)
    s.amg_form = "block_jacobi"
    assert s._opts().pc == 2
    s.amg_form = "schur"
    for pc_type, pc in (("gamg", 2), ("schur", 3), ("jacobi", 1), ("none", 0)):
        s.pc_type = pc_type
        assert s._opts().pc == pc
    s.pc_type = "hypre"
    s.use_P_mat = False
    assert s._opts().pc == 0
    s.use_P_mat = True
    with pytest.warns(UserWarning, match="fieldsplit"):
        s.pc_type = "fieldsplit"
        assert s._opts().pc == 2
    s.pc_type = "hypre"
    assert s._opts().ksp_type == 0
    with pytest.warns(UserWarning, match="not symmetric"):
        s.ksp_type = "cg"
        assert s._opts().ksp_type == 1
    s.ksp_type = "gmres"
    for attr, val, exc in (("pc_type", "ilu", NotImplementedError), ("ksp_type", "bicg", NotImplementedError),
                           ("norm_type", "unpreconditioned", NotImplementedError), ("amg_form", "bogus", ValueError)):
        old = getattr(s, attr)
        setattr(s, attr, val)
        with pytest.raises(exc):
            s._opts()
        setattr(s, attr, old)


def test_ignored_settings_warn(kb, tmp_path):
    """Settings that configure hypre / an unused stimulus variant in the reference are not silently dropped: the host mirror
    warns (KNPEMIx_solver.py:38-39,72; mixed_dim_problem.py:299-304)."""
    it = BASE.replace("solver: {direct: True, output: {save_xdmf: False}}",
                      "solver: {direct: False, ksp_settings: {ksp_rtol: 1.0e-9, ksp_type: gmres, pc_type: hypre, strong_threshold: 0.5, "
                      "norm_type: preconditioned, non_zero_init_guess: True}, output: {save_xdmf: False}}")
    it = it.replace("a_syn: 5.0e-4,", "a_syn: 5.0e-4, tau_syn_rise: 1.0e-4, tau_syn_decay: 5.0e-4,")
    p = kb.ProblemKNPEMI(write(tmp_path, it), verbose=False)
    p.solver_config["view_ksp"] = False
    with pytest.warns(UserWarning) as rec:
        kb.SolverKNPEMI(p, solver_config=p.solver_config)
    msgs = " ".join(str(r.message) for r in rec)
    assert "strong_threshold" in msgs and "tau_syn" in msgs


def test_point_evaluation_and_multiple_stimulus_directions_parse(kb, tmp_path):
    txt = BASE + "point_evaluation: {ics_points: [[0.5, 0.5]], ecs_points: [[0.1, 0.1]], gamma_points: [[0.25, 0.5]]}\n" \
               + "stimulus_region: {multiple: True, direction: [x, y], range: [[0.0, 0.5], [0.2, 0.6]]}\n"
    p = kb.ProblemKNPEMI(write(tmp_path, txt), verbose=False)
    assert p.point_evaluation and np.allclose(p.ics_points, [[0.5e-6, 0.5e-6]]) and np.allclose(p.gamma_points, [[0.25e-6, 0.5e-6]])
    assert p.multiple_stimulus_directions and p.stimulus_region_directions == [0, 1]
    assert np.allclose(p.stimulus_region_range, np.array([[0.0, 0.5], [0.2, 0.6]]) * 1e-6)


def test_steady_state_initial_conditions_from_config(kb, tmp_path):
    """No `initial_conditions` block -> the membrane ODE system is integrated to rest (KNPEMIx_problem.py:224-325) with the
    compartment sizes of the mesh (utils/mixed_dim_problem.py:813-849); the constants and the fields take the result."""
    # membrane tags = cell tags as in the production configs (the membrane area is taken over dS(neuron_tags))
    txt = BASE.replace("ics_tags: [1]", "ics_tags: !range [2, 6]").replace("ecs_tags: [2]", "ecs_tags: [1]") \
              .replace("membrane_tags: [4]", "membrane_tags: !range [2, 6]")
    txt = drop(drop(drop(txt, "cell_tag_file"), "facet_tag_file"), "initial_conditions") + \
        "\nsynthetic_mesh: {kind: cell_array, dim: 2, N: 16, cells_per_dim: 2, first_tag: 2, extra_tag: 1}\n"
    p = kb.ProblemKNPEMI(write(tmp_path, txt), verbose=False)
    assert p.find_initial_conditions
    p.set_initial_conditions()
    # 2 x 2 square cells of side 0.25 in the unit square x 1e-6
    assert abs(p.vol_i_n - 0.25e-12) < 1e-24 and abs(p.vol_e - 0.75e-12) < 1e-24 and abs(p.area_g_n - 4e-6) < 1e-18
    x = np.array(p.steady_state)
    assert x.shape == (10,) and np.all(np.isfinite(x))
    assert -0.09 < x[0] < -0.05 and np.all(x[1:7] > 0) and np.all((x[7:] > 0) & (x[7:] < 1))
    assert p.phi_m_init.value == x[0] and p.K_e_init.value == x[4] and p.h_init.value == x[9]
    assert np.all(p.wh[0][3].x.array == x[0]) and np.all(p.wh[1][1].x.array == x[4]) and np.all(p.wh[0][0].x.array == x[1])
    # electroneutral exchange: what leaves the cells arrives in the ECS (the ODE conserves the ion amounts)
    for k, (ci0, ce0) in enumerate([(10.0, 145.0), (130.0, 3.0), (5.0, 134.0)]):
        before = ci0 * p.vol_i_n + ce0 * p.vol_e
        after = x[1 + 2 * k] * p.vol_i_n + x[2 + 2 * k] * p.vol_e
        assert abs(after - before) < 1e-5 * before


def test_steady_state_matches_reference_ode_classes(kb):
    """steady_state.py against tests/golden/steady_state.json, written by scripts/make_golden_steady_state.py from the
    reference's own TwoCompartment / ThreeCompartmentMembraneODESystem (utils/membrane_ODE_systems.py) on the same
    constants and compartment sizes.  Both sides integrate with rtol 1e-6 / atol 1e-8, hence the 1e-5 comparison."""
    import importlib
    import json
    ss = importlib.import_module("knp-emi-cgx_b200.steady_state")
    path = os.path.join(os.path.dirname(__file__), "golden", "steady_state.json")
    gold = json.load(open(path))
    for name, case in gold.items():
        c, g = case["constants"], case["geometry"]
        consts = dict(R=c["R"], F=c["F"], T=c["T"], C_M=c["C_M"], g_Na_bar=c["g_Na_bar"], g_K_bar=c["g_K_bar"],
                      g_leak=(c["g_Na_leak"], c["g_K_leak"], c["g_Cl_leak"]),
                      g_leak_g=(c["g_Na_leak_g"], c["g_K_leak_g"], c["g_Cl_leak_g"]), phi_rest=c["phi_rest"],
                      phi_m=c["phi_m_init"], c_i=(c["Na_i_init"], c["K_i_init"], c["Cl_i_init"]),
                      c_e=(c["Na_e_init"], c["K_e_init"], c["Cl_e_init"]), phi_m_g=c["phi_m_g_init"],
                      c_i_g=(c["Na_i_g_init"], c["K_i_g_init"], c["Cl_i_g_init"]))
        geom = dict(vol_i_n=g["vol_i_n"], vol_e=g["vol_e"], area_n=g["area_g_n"])
        if case["glia"]:
            geom.update(vol_i_g=g["vol_i_g"], area_g=g["area_g_g"])
        x, t_end, reached = ss.MembraneSteadyState(consts, geom, glia=case["glia"]).solve()
        ref = np.array(case["steady_state"])
        assert reached and x.shape == ref.shape
        assert np.abs(x - ref).max() / np.abs(ref).max() < 1e-5 and np.all(np.abs(x - ref) <= 1e-4 * np.abs(ref) + 1e-9), name
