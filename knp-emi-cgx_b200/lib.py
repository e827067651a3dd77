"""ctypes binding of libknpemi_b200.so (C ABI in include/knpemi_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a context is
created, an exception is raised.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libknpemi_b200.so")

MODEL_PASSIVE, MODEL_KIRNA, MODEL_GLIAL_CT, MODEL_NEURONAL_CT, MODEL_ATP, MODEL_HH = 1, 2, 4, 8, 16, 32

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class MeshDesc(C.Structure):
    _fields_ = [("gdim", C.c_int32), ("n_vertices", C.c_int64), ("n_owned_vertices", C.c_int64),
                ("coords", c_f64p), ("n_cells", C.c_int64), ("cell_verts", c_i32p), ("cell_tags", c_i32p),
                ("n_intra_tags", C.c_int32), ("intra_tags", c_i32p), ("extra_tag", C.c_int32),
                ("n_mfacets", C.c_int64), ("mfacet_verts", c_i32p), ("mfacet_tags", c_i32p),
                ("cell_owned", c_u8p), ("mfacet_owned", c_u8p),
                ("n_quad", C.c_int32), ("quad_bary", c_f64p), ("quad_w", c_f64p), ("degree", C.c_int32)]


class Sizes(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int64), ("nnz", C.c_int64), ("nnz_P", C.c_int64),
                ("n_own", C.c_int64 * 2), ("n_loc", C.c_int64 * 2), ("n_mverts", C.c_int64),
                ("n_mfacets", C.c_int64), ("n_cells", C.c_int64 * 2), ("max_deg", C.c_int32),
                ("max_gdeg", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("dt", C.c_double), ("F", C.c_double), ("R", C.c_double), ("T", C.c_double), ("C_M", C.c_double),
                ("phi_rest", C.c_double), ("z", C.c_double * 3), ("D", C.c_double * 3),
                ("g_Na_bar", C.c_double), ("g_K_bar", C.c_double), ("g_leak", C.c_double * 3),
                ("g_leak_g", C.c_double * 3), ("g_syn_bar", C.c_double), ("a_syn", C.c_double),
                ("T_stim", C.c_double), ("scale_stimulus", C.c_int32), ("stim_dir", C.c_int32 * 3),
                ("stim_lo", C.c_double * 3), ("stim_hi", C.c_double * 3), ("K_e_init", C.c_double),
                ("K_i_g_init", C.c_double), ("ode_substeps", C.c_int32), ("rush_larsen", C.c_int32),
                ("stim_area", C.c_double)]


class TagModels(C.Structure):
    _fields_ = [("tag", C.c_int32), ("models", C.c_uint32), ("stimulated", C.c_int32)]


class SolveOpts(C.Structure):
    _fields_ = [("rtol", C.c_double), ("max_it", C.c_int32), ("restart", C.c_int32), ("pc", C.c_int32),
                ("project_nullspace", C.c_int32), ("zero_mean_solution", C.c_int32), ("refine", C.c_int32),
                ("field_scale", C.c_double * 8), ("ksp_type", C.c_int32)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("rnorm0", C.c_double), ("rnorm", C.c_double)]


# every symbol declared in include/knpemi_b200.h (tests check that the library exports all of them)
SYMBOLS = [
    "knp_p2_emulate_host", "knp_last_error", "knp_version", "knp_launch_count", "knp_create", "knp_destroy", "knp_get_sizes", "knp_csr_dev", "knp_csr_host",
    "knp_csr_P_host", "knp_dofmap_host", "knp_mverts_host", "knp_set_params", "knp_stimulus_area_local",
    "knp_set_state", "knp_get_state", "knp_state_dev", "knp_phi_m_host", "knp_gate_step", "knp_assemble",
    "knp_assemble_P", "knp_set_source", "knp_set_dirichlet", "knp_values_dev", "knp_spmv", "knp_pc_setup", "knp_pc_apply", "knp_pc_bytes", "knp_solve", "knp_step",
    "knp_step_host", "knp_set_time", "knp_get_time", "knp_l2_norm_sq", "knp_integral", "knp_membrane_area", "knp_probe_setup", "knp_probe_eval", "knp_stimulus_current", "knp_last_timings", "knp_amg_num_levels", "knp_amg_part_levels", "knp_amg_setup_host", "knp_amg_host_level", "knp_amg_setup_device", "knp_amg_setup_was_on_device", "knp_pattern_host", "knp_edge_tables_host", "knp_rowblocks_host",
    "knp_amg_dist_sim_host", "knp_amg_dist_sim_level", "knp_amg_dist_sim_perm",
    "knp_copy", "knp_amg_level_sizes", "knp_amg_level_host", "knp_nccl_unique_id", "knp_dist_init", "knp_halo_exchange", "knp_peer_direct",
    "knp_allreduce_sum",
]


class KnpError(RuntimeError):
    code = 0


_lib = None


def load():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KnpError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C knp-emi-cgx_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.knp_last_error.restype = C.c_char_p
    lib.knp_launch_count.restype = C.c_int64
    vp = C.c_void_p
    lib.knp_create.argtypes = [C.POINTER(vp), C.POINTER(MeshDesc), C.c_int]
    lib.knp_destroy.argtypes = [vp]
    lib.knp_get_sizes.argtypes = [vp, C.POINTER(Sizes)]
    lib.knp_csr_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.knp_csr_host.argtypes = [vp, vp, vp]
    lib.knp_csr_P_host.argtypes = [vp, vp, vp]
    lib.knp_dofmap_host.argtypes = [vp, vp, vp]
    lib.knp_mverts_host.argtypes = [vp, vp]
    lib.knp_set_params.argtypes = [vp, C.POINTER(Params), C.c_int32, C.POINTER(TagModels)]
    lib.knp_stimulus_area_local.argtypes = [vp, c_f64p]
    lib.knp_set_state.argtypes = [vp, vp, vp]
    lib.knp_get_state.argtypes = [vp, vp, vp]
    lib.knp_state_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    lib.knp_phi_m_host.argtypes = [vp, vp]
    lib.knp_gate_step.argtypes = [vp, vp]
    lib.knp_assemble.argtypes = [vp, C.c_double, vp, vp, vp]
    lib.knp_assemble_P.argtypes = [vp, vp, vp]
    lib.knp_set_source.argtypes = [vp, C.c_int32, vp, vp]
    lib.knp_set_dirichlet.argtypes = [vp, C.c_int32, vp, vp]
    lib.knp_values_dev.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.knp_spmv.argtypes = [vp, vp, vp, vp, vp]
    lib.knp_pc_setup.argtypes = [vp, C.POINTER(SolveOpts)]
    lib.knp_pc_apply.argtypes = [vp, vp, vp, vp]
    lib.knp_pc_bytes.argtypes = [vp, c_f64p]
    lib.knp_solve.argtypes = [vp, vp, vp, vp, C.POINTER(SolveOpts), C.POINTER(SolveInfo), vp]
    lib.knp_step.argtypes = [vp, C.POINTER(SolveOpts), C.POINTER(SolveInfo), vp]
    lib.knp_step_host.argtypes = [vp, vp, vp, C.POINTER(SolveOpts), C.POINTER(SolveInfo)]
    lib.knp_set_time.argtypes = [vp, C.c_double, C.c_int32]
    lib.knp_get_time.argtypes = [vp, c_f64p, c_i32p]
    lib.knp_l2_norm_sq.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, c_f64p]
    lib.knp_integral.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, c_f64p]
    lib.knp_membrane_area.argtypes = [vp, C.c_int32, c_f64p]
    lib.knp_probe_setup.argtypes = [vp, C.c_int32, vp, vp, vp]
    lib.knp_probe_eval.argtypes = [vp, vp]
    lib.knp_stimulus_current.argtypes = [vp, C.c_double, c_f64p]
    lib.knp_last_timings.argtypes = [vp, vp]
    lib.knp_copy.argtypes = [vp, vp, vp, C.c_int64, C.c_int32]
    lib.knp_amg_num_levels.argtypes = [vp]
    lib.knp_amg_part_levels.argtypes = [vp, C.c_int32]
    lib.knp_rowblocks_host.argtypes = [C.c_int32, vp, C.c_int32, vp, vp]
    lib.knp_pattern_host.argtypes = [vp, c_i64p, c_i64p, vp, vp, vp, vp, vp]
    lib.knp_edge_tables_host.argtypes = [vp, c_i32p, c_i64p, c_i32p, vp, vp, vp, vp]
    lib.knp_p2_emulate_host.argtypes = [vp, C.POINTER(Params), C.c_int32, C.POINTER(TagModels), C.c_double, C.c_int32,
                                        vp, vp, vp, vp]
    lib.knp_amg_setup_host.argtypes = [C.c_int32, vp, vp, vp, C.c_double, C.c_int32, vp]
    lib.knp_amg_setup_device.argtypes = [C.c_int32, vp, vp, vp, C.c_double, C.c_int32, C.c_int32, vp]
    lib.knp_amg_setup_was_on_device.argtypes = [vp]
    lib.knp_amg_host_level.argtypes = [C.c_int32, c_i64p, c_i64p, vp, vp, vp]
    lib.knp_amg_level_sizes.argtypes = [vp, C.c_int32, c_i64p, c_i64p]
    lib.knp_amg_dist_sim_host.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, vp, C.c_double, C.c_int64, vp]
    lib.knp_amg_dist_sim_level.argtypes = [C.c_int32, C.c_int32, c_i64p, c_i64p, c_i64p, c_f64p, vp, vp, vp]
    lib.knp_amg_dist_sim_perm.argtypes = [vp]
    lib.knp_amg_level_host.argtypes = [vp, C.c_int32, vp, vp, vp]
    lib.knp_nccl_unique_id.argtypes = [C.c_char_p]
    lib.knp_dist_init.argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p, C.c_int64, C.c_int32, vp, vp, vp, vp, vp]
    lib.knp_halo_exchange.argtypes = [vp, vp, vp]
    lib.knp_peer_direct.argtypes = [vp]
    lib.knp_allreduce_sum.argtypes = [vp, vp, C.c_int32, vp]
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().knp_last_error().decode(errors="replace")
        err = KnpError(f"libknpemi_b200 error {rc}: {msg}")
        err.code = rc
        raise err


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def mesh_desc(gdim, coords, cells, cell_tags, intra_tags, extra_tag, mf_verts, mf_tags, quad_bary, quad_w,
              n_owned_vertices=None, cell_owned=None, mfacet_owned=None, degree=1):
    """knp_mesh_desc for numpy arrays; returns (descriptor, dict of the arrays that must stay alive).  degree = 2: the
    arrays describe the P2 node mesh (mesh.py::p2_node_mesh): nodes for vertices, 6 / 10 nodes per cell, 3 / 6 per facet."""
    nt = gdim if degree != 2 else gdim * (gdim + 1) // 2
    if degree == 2 and np.asarray(cells).shape[1] != (gdim + 1) * (gdim + 2) // 2:
        raise KnpError("degree = 2 expects the P2 node mesh (mesh.p2_node_mesh)")
    k = dict(
        coords=np.ascontiguousarray(coords, np.float64),
        cells=np.ascontiguousarray(cells, np.int32),
        cell_tags=np.ascontiguousarray(cell_tags, np.int32),
        intra=np.ascontiguousarray(intra_tags, np.int32),
        mfv=np.ascontiguousarray(mf_verts, np.int32).reshape(-1, nt),
        mft=np.ascontiguousarray(mf_tags, np.int32),
        qb=np.ascontiguousarray(quad_bary, np.float64),
        qw=np.ascontiguousarray(quad_w, np.float64),
        co=None if cell_owned is None else np.ascontiguousarray(cell_owned, np.uint8),
        fo=None if mfacet_owned is None else np.ascontiguousarray(mfacet_owned, np.uint8),
    )
    d = MeshDesc()
    d.gdim = gdim
    d.n_vertices = k["coords"].shape[0]
    d.n_owned_vertices = d.n_vertices if n_owned_vertices is None else int(n_owned_vertices)
    d.coords = k["coords"].ctypes.data_as(c_f64p)
    d.n_cells = k["cells"].shape[0]
    d.cell_verts = k["cells"].ctypes.data_as(c_i32p)
    d.cell_tags = k["cell_tags"].ctypes.data_as(c_i32p)
    d.n_intra_tags = k["intra"].size
    d.intra_tags = k["intra"].ctypes.data_as(c_i32p)
    d.extra_tag = int(extra_tag)
    d.n_mfacets = k["mfv"].shape[0]
    d.mfacet_verts = k["mfv"].ctypes.data_as(c_i32p)
    d.mfacet_tags = k["mft"].ctypes.data_as(c_i32p)
    d.cell_owned = None if k["co"] is None else k["co"].ctypes.data_as(c_u8p)
    d.mfacet_owned = None if k["fo"] is None else k["fo"].ctypes.data_as(c_u8p)
    d.n_quad = k["qw"].size
    d.quad_bary = k["qb"].ctypes.data_as(c_f64p)
    d.quad_w = k["qw"].ctypes.data_as(c_f64p)
    d.degree = int(degree)
    return d, k


def pattern_host(gdim, coords, cells, cell_tags, intra_tags, extra_tag, mf_verts, mf_tags, quad_bary, quad_w, **kw):
    """(indptr, indices, dof_vert_i, dof_vert_e) of the system matrix as knp_create lays it out; host only, no GPU."""
    lib = load()
    d, keep = mesh_desc(gdim, coords, cells, cell_tags, intra_tags, extra_tag, mf_verts, mf_tags, quad_bary, quad_w, **kw)
    n, nnz = C.c_int64(), C.c_int64()
    own = (C.c_int32 * 4)()
    check(lib.knp_pattern_host(C.byref(d), C.byref(n), C.byref(nnz), own, None, None, None, None))
    indptr, indices = np.empty(n.value + 1, np.int32), np.empty(nnz.value, np.int32)
    vi, ve = np.empty(own[2], np.int32), np.empty(own[3], np.int32)      # local dofs: owned first, then ghosts
    check(lib.knp_pattern_host(C.byref(d), None, None, None, _ptr(indptr), _ptr(indices), _ptr(vi), _ptr(ve)))
    return indptr, indices, vi, ve


def edge_tables_host(gdim, coords, cells, cell_tags, intra_tags, extra_tag, mf_verts, mf_tags, quad_bary, quad_w, **kw):
    """Lane-group tables of the edge-lane row kernel (host only): dict(lgG, adjG[W, G], hitG[W, G, words], meta[W, 2],
    node_x[n_loc0 + n_loc1, gdim], n_own_loc) or None when the mesh does not fit the tables."""
    lib = load()
    d, keep = mesh_desc(gdim, coords, cells, cell_tags, intra_tags, extra_tag, mf_verts, mf_tags, quad_bary, quad_w, **kw)
    lg, ok, W = C.c_int32(), C.c_int32(), C.c_int64()
    check(lib.knp_edge_tables_host(C.byref(d), C.byref(lg), C.byref(W), C.byref(ok), None, None, None, None))
    if not ok.value:
        return None
    n = C.c_int64()
    own = (C.c_int32 * 4)()
    check(lib.knp_pattern_host(C.byref(d), C.byref(n), None, own, None, None, None, None))
    G, hw = 1 << lg.value, (1 if gdim == 2 else 4)
    adjG, hitG = np.empty((W.value, G), np.int32), np.empty((W.value, G, hw), np.uint32)
    meta, x = np.empty((W.value, 2), np.int32), np.empty((own[2] + own[3], gdim), np.float64)
    check(lib.knp_edge_tables_host(C.byref(d), None, None, None, _ptr(adjG), _ptr(hitG), _ptr(meta), _ptr(x)))
    return dict(lgG=lg.value, adjG=adjG, hitG=hitG, meta=meta, node_x=x, n_own_loc=tuple(own))


def p2_emulate_host(params: Params, tag_models, t, mode, u, gates, gdim, coords, cells, cell_tags, intra_tags, extra_tag,
                    mf_verts, mf_tags, quad_bary, quad_w, **kw):
    """TEST INFRASTRUCTURE (no GPU): one P2 assembly on the CPU with the functions the P2 kernels run per thread
    (knp_p2_emulate_host).  Returns (indptr, indices, values, b) for mode 0 and the values of P for mode 1."""
    lib = load()
    d, keep = mesh_desc(gdim, coords, cells, cell_tags, intra_tags, extra_tag, mf_verts, mf_tags, quad_bary, quad_w, degree=2, **kw)
    n, nnz = C.c_int64(), C.c_int64()
    check(lib.knp_pattern_host(C.byref(d), C.byref(n), C.byref(nnz), None, None, None, None, None))
    arr = (TagModels * max(1, len(tag_models)))(*tag_models)
    u = np.ascontiguousarray(u, np.float64)
    gates = None if gates is None else np.ascontiguousarray(gates, np.float64)
    if mode == 0:
        indptr, indices = np.empty(n.value + 1, np.int32), np.empty(nnz.value, np.int32)
        check(lib.knp_pattern_host(C.byref(d), None, None, None, _ptr(indptr), _ptr(indices), None, None))
        vals, b = np.full(nnz.value, np.nan), np.full(n.value, np.nan)
        check(lib.knp_p2_emulate_host(C.byref(d), C.byref(params), len(tag_models), arr, float(t), 0, _ptr(u), _ptr(gates),
                                      _ptr(vals), _ptr(b)))
        return indptr, indices, vals, b
    vals = np.full(nnz.value, np.nan)           # nnz_P <= nnz
    check(lib.knp_p2_emulate_host(C.byref(d), C.byref(params), len(tag_models), arr, float(t), 1, _ptr(u), None, _ptr(vals), None))
    return vals


class Context:
    """Thin object wrapper around a knp_ctx*; numpy in, numpy out; device pointers as ints."""

    def __init__(self, gdim, coords, cells, cell_tags, intra_tags, extra_tag, mf_verts, mf_tags, quad_bary, quad_w,
                 n_owned_vertices=None, cell_owned=None, mfacet_owned=None, device=0, degree=1):
        lib = load()
        self._lib = lib
        d, k = mesh_desc(gdim, coords, cells, cell_tags, intra_tags, extra_tag, mf_verts, mf_tags, quad_bary, quad_w,
                         n_owned_vertices, cell_owned, mfacet_owned, degree)
        self.degree = degree
        h = C.c_void_p()
        check(lib.knp_create(C.byref(h), C.byref(d), device))
        self.h = h
        s = Sizes()
        check(lib.knp_get_sizes(self.h, C.byref(s)))
        self.sizes = s
        self.n_rows, self.n_cols, self.nnz, self.nnz_P = s.n_rows, s.n_cols, s.nnz, s.nnz_P
        self.n_own = (s.n_own[0], s.n_own[1])
        self.n_loc = (s.n_loc[0], s.n_loc[1])
        self.n_mverts = s.n_mverts
        self.gdim = gdim

    def close(self):
        if getattr(self, "h", None):
            self._lib.knp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- structure ----
    def csr(self):
        indptr = np.empty(self.n_rows + 1, np.int32)
        indices = np.empty(self.nnz, np.int32)
        check(self._lib.knp_csr_host(self.h, _ptr(indptr), _ptr(indices)))
        return indptr, indices

    def csr_P(self):
        indptr = np.empty(self.n_rows + 1, np.int32)
        indices = np.empty(self.nnz_P, np.int32)
        check(self._lib.knp_csr_P_host(self.h, _ptr(indptr), _ptr(indices)))
        return indptr, indices

    def dofmaps(self):
        vi = np.empty(self.n_loc[0], np.int32)
        ve = np.empty(self.n_loc[1], np.int32)
        check(self._lib.knp_dofmap_host(self.h, _ptr(vi), _ptr(ve)))
        return vi, ve

    def mverts(self):
        v = np.empty(self.n_mverts, np.int32)
        check(self._lib.knp_mverts_host(self.h, _ptr(v)))
        return v

    # ---- parameters / state ----
    def set_params(self, params: Params, tag_models):
        arr = (TagModels * max(1, len(tag_models)))()
        for i, (tag, models, stim) in enumerate(tag_models):
            arr[i].tag, arr[i].models, arr[i].stimulated = int(tag), int(models), int(bool(stim))
        check(self._lib.knp_set_params(self.h, C.byref(params), len(tag_models), arr))

    def stimulus_area_local(self):
        out = C.c_double()
        check(self._lib.knp_stimulus_area_local(self.h, C.byref(out)))
        return out.value

    def set_state(self, u=None, gates=None):
        u = None if u is None else np.ascontiguousarray(u, np.float64)
        gates = None if gates is None else np.ascontiguousarray(gates, np.float64)
        if u is not None:
            assert u.size == self.n_cols
        if gates is not None:
            assert gates.size == 3 * self.n_mverts
        check(self._lib.knp_set_state(self.h, _ptr(u), _ptr(gates)))

    def get_state(self):
        u = np.empty(self.n_cols, np.float64)
        g = np.empty((3, self.n_mverts), np.float64)
        check(self._lib.knp_get_state(self.h, _ptr(u), _ptr(g)))
        return u, g

    def phi_m(self):
        out = np.empty(self.n_mverts, np.float64)
        check(self._lib.knp_phi_m_host(self.h, _ptr(out)))
        return out

    def dev_ptrs(self):
        a, b, p, x = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(self._lib.knp_values_dev(self.h, C.byref(a), C.byref(b), C.byref(p), C.byref(x)))
        g = C.c_void_p()
        u = C.c_void_p()
        check(self._lib.knp_state_dev(self.h, C.byref(u), C.byref(g)))
        return dict(A=a.value, b=b.value, P=p.value, u=u.value, gates=g.value)

    # ---- hot path ----
    def gate_step(self, stream=None):
        check(self._lib.knp_gate_step(self.h, stream))

    def assemble(self, t, A_ptr=None, b_ptr=None, stream=None):
        check(self._lib.knp_assemble(self.h, float(t), A_ptr, b_ptr, stream))

    def assemble_P(self, P_ptr=None, stream=None):
        check(self._lib.knp_assemble_P(self.h, P_ptr, stream))

    def set_source(self, rows, vals):
        """b[rows] += vals in every assembly (ion-injection terms); empty arrays clear."""
        rows = np.ascontiguousarray(rows, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        assert rows.shape == vals.shape and np.unique(rows).size == rows.size
        check(self._lib.knp_set_source(self.h, int(rows.size), _ptr(rows), _ptr(vals)))

    def set_dirichlet(self, cols, vals):
        """Constrained dofs (column layout, owned and ghost) and their values; empty arrays clear (knp_set_dirichlet)."""
        cols = np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        assert cols.shape == vals.shape
        check(self._lib.knp_set_dirichlet(self.h, int(cols.size), _ptr(cols), _ptr(vals)))

    def spmv(self, x_ptr, y_ptr, A_ptr=None, stream=None):
        check(self._lib.knp_spmv(self.h, A_ptr, x_ptr, y_ptr, stream))

    def pc_setup(self, opts: SolveOpts):
        check(self._lib.knp_pc_setup(self.h, C.byref(opts)))

    def pc_apply(self, r_ptr, z_ptr, stream=None):
        check(self._lib.knp_pc_apply(self.h, r_ptr, z_ptr, stream))

    def pc_bytes(self):
        out = C.c_double()
        check(self._lib.knp_pc_bytes(self.h, C.byref(out)))
        return out.value

    def solve(self, opts: SolveOpts, A_ptr=None, b_ptr=None, x_ptr=None, stream=None):
        info = SolveInfo()
        check(self._lib.knp_solve(self.h, A_ptr, b_ptr, x_ptr, C.byref(opts), C.byref(info), stream))
        return info

    def step(self, opts: SolveOpts, stream=None, raise_on_nonconvergence=True):
        """One timestep.  A Krylov solve that hits max_it makes knp_step return KNP_E_NOCONV (the state holds the last
        iterate); with raise_on_nonconvergence=False that case is returned like PETSc's KSP does it (info.converged == 0)."""
        info = SolveInfo()
        rc = self._lib.knp_step(self.h, C.byref(opts), C.byref(info), stream)
        if rc == -3 and not raise_on_nonconvergence and info.iterations >= opts.max_it:
            return info
        check(rc)
        return info

    def step_host(self, u_host, gates_host, opts: SolveOpts):
        info = SolveInfo()
        check(self._lib.knp_step_host(self.h, _ptr(u_host), _ptr(gates_host), C.byref(opts), C.byref(info)))
        return info

    def set_time(self, t, step_index):
        check(self._lib.knp_set_time(self.h, float(t), int(step_index)))

    def get_time(self):
        t, i = C.c_double(), C.c_int32()
        check(self._lib.knp_get_time(self.h, C.byref(t), C.byref(i)))
        return t.value, i.value

    def integral(self, subdomain, field, tags, power=1):
        """This rank's integral of u^power over the owned cells of `subdomain` tagged `tags` (power 0: their measure)."""
        tags = np.ascontiguousarray(np.atleast_1d(tags), np.int32)
        out = C.c_double()
        check(self._lib.knp_integral(self.h, subdomain, field, power, tags.size, _ptr(tags), C.byref(out)))
        return out.value

    def stimulus_current(self, t):
        """This rank's part of int stim_expr dS(stimulus_tags) at time t from the state on the device."""
        out = C.c_double()
        check(self._lib.knp_stimulus_current(self.h, float(t), C.byref(out)))
        return out.value

    def probe_setup(self, ptr, cols, weights):
        ptr, cols = np.ascontiguousarray(ptr, np.int32), np.ascontiguousarray(cols, np.int32)
        weights = np.ascontiguousarray(weights, np.float64)
        self._n_probe = ptr.size - 1
        check(self._lib.knp_probe_setup(self.h, self._n_probe, _ptr(ptr), _ptr(cols), _ptr(weights)))

    def probe_eval(self):
        out = np.zeros(getattr(self, "_n_probe", 0))
        check(self._lib.knp_probe_eval(self.h, _ptr(out)))
        return out

    def membrane_area(self, tag):
        out = C.c_double()
        check(self._lib.knp_membrane_area(self.h, int(tag), C.byref(out)))
        return out.value

    def l2_norm_sq(self, subdomain, field, tags):
        tags = np.ascontiguousarray(np.atleast_1d(tags), np.int32)
        out = C.c_double()
        check(self._lib.knp_l2_norm_sq(self.h, subdomain, field, tags.size, _ptr(tags), C.byref(out)))
        return out.value

    def last_timings(self):
        ms = np.zeros(5)
        check(self._lib.knp_last_timings(self.h, _ptr(ms)))
        return dict(gate=ms[0], facet=ms[1], rows=ms[2], solve=ms[3], total=ms[4])

    def to_host(self, dev_ptr, count, dtype=np.float64):
        out = np.empty(count, dtype)
        check(self._lib.knp_copy(self.h, _ptr(out), dev_ptr, out.nbytes, 2))
        return out

    def to_dev(self, dev_ptr, arr):
        arr = np.ascontiguousarray(arr)
        check(self._lib.knp_copy(self.h, dev_ptr, _ptr(arr), arr.nbytes, 1))

    def values_host(self):
        """(A values, b, P values) copied from the context's own buffers."""
        d = self.dev_ptrs()
        return (self.to_host(d["A"], self.nnz), self.to_host(d["b"], self.n_rows), self.to_host(d["P"], self.nnz_P))

    def amg_levels(self, part=None):
        """Level operators as scipy CSR; part=None: all hierarchies in order, else only hierarchy `part`."""
        import scipy.sparse as sp
        out = []
        n0 = self._lib.knp_amg_part_levels(self.h, 0)
        rng = range(self._lib.knp_amg_num_levels(self.h)) if part is None else (
            range(n0) if part == 0 else range(n0, n0 + self._lib.knp_amg_part_levels(self.h, part)))
        for l in rng:
            n, nnz = C.c_int64(), C.c_int64()
            check(self._lib.knp_amg_level_sizes(self.h, l, C.byref(n), C.byref(nnz)))
            ip = np.empty(n.value + 1, np.int32)
            ix = np.empty(nnz.value, np.int32)
            va = np.empty(nnz.value, np.float64)
            check(self._lib.knp_amg_level_host(self.h, l, _ptr(ip), _ptr(ix), _ptr(va)))
            out.append(sp.csr_matrix((va, ix, ip), shape=(n.value, n.value)))
        return out

    # ---- distributed ----
    def dist_init(self, rank, nranks, unique_id, n_phi_global, peers, send_ptr, send_cols, recv_ptr, recv_cols):
        peers = np.ascontiguousarray(peers, np.int32)
        send_ptr = np.ascontiguousarray(send_ptr, np.int64)
        recv_ptr = np.ascontiguousarray(recv_ptr, np.int64)
        send_cols = np.ascontiguousarray(send_cols, np.int32)
        recv_cols = np.ascontiguousarray(recv_cols, np.int32)
        check(self._lib.knp_dist_init(self.h, rank, nranks, unique_id, int(n_phi_global), peers.size, _ptr(peers),
                                      _ptr(send_ptr), _ptr(send_cols), _ptr(recv_ptr), _ptr(recv_cols)))

    def peer_direct(self):
        return bool(self._lib.knp_peer_direct(self.h))

    def halo_exchange(self, x_ptr=None, stream=None):
        check(self._lib.knp_halo_exchange(self.h, x_ptr, stream))


def launch_count():
    return int(load().knp_launch_count())


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    check(load().knp_nccl_unique_id(buf))
    return buf.raw


def amg_setup_host(A, theta=0.08, coarse_size=600, device=None):
    """Level operators (scipy CSR) of the library's smoothed-aggregation setup for a scipy CSR matrix: the host form
    (amg_setup.cpp, no GPU needed) or, with device = a CUDA device index, the device form (amg_device.cu)."""
    import scipy.sparse as sp
    lib = load()
    A = sp.csr_matrix(A)
    A.sort_indices()
    ip = np.ascontiguousarray(A.indptr, np.int32)
    ix = np.ascontiguousarray(A.indices, np.int32)
    va = np.ascontiguousarray(A.data, np.float64)
    nl = C.c_int32()
    if device is None:
        check(lib.knp_amg_setup_host(A.shape[0], _ptr(ip), _ptr(ix), _ptr(va), theta, coarse_size, C.byref(nl)))
    else:
        check(lib.knp_amg_setup_device(A.shape[0], _ptr(ip), _ptr(ix), _ptr(va), theta, coarse_size, int(device), C.byref(nl)))
    out = []
    for l in range(nl.value):
        n, nnz = C.c_int64(), C.c_int64()
        check(lib.knp_amg_host_level(l, C.byref(n), C.byref(nnz), None, None, None))
        lp, li, lv = np.empty(n.value + 1, np.int32), np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float64)
        check(lib.knp_amg_host_level(l, None, None, _ptr(lp), _ptr(li), _ptr(lv)))
        out.append(sp.csr_matrix((lv, li, lp), shape=(n.value, n.value)))
    return out


def amg_dist_sim_host(A, owner, nranks, theta=0.08, repl_threshold=600):
    """Row-distributed hierarchy setup on `nranks` simulated ranks (host only).  Returns (As, Ps, rhos, perm0): global level
    operators (the last one is the replicated level), prolongators between them, the smoother bounds of the distributed
    levels and the level-0 numbering (new index -> index in A)."""
    import scipy.sparse as sp
    lib = load()
    A = sp.csr_matrix(A)
    A.sort_indices()
    ip, ix = np.ascontiguousarray(A.indptr, np.int32), np.ascontiguousarray(A.indices, np.int32)
    va, ow = np.ascontiguousarray(A.data, np.float64), np.ascontiguousarray(owner, np.int32)
    nl = C.c_int32()
    check(lib.knp_amg_dist_sim_host(int(nranks), A.shape[0], _ptr(ip), _ptr(ix), _ptr(va), _ptr(ow), theta,
                                    int(repl_threshold), C.byref(nl)))

    def level(l, which):
        nr, nc, nnz, rho = C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
        check(lib.knp_amg_dist_sim_level(l, which, C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(rho), None, None, None))
        lp, li, lv = np.empty(nr.value + 1, np.int32), np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float64)
        check(lib.knp_amg_dist_sim_level(l, which, None, None, None, None, _ptr(lp), _ptr(li), _ptr(lv)))
        return sp.csr_matrix((lv, li, lp), shape=(nr.value, nc.value)), rho.value

    As = [level(l, 0) for l in range(nl.value)]
    Ps = [level(l, 1)[0] for l in range(nl.value - 1)]
    perm = np.empty(A.shape[0], np.int32)
    check(lib.knp_amg_dist_sim_perm(_ptr(perm)))
    return [a for a, _ in As], Ps, [r for _, r in As[:-1]], perm


def rowblocks_host(indptr):
    """Row blocks of the streaming SpMV for a CSR row-pointer array: (n_blocks, 4) int32 or None (fallback kernel)."""
    lib = load()
    ip = np.ascontiguousarray(indptr, np.int32)
    n = ip.size - 1
    nb = C.c_int32()
    check(lib.knp_rowblocks_host(n, _ptr(ip), 0, None, C.byref(nb)))
    if nb.value <= 0:
        return None
    out = np.empty((nb.value, 4), np.int32)
    check(lib.knp_rowblocks_host(n, _ptr(ip), nb.value, _ptr(out), C.byref(nb)))
    return out
