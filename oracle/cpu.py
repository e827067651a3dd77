"""ctypes wrapper of oracle/libknpemi_cpu.so, the C++/OpenMP restatement of the oracle's timestep (TEST / BASELINE
INFRASTRUCTURE: used by tests/test_cpu_baseline.py and by bench.py's cpu_baseline / --impl reference legs only).

It runs the same algorithm as oracle/knpemi.py + oracle/amg.py::SchurPC (which restate the reference's time-loop body,
KNPEMIx_solver.py:365-468, with the product's own preconditioner in place of hypre) on all host cores, so that the CPU
baseline is a measurement at the actual configuration size instead of a scaled single-core numpy sample."""
import ctypes as C
import os
import subprocess
import numpy as np

from .quadrature import facet_rule

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libknpemi_cpu.so")
FLAGS = {"Passive": 1, "KirNa": 2, "GlialCT": 4, "NeuronalCT": 8, "ATP": 16, "HH": 32}

_f64p, _i32p = C.POINTER(C.c_double), C.POINTER(C.c_int32)


class _Params(C.Structure):
    _fields_ = [("dt", C.c_double), ("F", C.c_double), ("R", C.c_double), ("T", C.c_double), ("C_M", C.c_double),
                ("phi_rest", C.c_double), ("z", C.c_double * 3), ("D", C.c_double * 3), ("g_Na_bar", C.c_double),
                ("g_K_bar", C.c_double), ("g_leak", C.c_double * 3), ("g_leak_g", C.c_double * 3), ("g_syn_bar", C.c_double),
                ("a_syn", C.c_double), ("T_stim", C.c_double), ("scale_stimulus", C.c_int32), ("stim_dir", C.c_int32 * 3),
                ("stim_lo", C.c_double * 3), ("stim_hi", C.c_double * 3), ("K_e_init", C.c_double), ("K_i_g_init", C.c_double),
                ("ode_substeps", C.c_int32), ("rush_larsen", C.c_int32)]


class _Mesh(C.Structure):
    _fields_ = [("gdim", C.c_int32), ("ns", C.c_int32 * 2), ("x", _f64p * 2), ("n_cells", C.c_int64 * 2), ("cells", _i32p * 2),
                ("n_mv", C.c_int32), ("n_mf", C.c_int32), ("nq", C.c_int32), ("mv_node", _i32p * 2), ("mf_mv", _i32p),
                ("mf_models", C.POINTER(C.c_uint32)), ("mf_stim", C.POINTER(C.c_uint8)), ("qb", _f64p), ("qw", _f64p),
                ("indptr", _i32p), ("indices", _i32p)]


def build():
    subprocess.run(["make", "-C", os.path.join(HERE, "cpu_baseline")], check=True, capture_output=True)


def load():
    if not os.path.exists(LIB):
        build()
    lib = C.CDLL(LIB)
    vp = C.c_void_p
    lib.kcpu_create.restype = vp
    lib.kcpu_create.argtypes = [C.POINTER(_Mesh), C.POINTER(_Params), C.c_int32]
    lib.kcpu_destroy.argtypes = [vp]
    lib.kcpu_nnz.restype = C.c_int64
    lib.kcpu_nnz.argtypes = [vp]
    lib.kcpu_error.restype = C.c_char_p
    lib.kcpu_error.argtypes = [vp]
    lib.kcpu_set_state.argtypes = [vp, vp, vp]
    lib.kcpu_get_state.argtypes = [vp, vp, vp]
    lib.kcpu_assemble.argtypes = [vp, C.c_double, vp, vp]
    lib.kcpu_gate_update.argtypes = [vp]
    lib.kcpu_pc_setup.argtypes = [vp]
    lib.kcpu_pc_apply.argtypes = [vp, vp, vp]
    lib.kcpu_spmv.argtypes = [vp, vp, vp]
    lib.kcpu_step.argtypes = [vp, C.c_double, C.c_int32, _i32p, vp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CpuBaseline:
    """One problem instance: mesh arrays (x, cells, cell_tags, membrane facets with tags), OracleParams and the model list as
    for KNPEMIOracle.  The library builds the CSR pattern of the system matrix itself (contract ordering, 64-bit row
    pointers); `pattern` = (indptr, indices) is an optional expectation -- tests pass the oracle's CSR and construction
    fails if the two differ in a single entry."""

    def __init__(self, gdim, x, cells, cell_tags, mf_verts, mf_tags, params, models, pattern=None, restart=30):
        self.lib = load()
        p = params
        nv = x.shape[0]
        is_in0 = np.isin(cell_tags, np.asarray(p.intra_tags))
        # restricted dof sets = vertices of the subdomain's cells, ascending (DofMapRestriction, KNPEMIx_problem.py:85-89)
        self.S = [np.unique(cells[is_in0].ravel()).astype(np.int64), np.unique(cells[cell_tags == p.extra_tag].ravel()).astype(np.int64)]
        indptr, indices = (None, None) if pattern is None else pattern[:2]
        r = []
        for S in self.S:
            a = np.full(nv, -1, np.int64)
            a[S] = np.arange(S.size)
            r.append(a)
        is_in = np.isin(cell_tags, np.asarray(p.intra_tags))
        is_ex = cell_tags == p.extra_tag
        k = self._keep = {}
        k["x"] = [np.ascontiguousarray(x[S], np.float64) for S in self.S]
        k["cells"] = [np.ascontiguousarray(r[0][cells[is_in]], np.int32), np.ascontiguousarray(r[1][cells[is_ex]], np.int32)]
        mverts = np.unique(mf_verts.ravel())
        self.mverts = mverts
        k["mv_node"] = [np.ascontiguousarray(r[s][mverts], np.int32) for s in range(2)]
        k["mf_mv"] = np.ascontiguousarray(np.searchsorted(mverts, mf_verts), np.int32)
        table = {}
        for name, tags in models:
            for t in (p.membrane_tags if tags is None else tags):
                table[int(t)] = table.get(int(t), 0) | FLAGS[name]
        k["mf_models"] = np.ascontiguousarray([table.get(int(t), 0) for t in mf_tags], np.uint32)
        k["mf_stim"] = np.ascontiguousarray(np.isin(mf_tags, np.asarray(p.stimulus_tags)), np.uint8)
        qb, qw = facet_rule(gdim)
        k["qb"], k["qw"] = np.ascontiguousarray(qb, np.float64), np.ascontiguousarray(qw, np.float64)
        if indptr is not None:
            k["indptr"], k["indices"] = np.ascontiguousarray(indptr, np.int32), np.ascontiguousarray(indices, np.int32)
        m = _Mesh()
        m.gdim = gdim
        for s in range(2):
            m.ns[s] = self.S[s].size
            m.x[s] = k["x"][s].ctypes.data_as(_f64p)
            m.n_cells[s] = k["cells"][s].shape[0]
            m.cells[s] = k["cells"][s].ctypes.data_as(_i32p)
            m.mv_node[s] = k["mv_node"][s].ctypes.data_as(_i32p)
        m.n_mv, m.n_mf, m.nq = mverts.size, mf_verts.shape[0], qw.size
        m.mf_mv = k["mf_mv"].ctypes.data_as(_i32p)
        m.mf_models = k["mf_models"].ctypes.data_as(C.POINTER(C.c_uint32))
        m.mf_stim = k["mf_stim"].ctypes.data_as(C.POINTER(C.c_uint8))
        m.qb, m.qw = k["qb"].ctypes.data_as(_f64p), k["qw"].ctypes.data_as(_f64p)
        if indptr is not None:
            m.indptr, m.indices = k["indptr"].ctypes.data_as(_i32p), k["indices"].ctypes.data_as(_i32p)
        P = _Params()
        P.dt, P.F, P.R, P.T, P.C_M, P.phi_rest = p.dt, p.F, p.R, p.T, p.C_M, p.phi_rest
        for i in range(3):
            P.z[i], P.D[i], P.g_leak[i], P.g_leak_g[i] = p.z[i], p.D[i], p.g_leak[i], p.g_leak_g[i]
        P.g_Na_bar, P.g_K_bar, P.g_syn_bar, P.a_syn, P.T_stim = p.g_Na_bar, p.g_K_bar, p.g_syn_bar, p.a_syn, p.T_stim
        P.scale_stimulus = int(p.scale_stimulus)
        for i in range(3):
            P.stim_dir[i] = -1
        if p.stimulus_region is not None:
            regions = p.stimulus_region if isinstance(p.stimulus_region[0], (tuple, list)) else [p.stimulus_region]
            for i, (d, lo, hi) in enumerate(regions):
                P.stim_dir[i], P.stim_lo[i], P.stim_hi[i] = d, lo, hi
        P.K_e_init, P.K_i_g_init = p.c_e_init[1], p.c_i_g_init[1]
        P.ode_substeps, P.rush_larsen = p.ode_substeps, int(p.rush_larsen)
        self.any_hh = any(nm == "HH" for nm, _ in models)
        self.n = 4 * (self.S[0].size + self.S[1].size)
        self.n_mv = mverts.size
        self.h = self.lib.kcpu_create(C.byref(m), C.byref(P), restart)
        if not self.h:
            raise RuntimeError("CPU baseline: the CSR pattern built by the library differs from the expected one")
        self.nnz = int(self.lib.kcpu_nnz(self.h))

    def close(self):
        if self.h:
            self.lib.kcpu_destroy(self.h)
            self.h = None

    def set_state(self, u, gates_on_mverts):
        u = np.ascontiguousarray(u, np.float64)
        g = np.ascontiguousarray(gates_on_mverts, np.float64)
        assert u.size == self.n and g.size == 3 * self.n_mv
        self.lib.kcpu_set_state(self.h, _p(u), _p(g))

    def get_state(self):
        u, g = np.empty(self.n), np.empty((3, self.n_mv))
        self.lib.kcpu_get_state(self.h, _p(u), _p(g))
        return u, g

    def assemble(self, t):
        A, b = np.empty(self.nnz), np.empty(self.n)
        self.lib.kcpu_assemble(self.h, float(t), _p(A), _p(b))
        return A, b

    def pc_setup(self):
        if self.lib.kcpu_pc_setup(self.h) != 0:
            raise RuntimeError(self.lib.kcpu_error(self.h).decode())

    def pc_apply(self, r):
        r = np.ascontiguousarray(r, np.float64)
        z = np.empty(self.n)
        self.lib.kcpu_pc_apply(self.h, _p(r), _p(z))
        return z

    def step(self, rtol=1e-9):
        its = C.c_int32()
        ms = np.zeros(3)
        rc = self.lib.kcpu_step(self.h, float(rtol), int(self.any_hh), C.byref(its), _p(ms))
        if rc != 0:
            raise RuntimeError(f"CPU baseline: GMRES failed ({rc})")
        return its.value, dict(assembly=ms[0], solve=ms[1], total=ms[2])

    @property
    def threads(self):
        return int(self.lib.kcpu_threads())
