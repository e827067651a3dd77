"""ProblemKNPEMI: host-side mirror of the reference's problem classes.

Mirrors src/CGx/utils/mixed_dim_problem.py (MixedDimensionalProblem: YAML schema :86-374, tag parsing
:376-433, init_ionic_models :435-465) and src/CGx/KNPEMI/KNPEMIx_problem.py (ProblemKNPEMI: spaces and
restrictions :28-94, initial conditions :220-452, constants :909-981, class defaults :983-997) with the same
public names, argument meaning and error behaviour.  What differs is *where the forms live*: instead of
building UFL forms that FFCx JIT-compiles, ``setup_variational_form`` creates a device context in
libknpemi_b200.so whose CUDA kernels evaluate exactly those forms (csrc/assembly.cu).
"""
import os
import re
import time
import collections.abc
import numpy as np
import yaml

from . import lib as _lib
from . import mesh as _mesh
from .comm import Comm, MPI


def range_constructor(loader, node):
    """!range [a, b] -> list(range(a, b))   (src/CGx/utils/misc.py:33-37)."""
    return list(range(*loader.construct_sequence(node)))


def flatten_list(input_list):
    return [item for sub in input_list for item in (sub if isinstance(sub, tuple) else [sub])]


class Constant:
    """Stand-in for dolfinx.fem.Constant: a mutable scalar with a ``.value``."""

    def __init__(self, value):
        # PyYAML reads `1e-9` (no dot) as a string; dolfinx.fem.Constant converts through numpy, and so do we
        self.value = float(value) if isinstance(value, str) else value

    def __float__(self):
        return float(self.value)

    def __repr__(self):
        return f"Constant({self.value})"


class _Vector:
    def __init__(self, fn):
        self._fn = fn

    @property
    def array(self):
        self._fn._problem._sync_host()
        return self._fn._data

    def scatter_forward(self):
        pass


class Function:
    """Stand-in for dolfinx.fem.Function on the whole (local) mesh: ``.x.array`` is a numpy view that is
    refreshed from the device state on access."""

    def __init__(self, problem, n, name=""):
        self._problem = problem
        self._data = np.zeros(n)
        self.name = name
        self.x = _Vector(self)


class _IndexMap:
    def __init__(self, size_local, num_ghosts=0):
        self.size_local = size_local
        self.num_ghosts = num_ghosts


class _Restriction:
    """Stand-in for multiphenicsx.fem.DofMapRestriction (KNPEMIx_problem.py:88-89)."""

    def __init__(self, dofs, n_owned):
        self.dofs = dofs
        self.index_map = _IndexMap(int(n_owned), int(dofs.size - n_owned))


class Measure:
    """problem.dx(tag): only used to take norms in the reference drivers (tests/KNPEMI/*.py:45-51)."""

    def __init__(self, problem, tags=None):
        self.problem = problem
        self.tags = tags

    def __call__(self, tags):
        return Measure(self.problem, tuple(np.atleast_1d(tags).tolist()))


class ProblemKNPEMI:
    # Class settings (KNPEMIx_problem.py:983-997)
    mesh_conversion_factor = 1.0
    fem_order = 1
    MMS_test = False
    dirichlet_bcs = False
    pin_ecs_potential = False
    # physical defaults used when the config gives no physical_constants (mixed_dim_problem.py:193-195)
    T_value = R_value = F_value = psi_value = 1.0

    def __init__(self, config_file: str, comm: Comm = None, device: int = None, verbose: bool = True):
        tic = time.perf_counter()
        self.comm = comm if comm is not None else Comm()
        self.verbose = verbose
        self.device = device
        self._print("Reading input data from " + str(config_file))
        self.read_config_file(config_file=str(config_file))
        self.setup_domain()
        self.t = Constant(0.0)
        self.dt = Constant(float(self.dt))
        self.setup_constants()
        self.setup_spaces()
        self.init()
        self.setup_boundary_conditions()
        if self.source_terms == "ion_injection":
            self.setup_source_terms()
        self.ionic_models = []
        self.gating_variables = False
        self.ode_substeps, self.rush_larsen = 25, True
        self._ctx = None
        self._host_stale = False
        self._print(f"Problem setup in {time.perf_counter() - tic:0.4f} seconds.\n")

    # ------------------------------------------------------------------ helpers
    def _print(self, *a):
        if self.verbose and self.comm.rank == 0:
            print(*a, flush=True)

    # ------------------------------------------------------------------ config
    def read_config_file(self, config_file):
        """utils/mixed_dim_problem.py:86-374 (same keys, same RuntimeErrors)."""
        yaml.add_constructor("!range", range_constructor, Loader=yaml.FullLoader)
        with open(config_file, "r") as fh:
            config = yaml.load(fh, Loader=yaml.FullLoader)
        self.config = config
        if "solver" in config:
            self.solver_config = config["solver"]
        else:
            raise RuntimeError("Provide solver configuration in input file.")
        input_dir = config.get("input_dir", "./")
        self.output_dir = config.get("output_dir", "./output/")
        if "cell_tag_file" in config and "facet_tag_file" in config:
            mesh_file = input_dir + config["cell_tag_file"]
            facet_file = input_dir + config["facet_tag_file"]
            self.input_files = {"mesh_file": mesh_file, "facet_file": facet_file}
            if "square" in mesh_file or mesh_file == facet_file:
                self.ct_name, self.ft_name = "ct", "ft"
            else:
                self.ct_name = self.ft_name = "mesh"
        elif "synthetic_mesh" in config:
            self.input_files = {"mesh_file": "<synthetic>", "facet_file": "<synthetic>"}
        else:
            raise RuntimeError("Provide cell_tag_file and facet_tag_file fields in input file.")
        self.synthetic_mesh = config.get("synthetic_mesh")
        if "dt" in config:
            self.dt = float(config["dt"])
        else:
            raise RuntimeError("Provide dt (timestep size) field in input file.")
        if "time_steps" in config:
            self.time_steps = int(config["time_steps"])
        elif "T" in config:
            self.time_steps = int(float(config["T"]) / float(config["dt"]))
        else:
            raise RuntimeError("Provide final time T or time_steps field in input file.")
        tags = {}
        if "ics_tags" in config:
            tags["intra"] = config["ics_tags"]
        else:
            raise RuntimeError("Provide ics_tags (intracellular space tags) field in input file.")
        if "ecs_tags" in config: tags["extra"] = config["ecs_tags"]
        if "boundary_tags" in config: tags["boundary"] = config["boundary_tags"]
        if "membrane_tags" in config: tags["membrane"] = config["membrane_tags"]
        if "stimulus_tags" in config:
            self.stimulus_tags = config["stimulus_tags"]
        else:
            self.stimulus_tags = tags.get("membrane", tags["intra"])
        if "glia_tags" in config:
            tags["glia"] = config["glia_tags"]
            tags["neuron"] = [t for t in tags["intra"] if t not in tags["glia"]]
        else:
            tags["neuron"] = tags["intra"]
        self.parse_tags(tags)
        if "physical_constants" in config:
            consts = config["physical_constants"]
            if "T" in consts: self.T_value = consts["T"]
            if "R" in consts: self.R_value = consts["R"]
            if "F" in consts: self.F_value = consts["F"]
            self.psi_value = self.R_value * self.T_value / self.F_value
        else:
            self.T_value = self.R_value = self.F_value = self.psi_value = 1.0
        self.C_M_value = config.get("C_M", 1.0)
        if "mesh_conversion_factor" in config:
            self.mesh_conversion_factor = float(config["mesh_conversion_factor"])
        if "fem_order" in config:
            self.fem_order = config["fem_order"]
        if "dirichlet_bcs" in config:
            self.dirichlet_bcs = config["dirichlet_bcs"]
        if "MMS_test" in config:
            raise NotImplementedError("MMS_test is verification tooling outside the per-timestep path (SURVEY.md section 2, #8)")
        if "ion_species" in config:
            raise NotImplementedError("custom ion_species tables are not supported by the B200 path yet (Na/K/Cl only)")
        self.source_terms = config.get("source_terms")
        # point probes (mixed_dim_problem.py:278-288)
        self.gamma_points = None
        if "point_evaluation" in config:
            self.point_evaluation = True
            pe = config["point_evaluation"]
            scale = float(config.get("mesh_conversion_factor", self.mesh_conversion_factor))
            self.ics_points = np.atleast_2d(np.array(pe["ics_points"], float)) * scale
            self.ecs_points = np.atleast_2d(np.array(pe["ecs_points"], float)) * scale
            if "gamma_points" in pe:
                self.gamma_points = np.atleast_2d(np.array(pe["gamma_points"], float)) * scale
        else:
            self.point_evaluation = False
        if "stimulus" in config:
            try:
                g_dict = config["stimulus"]["conductance"]
                self.g_syn_bar_val = g_dict["g_syn_bar"]
                self.a_syn_val = config["stimulus"]["a_syn"]
                self.T_stim_val = config["stimulus"]["T_stim"]
            except Exception:
                raise RuntimeError("For stimulus, provide g_syn_bar, a_syn and T_stim in input file.")
            if "tau_syn_rise" in config["stimulus"] or "tau_syn_decay" in config["stimulus"]:
                try:
                    self.tau_syn_rise = config["stimulus"]["tau_syn_rise"]
                    self.tau_syn_decay = config["stimulus"]["tau_syn_decay"]
                except Exception:
                    raise RuntimeError("For rise and decay stimulus, provide tau_syn_rise and tau_syn_decay in input file.")
            if "scale" in config["stimulus"]:
                self.scale_stimulus = config["stimulus"]["scale"]
            else:
                raise RuntimeError("Provide whether to scale stimulus strength by surface area in stimulus configuration in input file.")
            self.g_Na_bar_val = g_dict.get("g_Na_bar", 1200.0)
            self.g_K_bar_val = g_dict.get("g_K_bar", 360.0)
            self.g_Na_leak_val = g_dict.get("g_Na_leak", 0.3)
            self.g_Na_leak_g_val = g_dict.get("g_Na_leak_g", 1.0)
            self.g_K_leak_val = g_dict.get("g_K_leak", 0.1)
            self.g_K_leak_g_val = g_dict.get("g_K_leak_g", 16.96)
            self.g_Cl_leak_val = g_dict.get("g_Cl_leak", 0.25)
            self.g_Cl_leak_g_val = g_dict.get("g_Cl_leak_g", 2.0)
        else:
            self.g_syn_bar_val, self.a_syn_val, self.T_stim_val, self.scale_stimulus = 40.0, 5e-4, 1.0, False
            self.g_Na_bar_val, self.g_K_bar_val = 1200, 360
            self.g_Na_leak_val, self.g_Na_leak_g_val = 1.0, 1.0
            self.g_K_leak_val, self.g_K_leak_g_val = 4.0, 16.96
            self.g_Cl_leak_val, self.g_Cl_leak_g_val = 0.25, 0.50
        if "stimulus_region" in config:
            self.stimulus_region = True
            self.stimulus_region_range = np.array(config["stimulus_region"]["range"]) * self.mesh_conversion_factor
            axes = {"x": 0, "y": 1, "z": 2}
            if config["stimulus_region"].get("multiple", False):
                # several directions: range = [[lo, hi], ...], direction = [axis, ...] (mixed_dim_problem.py:346-351)
                self.multiple_stimulus_directions = True
                self.stimulus_region_directions = [axes[str(d)] for d in config["stimulus_region"]["direction"]]
                if len(self.stimulus_region_directions) > 3 or np.shape(self.stimulus_region_range) != (len(self.stimulus_region_directions), 2):
                    raise RuntimeError("stimulus_region with multiple: give one [lo, hi] range per direction (at most 3).")
            else:
                self.multiple_stimulus_directions = False
                self.stimulus_region_direction = axes[str(config["stimulus_region"]["direction"])]
        else:
            self.stimulus_region = False
        if "initial_conditions" in config:
            self.initial_conditions = config["initial_conditions"]
            self.find_initial_conditions = False
        else:
            self.find_initial_conditions = True
        if "membrane_data_tag" in config:
            self.membrane_data_tag = int(config["membrane_data_tag"])
        else:
            self.membrane_data_tag = self.stimulus_tags[0] if len(self.stimulus_tags) > 0 else self.gamma_tags[0]

    def parse_tags(self, tags: dict):
        """utils/mixed_dim_problem.py:376-433."""
        allowed = {"intra", "extra", "membrane", "boundary", "glia", "neuron"}
        if not set(tags).issubset(allowed):
            raise ValueError(f"Mismatch in tags.\nAllowed tags: {allowed}\nInput tags: {set(tags)}")
        if isinstance(tags["intra"], collections.abc.Sequence):
            self._print(f"# Cell tags = {len(tags['intra'])}.")
        self.intra_tags = tags["intra"]
        self.extra_tag = tags.get("extra", 1)
        self.gamma_tags = tags.get("membrane", self.intra_tags)
        if "glia" in tags:
            self.glia_tags = tags["glia"]
            self.glia_flag = len(self.glia_tags) > 0
        else:
            self.glia_tags, self.glia_flag = None, False
        self.neuron_tags = tags["neuron"]
        self.boundary_tags = tags.get("boundary", ())
        as_tuple = lambda v: tuple(int(t) for t in (v if isinstance(v, collections.abc.Sequence) else (v,)))
        self.intra_tags = as_tuple(self.intra_tags)
        self.extra_tag = as_tuple(self.extra_tag)
        self.boundary_tags = as_tuple(self.boundary_tags)
        self.gamma_tags = as_tuple(self.gamma_tags)
        self.neuron_tags = as_tuple(self.neuron_tags)
        self.stimulus_tags = as_tuple(self.stimulus_tags)
        if self.glia_flag:
            self.glia_tags = as_tuple(self.glia_tags)
        if len(self.extra_tag) != 1:
            raise NotImplementedError("exactly one extracellular tag is supported (as in every reference config)")

    # ------------------------------------------------------------------ domain
    def setup_domain(self):
        """Mesh ingest (utils/mixed_dim_problem.py:634-733): an existing XDMF file is read (xdmf.py / hdf5_min.py); otherwise
        the fixture is generated in memory from ``synthetic_mesh`` or from the file name (``square{N}.xdmf`` /
        ``cube{N}.xdmf``, as written by utils/generate_square_mesh.py)."""
        local_info = None
        sm = self.synthetic_mesh
        if sm is not None and self.comm.size > 1 and sm.get("kind") == "cell_array" and self.gamma_tags == self.intra_tags:
            # structured tissue block on several GPUs: every rank generates only its own slab (block partition by formula)
            gdim, n = int(sm.get("dim", 2)), int(sm.get("N", 32))
            m, local_info = _mesh.cell_array_mesh_local(gdim, n, int(sm.get("cells_per_dim", 8)), self.comm.rank, self.comm.size,
                                                        self.mesh_conversion_factor, float(sm.get("fill", 0.5)),
                                                        int(sm.get("first_tag", 2)), int(sm.get("extra_tag", 1)), sm.get("shape"))
            self.global_mesh_info = dict(n_vertices=(n + 1) ** gdim, n_cells=(2 if gdim == 2 else 6) * n ** gdim)
        elif sm is not None:
            m = _mesh.from_descriptor(self.synthetic_mesh, self.mesh_conversion_factor)
        else:
            base = os.path.basename(self.input_files["mesh_file"])
            mt = re.match(r"(square|cube)(\d+)\.xdmf$", base)
            if os.path.exists(self.input_files["mesh_file"]):
                # a mesh file that exists is read (never replaced by a same-named synthetic fixture)
                self._print("Reading mesh from XDMF file...")
                m = _mesh.from_xdmf(self.input_files["mesh_file"], self.input_files["facet_file"], self.ct_name, self.ft_name,
                                    self.intra_tags, self.extra_tag[0], self.boundary_tags, self.mesh_conversion_factor)
                mt = None
            elif not mt:
                raise RuntimeError(f"Cannot read {self.input_files['mesh_file']}: the file does not exist (generated fixtures are "
                                   "available under the names square{N}.xdmf / cube{N}.xdmf or through a synthetic_mesh block).")
        if sm is None and mt:
            n = int(mt.group(2))
            m = (_mesh.unit_square_fixture if mt.group(1) == "square" else _mesh.unit_cube_fixture)(
                n, self.mesh_conversion_factor)
        if self.fem_order == 2:
            # ("Lagrange", 2) spaces (KNPEMIx_problem.py:38-42): the host code below works on the nodes of the P2 space
            if self.comm.size > 1:
                raise NotImplementedError("fem_order = 2 runs on one GPU")
            m = _mesh.p2_node_mesh(m)
        if not (all(t < self.extra_tag[0] for t in self.intra_tags) or all(t > self.extra_tag[0] for t in self.intra_tags)):
            raise RuntimeError("Intracellular tags must be all smaller or all larger than extracellular tag.")
        # keep only membrane facets whose tag is listed, check the listed tags exist
        keep = np.isin(m.mf_tags, np.asarray(self.gamma_tags))
        m.mf_verts, m.mf_tags = m.mf_verts[keep], m.mf_tags[keep]
        m.intra_tags, m.extra_tag = self.intra_tags, self.extra_tag[0]
        if local_info is not None:
            if m.mf_owned is not None:
                m.mf_owned = m.mf_owned[keep]
            self.halo = local_info
        elif self.comm.size > 1:
            self.global_mesh_info = dict(n_vertices=m.x.shape[0], n_cells=m.cells.shape[0])
            from .partition import partition_mesh
            m, self.halo = partition_mesh(m, self.comm.rank, self.comm.size)
        else:
            self.global_mesh_info = dict(n_vertices=m.x.shape[0], n_cells=m.cells.shape[0])
            self.halo = None
        self.mesh = m
        self.dx = Measure(self)
        self.dS = Measure(self)

    # ------------------------------------------------------------------ constants / spaces
    def setup_constants(self):
        """KNPEMIx_problem.py:909-981."""
        C = Constant
        self.C_M, self.T, self.F, self.R, self.psi = C(self.C_M_value), C(self.T_value), C(self.F_value), C(self.R_value), C(self.psi_value)
        self.g_Na_bar, self.g_K_bar = C(self.g_Na_bar_val), C(self.g_K_bar_val)
        self.g_Na_leak, self.g_Na_leak_g = C(self.g_Na_leak_val), C(self.g_Na_leak_g_val)
        self.g_K_leak, self.g_K_leak_g = C(self.g_K_leak_val), C(self.g_K_leak_g_val)
        self.g_Cl_leak, self.g_Cl_leak_g = C(self.g_Cl_leak_val), C(self.g_Cl_leak_g_val)
        self.g_syn_bar, self.a_syn, self.T_stim = C(self.g_syn_bar_val), C(self.a_syn_val), C(self.T_stim_val)
        self.D_Na, self.D_K, self.D_Cl = C(1.33e-9), C(1.96e-9), C(2.03e-9)
        self.phi_rest = C(-0.065)
        self.phi_m_init = C(-0.070)
        self.Na_i_init, self.Na_e_init = C(10.0), C(145.0)
        self.K_i_init, self.K_e_init = C(130.0), C(3.0)
        self.Cl_i_init, self.Cl_e_init = C(5.0), C(134.0)
        self.phi_m_n_init, self.phi_m_g_init = C(self.phi_m_init.value), C(-0.085)
        self.Na_i_n_init, self.K_i_n_init, self.Cl_i_n_init = C(self.Na_i_init.value), C(self.K_i_init.value), C(self.Cl_i_init.value)
        self.Na_i_g_init, self.K_i_g_init, self.Cl_i_g_init = C(15.0), C(100.0), C(5.0)
        self.n_init, self.m_init, self.h_init = C(0.24458654944007155), C(0.028905534475191896), C(0.7540796658225248)
        mk = lambda name, gl, glg, D, ki, ke, kin, kig, z: {"name": name, "g_leak": gl, "g_leak_g": glg, "Di": D, "De": D,
                                                              "ki_init": ki, "ke_init": ke, "ki_init_n": kin, "ki_init_g": kig,
                                                              "z": C(z), "f_e": C(0.0), "f_i": C(0.0)}
        self.Na = mk("Na", self.g_Na_leak, self.g_Na_leak_g, self.D_Na, self.Na_i_init, self.Na_e_init, self.Na_i_n_init, self.Na_i_g_init, 1.0)
        self.K = mk("K", self.g_K_leak, self.g_K_leak_g, self.D_K, self.K_i_init, self.K_e_init, self.K_i_n_init, self.K_i_g_init, 1.0)
        self.Cl = mk("Cl", self.g_Cl_leak, self.g_Cl_leak_g, self.D_Cl, self.Cl_i_init, self.Cl_e_init, self.Cl_i_n_init, self.Cl_i_g_init, -1.0)
        self.ion_list = [self.Na, self.K, self.Cl]
        self.N_ions = len(self.ion_list)

    def setup_spaces(self):
        """KNPEMIx_problem.py:28-94: ("Lagrange", fem_order) space on the whole mesh, 8 fields, restrictions to the dofs of the
        intracellular / extracellular cells (membrane dofs belong to both).  fem_order = 2: the mesh is the P2 node mesh
        (mesh.p2_node_mesh; dof = node = vertex or edge midpoint), everything below is written per node."""
        if self.fem_order not in (1, 2):
            raise NotImplementedError(f"fem_order = {self.fem_order}: P1 and P2 elements are implemented")
        self._print("Setting up function spaces ...")
        m = self.mesh
        self.num_variables = self.N_ions + 1
        self.num_variables_total = 2 * self.num_variables
        nv = m.x.shape[0]
        names_i = [f"{ion['name']}_i" for ion in self.ion_list] + ["phi_i"]
        names_e = [f"{ion['name']}_e" for ion in self.ion_list] + ["phi_e"]
        self.wh = [[Function(self, nv, nm) for nm in names_i], [Function(self, nv, nm) for nm in names_e]]
        self.u_out_i, self.u_out_e = list(self.wh[0]), list(self.wh[1])
        self._print("Creating mesh restrictions ...")
        is_in = np.isin(m.cell_tags, np.asarray(self.intra_tags))
        is_ex = m.cell_tags == self.extra_tag[0]
        self.dofs_intra = np.unique(m.cells[is_in].ravel()).astype(np.int32)
        self.dofs_extra = np.unique(m.cells[is_ex].ravel()).astype(np.int32)
        n_owned = nv if m.n_owned is None else m.n_owned
        self.interior = _Restriction(self.dofs_intra, int((self.dofs_intra < n_owned).sum()))
        self.exterior = _Restriction(self.dofs_extra, int((self.dofs_extra < n_owned).sum()))
        self.restriction = [self.interior] * self.num_variables + [self.exterior] * self.num_variables
        self.neuron_cells = np.flatnonzero(np.isin(m.cell_tags, np.asarray(self.neuron_tags)))
        if self.glia_flag:
            self.glia_cells = np.flatnonzero(np.isin(m.cell_tags, np.asarray(self.glia_tags)))

    def init(self):
        pass

    def setup_boundary_conditions(self):
        """KNPEMIx_problem.py:96-198.  ``dirichlet_bcs``: every field keeps its initial value (k_init; phi_m_init inside, 0
        outside) on the vertices of the facets tagged ``boundary_tags`` (:139-161).  ``pin_ecs_potential`` (class switch,
        :997): phi_e = 0 at one extracellular vertex off the membrane (:163-194; the reference takes geometry point 0 or a
        random one, here the lowest-numbered admissible vertex).  The constrained dofs go to the device once the dof maps
        exist (_upload_bcs -> knp_set_dirichlet); the device applies them in every assembly.  Shipped configs use neither
        (bcs = [])."""
        self._print("Setting up boundary conditions ...")
        self.bcs = []
        self._bc_verts = None
        if self.dirichlet_bcs:
            m = self.mesh
            if m.bc_verts is not None:                       # ingested mesh: facets tagged boundary_tags
                self._bc_verts = np.asarray(m.bc_verts, np.int32)
            else:                                            # generated fixtures: the exterior boundary carries ONE tag
                sm = self.synthetic_mesh or {}           # square / cube: PARTIAL_OMEGA = 3 (utils/misc.py:139); tissue blocks: 1
                btag = int(sm.get("boundary_tag", 3 if sm.get("kind", "square") in ("square", "cube") else 1))
                self._bc_verts = _mesh.boundary_vertices(m) if btag in self.boundary_tags else np.zeros(0, np.int32)
            self.bcs = ["dirichlet: %d boundary vertices" % self._bc_verts.size]
        elif self.pin_ecs_potential:
            self.bcs = ["phi_e pinned at one vertex"]

    def _bc_entries(self, node_vert, mverts):
        """Constrained dofs in the column layout (owned and ghost columns of this rank) with their values; node_vert = the
        restricted dof maps (knp_dofmap_host), mverts = the membrane dofs (knp_mverts_host)."""
        from .partition import Layout
        m = self.mesh
        n_owned = m.x.shape[0] if m.n_owned is None else m.n_owned
        lay = Layout(node_vert, n_owned)
        inv = []
        for s in range(2):
            a = np.full(m.x.shape[0], -1, np.int64)
            a[node_vert[s]] = np.arange(node_vert[s].size)
            inv.append(a)
        cols, vals = [], []
        if self.dirichlet_bcs:
            for s, suffix in ((0, "i"), (1, "e")):
                q = inv[s][self._bc_verts]
                q = q[q >= 0]
                for k, ion in enumerate(self.ion_list):
                    cols.append(lay.col(s, k, q))
                    vals.append(np.full(q.size, ion[f"k{suffix}_init"].value))
                cols.append(lay.col(s, self.N_ions, q))
                vals.append(np.full(q.size, self.phi_m_init.value if s == 0 else 0.0))
        else:
            gid = np.arange(m.x.shape[0], dtype=np.int64) if m.vert_global is None else np.asarray(m.vert_global, np.int64)
            ok = inv[1] >= 0
            ok[mverts] = False
            ok[n_owned:] = False
            mine = int(gid[ok].min()) if ok.any() else np.iinfo(np.int64).max
            pin = int(self.comm.allreduce(float(mine), op=MPI.MIN))
            self.pinned_vertex = pin
            loc = np.flatnonzero(gid == pin)                  # the owner constrains the row, neighbours the ghost column
            q = inv[1][loc]
            q = q[q >= 0]
            cols.append(lay.col(1, self.N_ions, q))
            vals.append(np.zeros(q.size))
            self._print("Phi_e pinned at (vertex, point):", pin, m.x[loc[0]] if loc.size else "")
        return np.concatenate(cols), np.concatenate(vals)

    def _upload_bcs(self):
        if not self.bcs:
            return
        self._ctx.set_dirichlet(*self._bc_entries(self._node_vert, self._mverts))

    injection_current = 5e-9      # [A], KNPEMIx_problem.py:211

    def _cell_volumes(self, cells):
        d = self.mesh.gdim
        x = self.mesh.x[cells[:, :d + 1]]                  # the vertices come first (P2 node mesh)
        det = np.linalg.det(x[:, 1:] - x[:, :1])
        return np.abs(det) / (2.0 if d == 2 else 6.0)

    def setup_source_terms(self):
        """Ion injection (KNPEMIx_problem.py:200-218; site: utils/mixed_dim_problem.py:496-540,806-811): K+ and Cl- enter the
        extracellular space at I / F mol/s, spread over the cells whose vertices all lie within (x_max - x_min)/10 of the
        centre of the mesh's bounding box.  f_e is a P1 function (value on every vertex of an injection cell); the form
        dt (f_e, v) dx_e only sees extracellular cells.  The device adds the resulting constant entries to b every step
        (knp_set_source); they are formed in _upload_source() once the dof maps exist."""
        m = self.mesh
        lo = np.array([self.comm.allreduce(float(v), op=MPI.MIN) for v in m.x.min(axis=0)])
        hi = np.array([self.comm.allreduce(float(v), op=MPI.MAX) for v in m.x.max(axis=0)])
        centre, delta, tol = 0.5 * (lo + hi), (hi[0] - lo[0]) / 10.0, 1e-14
        self.x_L, self.y_L = centre[0] - delta, centre[1] - delta
        self.x_U, self.y_U = centre[0] + delta, centre[1] + delta
        inside = np.all((m.x >= centre - delta - tol) & (m.x <= centre + delta + tol), axis=1)
        self.injection_cells = np.flatnonzero(inside[m.cells].all(axis=1))
        vol = self._cell_volumes(m.cells[self.injection_cells])
        if m.cell_owned is not None:
            vol = vol * (np.asarray(m.cell_owned)[self.injection_cells] != 0)
        self.injection_volume = self.comm.allreduce(float(vol.sum()), op=MPI.SUM)
        if not self.injection_volume > 0.0:
            raise RuntimeError("ion_injection: no mesh cell lies inside the injection site")
        src_term = (self.injection_current / (1 * self.F.value)) / self.injection_volume      # [mol / (m^3 s)] = [mM / s]
        nv = m.x.shape[0]
        verts = np.unique(m.cells[self.injection_cells].ravel())
        for ion in (self.K, self.Cl):
            f = Function(self, nv, f"f_e_{ion['name']}")
            f._data[verts] = src_term
            ion["f_e"] = f

    def _source_entries(self, node_vert):
        """(rows, values) of the entries dt * sum_{c in ECS} (M_c f_e)|_p of the right-hand side for the owned extracellular
        dofs (the local mesh holds every cell that touches an owned vertex); node_vert = the restricted dof maps
        (knp_dofmap_host).  P1 and P2: M_c = |c| x the reference mass matrix of the element."""
        from .partition import Layout
        m = self.mesh
        d = m.gdim
        n_owned = m.x.shape[0] if m.n_owned is None else m.n_owned
        lay = Layout(node_vert, n_owned)
        node_of = np.full(m.x.shape[0], -1, np.int64)
        node_of[node_vert[1]] = np.arange(node_vert[1].size)
        cells = m.cells[m.cell_tags == self.extra_tag[0]]
        vol = self._cell_volumes(cells)
        Mref = _mesh.reference_mass(d, m.degree)                             # int N_a N_b / |cell|
        rows, vals = [], []
        for k, ion in enumerate(self.ion_list):
            f = ion["f_e"]
            if not isinstance(f, Function) or not f._data.any():
                continue
            fc = f._data[cells]                                              # (nc, dofs per cell)
            contrib = vol[:, None] * (fc @ Mref)                            # M_c f  (P1: vol/((d+1)(d+2)) (f_p + sum_q f_q))
            sv = np.bincount(cells.ravel(), weights=contrib.ravel(), minlength=m.x.shape[0]) * float(self.dt.value)
            nodes = node_of[np.flatnonzero(sv)]
            nodes = nodes[(nodes >= 0) & (nodes < lay.n_own[1])]
            rows.append(lay.col(1, k, nodes))
            vals.append(sv[node_vert[1][nodes]])
        if not rows:
            return np.zeros(0, np.int64), np.zeros(0)
        return np.concatenate(rows), np.concatenate(vals)

    def _upload_source(self):
        rows, vals = self._source_entries(self._node_vert)
        if rows.size:
            self._ctx.set_source(rows, vals)

    # ------------------------------------------------------------------ initial conditions
    def set_initial_conditions(self):
        """KNPEMIx_problem.py:220-452 (config-provided initial conditions, :326-353 and :386-447)."""
        if self.find_initial_conditions:
            self._find_steady_state_initial_conditions()
            nv = self.mesh.x.shape[0]
            self.phi_m_prev = Function(self, nv, "phi_m")
            self._fill_initial_fields()
            self._print("Initial conditions set.")
            return
        self._print("Setting initial conditions from input file ...")
        ic = self.initial_conditions
        pick = lambda a, b: ic[a] if a in ic else ic[b]
        if not self.glia_flag:
            self.phi_m_init.value = pick("phi_m", "phi_m_n")
            self.Na_i_init.value = pick("Na_i", "Na_i_n")
            self.K_i_init.value = pick("K_i", "K_i_n")
            self.Cl_i_init.value = pick("Cl_i", "Cl_i_n")
        else:
            self.phi_m_n_init.value, self.phi_m_g_init.value = ic["phi_m_n"], ic["phi_m_g"]
            self.Na_i_n_init.value, self.Na_i_g_init.value = ic["Na_i_n"], ic["Na_i_g"]
            self.K_i_n_init.value, self.K_i_g_init.value = ic["K_i_n"], ic["K_i_g"]
            self.Cl_i_n_init.value, self.Cl_i_g_init.value = ic["Cl_i_n"], ic["Cl_i_g"]
        self.Na_e_init.value, self.K_e_init.value, self.Cl_e_init.value = ic["Na_e"], ic["K_e"], ic["Cl_e"]
        self.n_init.value, self.m_init.value, self.h_init.value = ic["n"], ic["m"], ic["h"]
        nv = self.mesh.x.shape[0]
        self.phi_m_prev = Function(self, nv, "phi_m")
        self._fill_initial_fields()
        self._print("Initial conditions set.")

    def calculate_compartment_volumes_and_surface_areas(self):
        """utils/mixed_dim_problem.py:813-849: volumes [m^3] of the neuronal / glial intracellular space and of the ECS,
        membrane areas [m^2] of neurons / glia (host sums over the owned cells / facets, all-reduced)."""
        m = self.mesh
        vol = self._cell_volumes(m.cells)
        if m.cell_owned is not None:
            vol = vol * (np.asarray(m.cell_owned) != 0)
        xf = m.x[m.mf_verts]
        if m.gdim == 2:
            area = np.linalg.norm(xf[:, 1] - xf[:, 0], axis=1)
        else:
            area = 0.5 * np.linalg.norm(np.cross(xf[:, 1] - xf[:, 0], xf[:, 2] - xf[:, 0]), axis=1)
        if m.mf_owned is not None:
            area = area * (np.asarray(m.mf_owned) != 0)
        tot = lambda v: self.comm.allreduce(float(v), op=MPI.SUM)
        self.vol_i_n = tot(vol[np.isin(m.cell_tags, np.asarray(self.neuron_tags))].sum())
        self.area_g_n = tot(area[np.isin(m.mf_tags, np.asarray(self.neuron_tags))].sum())
        self.vol_e = tot(vol[m.cell_tags == self.extra_tag[0]].sum())
        if self.glia_flag:
            self.vol_i_g = tot(vol[np.isin(m.cell_tags, np.asarray(self.glia_tags))].sum())
            self.area_g_g = tot(area[np.isin(m.mf_tags, np.asarray(self.glia_tags))].sum())

    def _find_steady_state_initial_conditions(self):
        """KNPEMIx_problem.py:224-325: no initial_conditions in the config -> integrate the well-mixed membrane ODE system to
        rest on rank 0 (steady_state.py, mirror of utils/membrane_ODE_systems.py), broadcast, overwrite the constants."""
        from .steady_state import MembraneSteadyState
        self._print("Solving ODE system to find steady-state initial conditions ...")
        self.calculate_compartment_volumes_and_surface_areas()
        v = lambda c: float(c.value)
        consts = dict(R=v(self.R), F=v(self.F), T=v(self.T), C_M=v(self.C_M), g_Na_bar=v(self.g_Na_bar), g_K_bar=v(self.g_K_bar),
                      g_leak=(v(self.g_Na_leak), v(self.g_K_leak), v(self.g_Cl_leak)),
                      g_leak_g=(v(self.g_Na_leak_g), v(self.g_K_leak_g), v(self.g_Cl_leak_g)), phi_rest=v(self.phi_rest),
                      phi_m=v(self.phi_m_init), c_i=(v(self.Na_i_init), v(self.K_i_init), v(self.Cl_i_init)),
                      c_e=(v(self.Na_e_init), v(self.K_e_init), v(self.Cl_e_init)), phi_m_g=v(self.phi_m_g_init),
                      c_i_g=(v(self.Na_i_g_init), v(self.K_i_g_init), v(self.Cl_i_g_init)))
        geom = dict(vol_i_n=self.vol_i_n, vol_e=self.vol_e, area_n=self.area_g_n)
        if self.glia_flag:
            geom.update(vol_i_g=self.vol_i_g, area_g=self.area_g_g)
        x = None
        if self.comm.rank == 0:
            x, t_end, reached = MembraneSteadyState(consts, geom, glia=self.glia_flag).solve()
            self._print("Steady state reached. Derivatives zero to within tolerance." if reached
                        else "Max time exceeded without finding steady state. Exiting.")
            x = [float(val) for val in x]
        x = self.comm.bcast(x, root=0)
        self.steady_state = x
        if not self.glia_flag:
            (self.phi_m_init.value, self.Na_i_init.value, self.Na_e_init.value, self.K_i_init.value, self.K_e_init.value,
             self.Cl_i_init.value, self.Cl_e_init.value, self.n_init.value, self.m_init.value, self.h_init.value) = x
        else:
            (self.phi_m_n_init.value, self.Na_i_n_init.value, self.Na_e_init.value, self.K_i_n_init.value, self.K_e_init.value,
             self.Cl_i_n_init.value, self.Cl_e_init.value, self.phi_m_g_init.value, self.Na_i_g_init.value,
             self.K_i_g_init.value, self.Cl_i_g_init.value, self.n_init.value, self.m_init.value, self.h_init.value) = x

    def _fill_initial_fields(self):
        """Write the initial conditions into wh / phi_m_prev (also used by the iterative solver's
        initial-guess reset, KNPEMIx_solver.py:179-199)."""
        ui, ue = self.wh[0], self.wh[1]
        N = self.N_ions
        if not self.glia_flag:
            self.phi_m_prev._data[:] = self.phi_m_init.value
            ui[N]._data[:] = self.phi_m_init.value
            ue[N]._data[:] = 0.0
            for idx, ion in enumerate(self.ion_list):
                ui[idx]._data[:] = ion["ki_init"].value
                ue[idx]._data[:] = ion["ke_init"].value
        else:
            m = self.mesh
            self.neuron_dofs = np.unique(m.cells[self.neuron_cells].ravel())
            self.glia_dofs = np.unique(m.cells[self.glia_cells].ravel())
            self.phi_m_prev._data[self.neuron_dofs] = self.phi_m_n_init.value
            self.phi_m_prev._data[self.glia_dofs] = self.phi_m_g_init.value
            ui[N]._data[self.neuron_dofs] = self.phi_m_n_init.value
            ui[N]._data[self.glia_dofs] = self.phi_m_g_init.value
            ue[N]._data[:] = 0.0
            for idx, ion in enumerate(self.ion_list):
                ui[idx]._data[self.neuron_dofs] = ion["ki_init_n"].value
                ui[idx]._data[self.glia_dofs] = ion["ki_init_g"].value
                ue[idx]._data[:] = ion["ke_init"].value
        self._apply_initial_perturbation()
        self._host_stale = False
        if self._ctx is not None:
            self._push_state()

    def _apply_initial_perturbation(self):
        """Extension key ``initial_perturbation`` (synthetic benchmark configs only):
        concentrations x (1 + r sin 2 pi X sin 2 pi Y), phi_m = phi_m + a cos 2 pi X with X, Y in unit-mesh
        coordinates, so that the assembled coefficients are not constant."""
        pert = self.config.get("initial_perturbation")
        if not pert:
            return
        X = self.mesh.x / self.mesh_conversion_factor
        r, a = float(pert.get("relative", 0.0)), float(pert.get("phi_m_amplitude", 0.0))
        fac = 1.0 + r * np.sin(2 * np.pi * X[:, 0]) * np.sin(2 * np.pi * X[:, 1])
        N = self.N_ions
        for s in range(2):
            for k in range(N):
                self.wh[s][k]._data *= fac
        dphi = a * np.cos(2 * np.pi * X[:, 0])
        self.phi_m_prev._data += dphi
        self.wh[0][N]._data += dphi

    def init_ionic_models(self, ionic_models):
        """utils/mixed_dim_problem.py:435-465."""
        if not isinstance(ionic_models, (list, tuple)):
            ionic_models = [ionic_models]
        self.ionic_models = list(ionic_models)
        self.gating_variables = False
        ionic_tags = set()
        from .ionic_models import HodgkinHuxley
        nv = self.mesh.x.shape[0]
        for model in self.ionic_models:
            model._init()
            ionic_tags.update(model.tags)
            self._print("Added tags for ionic model: ", str(model))
            if isinstance(model, HodgkinHuxley):
                self.gating_variables = True
                self.n, self.m, self.h = Function(self, nv, "n"), Function(self, nv, "m"), Function(self, nv, "h")
                self.n._data[:], self.m._data[:], self.h._data[:] = self.n_init.value, self.m_init.value, self.h_init.value
                self._print("Gating variables flag set to True.")
        ionic_tags = sorted(ionic_tags)
        gamma_tags = sorted(flatten_list([self.gamma_tags]))
        if ionic_tags != gamma_tags and not self.MMS_test and len(ionic_tags) != 0:
            raise RuntimeError("Mismatch between membrane tags and ionic models tags."
                               + f"\nIonic models tags: {ionic_tags}\nMembrane tags: {gamma_tags}")
        self._print("# Membrane tags = ", len(gamma_tags))
        self._print("# Ionic models  = ", len(self.ionic_models), "\n")

    # ------------------------------------------------------------------ device context
    def _require_context(self):
        if self._ctx is None:
            raise RuntimeError("call setup_variational_form() first (it creates the CUDA context)")
        return self._ctx

    def setup_variational_form(self):
        """KNPEMIx_problem.py:454-655.  The reference builds the UFL forms a, L and JIT-compiles them; here
        the same forms are hard-wired in the CUDA kernels, so this step creates the device context (dof maps,
        CSR pattern, gather maps), uploads constants / model table and the initial state."""
        self._print("Setting up variational form ...")
        m = self.mesh
        qb, qw = _mesh.facet_quadrature(m.gdim)
        if self.device is None:
            self.device = int(os.environ.get("LOCAL_RANK", "0"))
        self._ctx = _lib.Context(m.gdim, m.x, m.cells, m.cell_tags, self.intra_tags, self.extra_tag[0], m.mf_verts,
                                 m.mf_tags, qb, qw, n_owned_vertices=m.n_owned, cell_owned=m.cell_owned,
                                 mfacet_owned=m.mf_owned, device=self.device, degree=m.degree)
        ctx = self._ctx
        self._node_vert = ctx.dofmaps()
        self._mverts = ctx.mverts()
        if self.comm.size > 1:
            from .partition import init_halo
            init_halo(self, ctx)
        self._upload_params()
        self._push_state()
        if self.source_terms == "ion_injection":
            self._upload_source()
        self._upload_bcs()
        if self.point_evaluation:
            self._setup_probes()
        self.a = self.L = "device-resident forms (csrc/assembly.cu)"

    # ------------------------------------------------------------------ point probes
    def _locate(self, point, cells):
        """Index into `cells` of the first cell containing `point` and its barycentric coordinates, or (None, None)."""
        m = self.mesh
        d = m.gdim
        xc = m.x[cells[:, :d + 1]]                                         # (nc, d+1, d): the vertices come first
        tol = 1e-9 * float(np.abs(m.x).max())
        cand = np.flatnonzero(np.all(xc.min(1) <= point + tol, axis=1) & np.all(xc.max(1) >= point - tol, axis=1))
        if cand.size == 0:
            return None, None
        T = np.transpose(xc[cand, 1:] - xc[cand, :1], (0, 2, 1))           # columns = edge vectors
        lam = np.linalg.solve(T, (point - xc[cand, 0])[:, :, None])[:, :, 0]
        bary = np.concatenate([1.0 - lam.sum(1, keepdims=True), lam], 1)
        ok = np.flatnonzero(np.all(bary >= -1e-9, axis=1))
        if ok.size == 0:
            return None, None
        return int(cand[ok[0]]), bary[ok[0]]

    def _setup_probes(self):
        ptr, cols, wts = self._probe_tables(self._node_vert)
        self._ctx.probe_setup(ptr, cols if cols else [0], wts if wts else [0.0])

    def _probe_tables(self, node_vert):
        """Containing cells and shape-function weights of the probe points (scifem.evaluate_function, KNPEMIx_solver.py:612-643),
        found once on the host; the device evaluates the resulting sparse functionals of the state every step.  Output order:
        [ics point p, variable j] , [ecs point p, variable j] , [gamma point p] (phi_m = phi_i - phi_e on the membrane).
        Returns (ptr, columns in the column layout, weights); node_vert = the restricted dof maps (knp_dofmap_host)."""
        from .partition import Layout
        m = self.mesh
        d = m.gdim
        n_owned = m.x.shape[0] if m.n_owned is None else m.n_owned
        lay = Layout(node_vert, n_owned)
        inv = []
        for s in range(2):
            a = np.full(m.x.shape[0], -1, np.int64)
            a[node_vert[s]] = np.arange(node_vert[s].size)
            inv.append(a)
        owned = np.ones(m.cells.shape[0], bool) if m.cell_owned is None else m.cell_owned.astype(bool)
        is_in = np.isin(m.cell_tags, np.asarray(self.intra_tags)) & owned
        is_ex = (m.cell_tags == self.extra_tag[0]) & owned
        ptr, cols, wts = [0], [], []
        found = []

        def claim(hit):
            # exactly one rank evaluates a point (the lowest rank that holds a containing owned cell)
            mine = float(self.comm.rank) if hit else float(self.comm.size)
            return self.comm.allreduce(mine, op=MPI.MIN) == float(self.comm.rank) and hit

        for s, pts, sel in ((0, self.ics_points, is_in), (1, self.ecs_points, is_ex)):
            idx = np.flatnonzero(sel)
            for pt in pts:
                c, bary = self._locate(pt[:d], m.cells[idx])
                take = claim(c is not None)
                found.append(self.comm.allreduce(float(c is not None), op=MPI.MAX) > 0)
                for f in range(4):
                    if take:
                        nodes = inv[s][m.cells[idx[c]]]
                        cols += [int(v) for v in lay.col(s, f, nodes)]
                        wts += [float(b) for b in (bary if m.degree == 1 else _mesh.p2_shape(bary))]
                    ptr.append(len(cols))
        if self.gamma_points is not None:
            fowned = np.ones(m.mf_verts.shape[0], bool) if m.mf_owned is None else m.mf_owned.astype(bool)
            fidx = np.flatnonzero(fowned)
            for pt in self.gamma_points:
                best, bw = None, None
                if fidx.size:
                    xf = m.x[m.mf_verts[fidx][:, :d]]                      # (nf, d, d): closest point by facet barycentrics
                    E = np.transpose(xf[:, 1:] - xf[:, :1], (0, 2, 1))     # (nf, d, d-1)
                    rhs = (pt[:d] - xf[:, 0])[:, :, None]
                    G = np.transpose(E, (0, 2, 1)) @ E
                    lam = np.linalg.solve(G, np.transpose(E, (0, 2, 1)) @ rhs)[:, :, 0]
                    bary = np.concatenate([1.0 - lam.sum(1, keepdims=True), lam], 1)
                    inside = np.all(bary >= -1e-9, axis=1)
                    proj = np.einsum("fa,fai->fi", bary, xf)
                    dist = np.linalg.norm(proj - pt[:d], axis=1)
                    dist[~inside] = np.inf
                    k = int(np.argmin(dist))
                    if np.isfinite(dist[k]) and dist[k] <= 1e-6 * float(np.abs(m.x).max()):
                        best, bw = k, bary[k]
                take = claim(best is not None)
                found.append(self.comm.allreduce(float(best is not None), op=MPI.MAX) > 0)
                if take:
                    verts = m.mf_verts[fidx[best]]
                    if m.degree == 2:
                        bw = _mesh.p2_shape(bw)                             # trace basis on the facet's 3 / 6 nodes
                    cols += [int(v) for v in lay.col(0, 3, inv[0][verts])] + [int(v) for v in lay.col(1, 3, inv[1][verts])]
                    wts += [float(b) for b in bw] + [-float(b) for b in bw]
                ptr.append(len(cols))
        if not all(found):
            raise RuntimeError("point_evaluation: a probe point lies outside its subdomain (ics_points must be inside "
                               "intracellular cells, ecs_points inside extracellular cells, gamma_points on the membrane)")
        return ptr, cols, wts

    def evaluate_probes(self):
        """(ics values [variable, point], ecs values [variable, point], gamma values [point]) from the device state."""
        v = np.asarray(self.comm.allreduce(self._ctx.probe_eval(), op=MPI.SUM))
        ni, ne = len(self.ics_points), len(self.ecs_points)
        ng = 0 if self.gamma_points is None else len(self.gamma_points)
        ics = v[:4 * ni].reshape(ni, 4).T
        ecs = v[4 * ni:4 * (ni + ne)].reshape(ne, 4).T
        return ics, ecs, v[4 * (ni + ne):4 * (ni + ne) + ng]

    def _tag_table(self):
        from .ionic_models import HodgkinHuxley
        table = {int(t): 0 for t in self.gamma_tags}
        for model in self.ionic_models:
            for t in model.tags:
                table[int(t)] = table.get(int(t), 0) | model.flag
        return [(t, fl, (t in self.stimulus_tags)) for t, fl in sorted(table.items())]

    def _upload_params(self, stim_area=0.0):
        P = _lib.Params()
        P.dt, P.F, P.R, P.T, P.C_M = self.dt.value, self.F.value, self.R.value, self.T.value, self.C_M.value
        P.phi_rest = self.phi_rest.value
        for k, ion in enumerate(self.ion_list):
            P.z[k], P.D[k] = ion["z"].value, ion["Di"].value
            P.g_leak[k], P.g_leak_g[k] = ion["g_leak"].value, ion["g_leak_g"].value
        P.g_Na_bar, P.g_K_bar = self.g_Na_bar.value, self.g_K_bar.value
        P.g_syn_bar, P.a_syn, P.T_stim = self.g_syn_bar.value, self.a_syn.value, self.T_stim.value
        P.scale_stimulus = int(bool(self.scale_stimulus))
        for i in range(3):
            P.stim_dir[i], P.stim_lo[i], P.stim_hi[i] = -1, 0.0, 0.0
        if self.stimulus_region and self.multiple_stimulus_directions:
            for i, d in enumerate(self.stimulus_region_directions):
                P.stim_dir[i] = int(d)
                P.stim_lo[i], P.stim_hi[i] = float(self.stimulus_region_range[i][0]), float(self.stimulus_region_range[i][1])
        elif self.stimulus_region:
            P.stim_dir[0] = int(self.stimulus_region_direction)
            P.stim_lo[0], P.stim_hi[0] = float(self.stimulus_region_range[0]), float(self.stimulus_region_range[1])
        P.K_e_init, P.K_i_g_init = self.K_e_init.value, self.K_i_g_init.value
        P.ode_substeps, P.rush_larsen = int(self.ode_substeps), int(self.rush_larsen)
        P.stim_area = stim_area
        table = self._tag_table()
        self._ctx.set_params(P, table)
        if self.scale_stimulus and stim_area == 0.0:
            # p.stimulus_area = allreduce(assemble_scalar(mask * dS(stimulus_tags)))  (KNPEMIx_ionic_model.py:591-601)
            area = self.comm.allreduce(self._ctx.stimulus_area_local(), op=MPI.SUM)
            self.stimulus_area = area
            if any(st for _, _, st in table) and self.gating_variables:
                self._print(f"Stimulus area on tag {self.stimulus_tags[0]}: {area:0.6e} m^2")
            if area > 0.0:
                P.stim_area = area
                self._ctx.set_params(P, table)

    def setup_preconditioner(self, use_block_jacobi: bool = True):
        """KNPEMIx_problem.py:657-744: block-diagonal preconditioner matrix P from the current fields."""
        self._print("Setting up preconditioner ...")
        if not use_block_jacobi:
            raise NotImplementedError("only the block-Jacobi preconditioner form (the reference default) is implemented")
        ctx = self._require_context()
        if not self._host_stale:
            self._push_state()
        ctx.assemble_P()
        self.P = "device-resident preconditioner matrix (block-Jacobi form)"

    # ------------------------------------------------------------------ host <-> device state
    def _pack_u(self):
        ctx = self._ctx
        u = np.zeros(ctx.n_cols)
        off_row = [0, 4 * ctx.n_own[0]]
        gh = [ctx.n_loc[0] - ctx.n_own[0], ctx.n_loc[1] - ctx.n_own[1]]
        off_gh = [ctx.n_rows, ctx.n_rows + 4 * gh[0]]
        for s in range(2):
            verts = self._node_vert[s]
            no = ctx.n_own[s]
            for f in range(4):
                vals = self.wh[s][f]._data[verts]
                u[off_row[s] + f * no: off_row[s] + (f + 1) * no] = vals[:no]
                u[off_gh[s] + f * gh[s]: off_gh[s] + (f + 1) * gh[s]] = vals[no:]
        return u

    def _unpack_u(self, u):
        ctx = self._ctx
        off_row = [0, 4 * ctx.n_own[0]]
        gh = [ctx.n_loc[0] - ctx.n_own[0], ctx.n_loc[1] - ctx.n_own[1]]
        off_gh = [ctx.n_rows, ctx.n_rows + 4 * gh[0]]
        for s in range(2):
            verts = self._node_vert[s]
            no = ctx.n_own[s]
            for f in range(4):
                d = self.wh[s][f]._data
                d[verts[:no]] = u[off_row[s] + f * no: off_row[s] + (f + 1) * no]
                d[verts[no:]] = u[off_gh[s] + f * gh[s]: off_gh[s] + (f + 1) * gh[s]]

    def _push_state(self):
        ctx = self._ctx
        gates = None
        if self.gating_variables:
            gates = np.stack([self.n._data[self._mverts], self.m._data[self._mverts], self.h._data[self._mverts]])
        elif ctx.n_mverts:
            gates = np.zeros((3, ctx.n_mverts))
        ctx.set_state(self._pack_u(), gates)
        self._host_stale = False

    def _mark_device_newer(self):
        self._host_stale = True

    def _sync_host(self):
        """Refresh wh / phi_m_prev / gates from the device (KNPEMIx_solver.py:451-468 done lazily)."""
        if not self._host_stale or self._ctx is None:
            return
        self._host_stale = False
        u, g = self._ctx.get_state()
        self._unpack_u(u)
        N = self.N_ions
        self.phi_m_prev._data[:] = self.wh[0][N]._data - self.wh[1][N]._data
        if self.gating_variables:
            self.n._data[self._mverts], self.m._data[self._mverts], self.h._data[self._mverts] = g[0], g[1], g[2]

    # ------------------------------------------------------------------ functionals
    def l2_norm_squared(self, function, tags):
        """Local (this rank's) integral of function^2 over the cells tagged `tags`; all-reduce with
        comm.allreduce(..., op=MPI.SUM) like the reference drivers do."""
        tags = tuple(np.atleast_1d(tags).tolist())
        ctx = self._require_context()
        for s in range(2):
            for f in range(4):
                if self.wh[s][f] is function:
                    in_sub = all(t in self.intra_tags for t in tags) if s == 0 else all(t == self.extra_tag[0] for t in tags)
                    if in_sub:
                        return ctx.l2_norm_sq(s, f, tags)
        # generic host path for any other field/tag combination (post-processing, not the hot path)
        self._sync_host()
        m = self.mesh
        sel = np.isin(m.cell_tags, np.asarray(tags))
        if m.cell_owned is not None:
            sel &= m.cell_owned.astype(bool)
        cells = m.cells[sel]
        uc = function._data[cells]
        return float((self._cell_volumes(cells) * np.einsum("ca,ab,cb->c", uc, _mesh.reference_mass(m.gdim, m.degree), uc)).sum())

    def l2_norm(self, function, tags):
        return float(np.sqrt(self.comm.allreduce(self.l2_norm_squared(function, tags), op=MPI.SUM)))

    def conservation(self):
        """The numbers ProblemKNPEMI.print_conservation prints (KNPEMIx_problem.py:807-843), computed on the device:
        total amount of every ion over both subdomains and, per intracellular tag, volume, membrane area and charge
        (N_Na + N_K - N_Cl) F.  Returns {"totals": {ion: mol}, "cells": {tag: {"volume", "area", "charge"}}}."""
        ctx = self._require_context()
        red = lambda v: float(self.comm.allreduce(float(v), op=MPI.SUM))
        itags, etag = list(self.intra_tags), [self.extra_tag[0]]
        names = [ion["name"] if isinstance(ion, dict) and "name" in ion else n for ion, n in zip(self.ion_list, ("Na", "K", "Cl"))]
        totals = {n: red(ctx.integral(0, k, itags) + ctx.integral(1, k, etag)) for k, n in enumerate(names)}
        cells = {}
        mtags = set(int(t) for t in np.atleast_1d(self.gamma_tags)) if hasattr(self, "gamma_tags") else set()
        for tag in itags:
            amount = [red(ctx.integral(0, k, [tag])) for k in range(3)]
            cells[int(tag)] = {"volume": red(ctx.integral(0, 0, [tag], power=0)),
                               "area": red(ctx.membrane_area(tag)) if (not mtags or int(tag) in mtags) else 0.0,
                               "charge": (amount[0] + amount[1] - amount[2]) * float(self.F.value)}
        return {"totals": totals, "cells": cells}

    def stimulus_current(self):
        """Total stimulus current int stim_expr dS(stimulus_tags) at the current time and state, like
        SolverKNPEMI.save_png accumulates it in stim_t (KNPEMIx_solver.py:578-610); computed on the device."""
        ctx = self._require_context()
        return float(self.comm.allreduce(ctx.stimulus_current(float(self.t.value)), op=MPI.SUM))

    def print_conservation(self):
        """KNPEMIx_problem.py:807-843."""
        c = self.conservation()
        self._print(f"Time {float(self.t.value) * 1e3:.2f} ms")
        for n, label in zip(c["totals"], ("Na+", "K+ ", "Cl-")):
            self._print(f"Total {label} concentration: {c['totals'][n]:.2e} mol")
        for tag, v in c["cells"].items():
            self._print(f"  Intra tag {tag}: Volume = {v['volume']:.2e} m^3, Area = {v['area']:.2e} m^2, Charge = {v['charge']:.2e} C")
