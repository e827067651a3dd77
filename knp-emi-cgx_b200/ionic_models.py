"""Membrane models: host-side mirror of src/CGx/KNPEMI/KNPEMIx_ionic_model.py.

In the reference every model returns a UFL expression (``_eval``) that FFCx compiles into the facet
kernels.  Here a model is a *selector*: it contributes a flag per membrane tag, and the CUDA facet kernel
(csrc/assembly.cu::facet_kernel) evaluates the closed-form currents of the selected models at the facet
quadrature points.  Constructor signatures, ``tags`` handling, ``__str__`` and the HH time-stepping
attributes follow the reference (file:line in each docstring).
"""
from abc import ABC, abstractmethod
import numpy as np

from . import lib as _lib


class IonicModel(ABC):
    """KNPEMIx_ionic_model.py:11-48."""
    flag = 0

    def __init__(self, KNPEMIx_problem, tags: tuple = None):
        self.problem = KNPEMIx_problem
        self.tags = tags
        if self.tags is None:
            self.tags = self.problem.gamma_tags
        if isinstance(self.tags, (int, np.integer)):
            self.tags = (int(self.tags),)
        self.tags = tuple(int(t) for t in self.tags)

    @abstractmethod
    def _init(self):
        pass

    def _eval(self, ion_idx):
        """The reference returns a UFL expression here; the B200 path evaluates the current inside the
        CUDA facet kernel, so the host object only carries the selection."""
        raise NotImplementedError("channel currents are evaluated on the device (facet_kernel); "
                                  "use Solver/Context.assemble and inspect b")


class PassiveModel(IonicModel):
    """KNPEMIx_ionic_model.py:77-91: I_ch,k = phi_m."""
    flag = _lib.MODEL_PASSIVE

    def _init(self):
        pass

    def __str__(self):
        return "Passive model"


class KirNaKPumpModel(IonicModel):
    """KNPEMIx_ionic_model.py:93-222 (glial Kir4.1 + Na/K/ATPase)."""
    flag = _lib.MODEL_KIRNA
    rho_pump_val = 1.1 * 1.12e-6
    P_Na_i_val = 10.0
    P_K_e_val = 1.5

    def _init(self):
        pass

    def __str__(self):
        return "Na/K/ATPase pump with passive inward-rectifying K current"


class GlialCotransporters(IonicModel):
    """KNPEMIx_ionic_model.py:224-298 (KCC1/NKCC1)."""
    flag = _lib.MODEL_GLIAL_CT

    def _init(self):
        pass

    def __str__(self):
        return "KCC1/NKCC1 Cotransporters"


class NeuronalCotransporters(IonicModel):
    """KNPEMIx_ionic_model.py:300-369 (KCC2/NKCC1)."""
    flag = _lib.MODEL_NEURONAL_CT

    def _init(self):
        pass

    def __str__(self):
        return "KCC2/NKCC1 Cotransporters"


class ATPPump(IonicModel):
    """KNPEMIx_ionic_model.py:371-424."""
    flag = _lib.MODEL_ATP

    def _init(self):
        pass

    def __str__(self):
        return "Na/K/ATPase pump"


class HodgkinHuxley(IonicModel):
    """KNPEMIx_ionic_model.py:426-674.  Gates n, m, h live on the membrane vertices of the device context;
    ``update_gating_variables`` runs the CUDA gate kernel (Rush-Larsen, `time_steps_ODE` frozen-coefficient
    sub-steps)."""
    flag = _lib.MODEL_HH

    def __init__(self, KNPEMIx_problem, tags: tuple = None, use_Rush_Larsen: bool = True, time_steps_ODE: int = 25):
        super().__init__(KNPEMIx_problem, tags)
        self.use_Rush_Larsen = use_Rush_Larsen
        self.time_steps_ODE = time_steps_ODE
        self.dt_ode = KNPEMIx_problem.dt.value / self.time_steps_ODE
        self.T_stim = KNPEMIx_problem.T_stim_val
        self.t_mod = 0.0

    def __str__(self):
        return "Hodgkin-Huxley"

    def _init(self):
        p = self.problem
        p.n_init_value, p.m_init_value, p.h_init_value = p.n_init.value, p.m_init.value, p.h_init.value
        p.ode_substeps = int(self.time_steps_ODE)
        p.rush_larsen = bool(self.use_Rush_Larsen)
        p._print(f"Initial n = {p.n_init.value}\nm = {p.m_init.value}\nh = {p.h_init.value}")

    def update_t_mod(self, tol: float = 1e-12):
        self.t_mod = float(np.mod(self.problem.t.value + tol, self.T_stim))

    def update_gating_variables(self):
        ctx = self.problem._require_context()
        ctx.gate_step()
        self.problem._mark_device_newer()
