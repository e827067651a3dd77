"""knp-emi-cgx_b200: B200-native implementation of CGx's per-timestep KNP-EMI hot path.

Public surface = the reference's own (src/CGx/KNPEMI): ProblemKNPEMI, SolverKNPEMI and the ionic models,
driven by the same YAML schema.  The compute path is libknpemi_b200.so (hand-written sm_100a CUDA kernels,
C ABI in include/knpemi_b200.h); importing this package does not require a GPU, creating a device context does.
The directory name contains '-', so import it with ``importlib.import_module("knp-emi-cgx_b200")`` or through
the ``cgx_b200`` alias module / the ``CGx`` drop-in shim at the repository root.
"""
from .comm import Comm, MPI
from .ionic_models import (IonicModel, PassiveModel, KirNaKPumpModel, GlialCotransporters,
                           NeuronalCotransporters, ATPPump, HodgkinHuxley)
from .problem import ProblemKNPEMI, Constant, Function
from .solver import SolverKNPEMI
from . import lib, mesh

__all__ = ["ProblemKNPEMI", "SolverKNPEMI", "IonicModel", "PassiveModel", "KirNaKPumpModel", "GlialCotransporters",
           "NeuronalCotransporters", "ATPPump", "HodgkinHuxley", "Comm", "MPI", "Constant", "Function", "lib", "mesh"]
