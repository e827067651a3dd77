"""Reference import path src/CGx/KNPEMI/KNPEMIx_solver.py -> B200-native SolverKNPEMI."""
from cgx_b200.solver import SolverKNPEMI  # noqa: F401
