"""Reference import path src/CGx/KNPEMI/KNPEMIx_problem.py -> B200-native ProblemKNPEMI."""
from cgx_b200.problem import ProblemKNPEMI  # noqa: F401
