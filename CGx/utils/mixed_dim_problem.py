"""Reference import path src/CGx/utils/mixed_dim_problem.py: the base-class role is merged into ProblemKNPEMI."""
from cgx_b200.problem import ProblemKNPEMI as MixedDimensionalProblem  # noqa: F401
