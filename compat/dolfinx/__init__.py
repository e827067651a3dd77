"""Stand-in for the part of dolfinx the reference's driver scripts touch after the solve: dolfinx.fem.form and
dolfinx.fem.assemble_scalar on ``ufl.inner(u, v) * problem.dx(tag)``."""
import numpy as _np
from . import fem  # noqa: F401

default_scalar_type = _np.float64
