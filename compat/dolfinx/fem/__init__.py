"""dolfinx.fem.form / assemble_scalar for ``ufl.inner(u, v) * problem.dx(tags)`` (tests/KNPEMI/*.py:45-51 of the reference):
this rank's part of the integral; the caller all-reduces it like the reference does."""
import numpy as np


def form(f, **kwargs):
    return f


def assemble_scalar(f):
    u, v, measure = f.integrand.a, f.integrand.b, f.measure
    problem, tags = measure.problem, measure.tags
    if tags is None:
        raise ValueError("assemble_scalar: give the measure a subdomain tag, e.g. problem.dx(1)")
    if u is v:
        return problem.l2_norm_squared(u, tags)          # device functional (knp_l2_norm_sq) for the solution fields
    # generic P1 mass-matrix product on the host for two different fields
    problem._sync_host()
    m = problem.mesh
    sel = np.isin(m.cell_tags, np.asarray(tags))
    if m.cell_owned is not None:
        sel &= m.cell_owned.astype(bool)
    cells = m.cells[sel]
    xx = m.x[cells]
    d = m.gdim
    vol = np.abs(np.linalg.det(xx[:, 1:] - xx[:, :1])) / (2.0 if d == 2 else 6.0)
    uc, vc = u._data[cells], v._data[cells]
    return float((vol / ((d + 1) * (d + 2)) * ((uc * vc).sum(1) + uc.sum(1) * vc.sum(1))).sum())
