"""Facet quadrature tables (oracle side; test infrastructure).

The reference integrates every dS term with ``quadrature_degree = 10``
(/root/reference/src/CGx/utils/mixed_dim_problem.py:732-733).  basix 0.9.0 (not
vendored, not installable here) maps that to a 6-point Gauss-Jacobi(0,0) =
Gauss-Legendre rule on an interval and to a 25-point Xiao-Gimbutas rule on a
triangle.  The interval rule is reproduced exactly.  The Xiao-Gimbutas table is
not available offline, so triangle facets use a collapsed-coordinate (Duffy)
Gauss-Legendre x Gauss-Jacobi(1,0) rule with 6 x 6 points, which is also exact
to degree >= 10; the survey measured <= 2e-11 relative sensitivity of the
golden norms to the choice among degree-10-exact rules.
The same tables (as data) are handed to the CUDA facet kernel.
"""
import numpy as np
from scipy.special import roots_jacobi


def interval_rule(npts=6):
    """Barycentric points (npts,2) and weights summing to 1 on a segment."""
    t, w = np.polynomial.legendre.leggauss(npts)
    s = 0.5 * (t + 1.0)
    return np.stack([1.0 - s, s], 1), 0.5 * w


def triangle_rule(npts=6):
    """Barycentric points (npts^2,3) and weights summing to 1 on a triangle."""
    tu, wu = np.polynomial.legendre.leggauss(npts)
    u = 0.5 * (tu + 1.0)
    wu = 0.5 * wu
    tv, wv = roots_jacobi(npts, 1.0, 0.0)      # weight (1-t) on [-1,1]
    v = 0.5 * (tv + 1.0)
    wv = 0.25 * wv                              # -> weight (1-v) on [0,1], integrates to 1/2
    U, V = np.meshgrid(u, v, indexing="ij")
    W = np.outer(wu, wv)
    l1 = V.ravel()
    l2 = (U * (1.0 - V)).ravel()
    l0 = 1.0 - l1 - l2
    w = W.ravel() * 2.0                         # reference triangle has area 1/2 -> weights sum to 1
    return np.stack([l0, l1, l2], 1), w


def facet_rule(gdim):
    return interval_rule(6) if gdim == 2 else triangle_rule(6)
