"""Design experiment (CPU, oracle-side): damped Jacobi vs Chebyshev smoothing inside the Schur preconditioner.
Result (N = 256): Jacobi V(1,1) 29/22/20, Chebyshev(2) 22/15/14, Jacobi V(2,2) 24/17/16 iterations -- 30 % fewer iterations
for 70 % more work per cycle: no gain.  Usage: python smoother_comparison.py 256 3
"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pc_experiment import *
import oracle.amg as oamg
from oracle.amg import SchurPC
class ChebAMG(oamg.SAAMG):
    deg = 2; frac = 4.0; nu = 1
    def _smooth(self, lv, x, b):
        A, dinv, rho = lv["A"], lv["dinv"], lv["rho"]
        if self.deg == 1:
            for _ in range(self.nu): x = x + (4.0 / 3.0 / rho) * dinv * (b - A @ x)
            return x
        lmax, lmin = 1.05 * rho, rho / self.frac
        th, de = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        sigma = th / de; rhok = 1.0 / sigma
        r = dinv * (b - A @ x); d = r / th; x = x + d
        for _ in range(self.deg - 1):
            rhok1 = 1.0 / (2 * sigma - rhok); r = dinv * (b - A @ x)
            d = rhok1 * rhok * d + (2 * rhok1 / de) * r; x = x + d; rhok = rhok1
        return x
n = int(sys.argv[1]); steps = int(sys.argv[2])
for nm, deg, frac, nu in [("jacobi V(1,1)", 1, 0, 1), ("cheb2 [rho/4]", 2, 4.0, 1), ("cheb2 [rho/8]", 2, 8.0, 1), ("cheb3 [rho/8]", 3, 8.0, 1), ("jacobi V(2,2)", 1, 0, 2)]:
    def f(o, A):
        ChebAMG.deg, ChebAMG.frac, ChebAMG.nu = deg, frac, nu
        SA = oamg.SAAMG; oamg.SAAMG = ChebAMG
        try: pc = SchurPC(o)
        finally: oamg.SAAMG = SA
        return pc
    run(make(n), f, steps, label=nm)
