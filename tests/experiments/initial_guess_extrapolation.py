"""Design experiment (CPU, oracle-side): previous solution vs linear / quadratic extrapolation in time as GMRES initial
guess.  Result (N = 256, 8 steps): 29 22 20 18 18 16 16 15 vs 29 29 20 18 17 16 14 13 -- not worth the extra state vector.
Usage: python initial_guess_extrapolation.py 256 8
"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from pc_experiment import *
from oracle.amg import SchurPC
n = int(sys.argv[1]); steps = int(sys.argv[2])
for mode in ("previous", "linear", "quadratic"):
    o = make(n); x = o.pack(); hist = [x.copy()]; its = []; pc = None
    for i in range(steps):
        o.t += o.p.dt; o.gate_update(); A, b = o.assemble(o.t); ns = o.nullspace()
        if i == 0: b = b - ns * (ns @ b)
        if pc is None: pc = SchurPC(o)
        if mode == "linear" and len(hist) >= 2: x0 = 2 * hist[-1] - hist[-2]
        elif mode == "quadratic" and len(hist) >= 3: x0 = 3 * hist[-1] - 3 * hist[-2] + hist[-3]
        elif mode == "quadratic" and len(hist) == 2: x0 = 2 * hist[-1] - hist[-2]
        else: x0 = hist[-1]
        x, k = o.solve_gmres(A, b, x0, ns, pc, 1e-9); o.unpack(x); its.append(k); hist.append(x.copy())
    print(f"{mode:10s} its {its}", flush=True)
