"""Development smoke check on a GPU box: C1/C2 structure, values, time loop against the oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgx_b200 as kb
from oracle.fixtures import unit_square
from oracle.knpemi import KNPEMIOracle, OracleParams

cfgdir = os.path.join(os.path.dirname(kb.__file__), "configs")

def make(cfg):
    p = kb.ProblemKNPEMI(os.path.join(cfgdir, cfg), verbose=False)
    HH, ATP, NCT = kb.HodgkinHuxley(p), kb.ATPPump(p), kb.NeuronalCotransporters(p)
    p.set_initial_conditions(); p.init_ionic_models([NCT, HH, ATP]); p.setup_variational_form()
    p.solver_config['view_ksp'] = False
    return p

p = make("c1_square32_direct.yaml")
ctx = p._ctx
print("sizes", ctx.n_rows, ctx.nnz, ctx.nnz_P, ctx.n_own, ctx.n_mverts)
o = KNPEMIOracle(unit_square(32), OracleParams(), [("NeuronalCT", None), ("HH", None), ("ATP", None)])
ip, ix = ctx.csr()
# one assembly at t = dt after a gate step (as the solver loop does)
o.t += o.p.dt; o.gate_update(); A, b = o.assemble(o.t)
print("indptr equal", np.array_equal(ip, A.indptr), "indices equal", np.array_equal(ix, A.indices))
ctx.gate_step(); ctx.assemble(o.p.dt)
Av, bv, _ = ctx.values_host()
u, g = ctx.get_state()
print("gates diff", np.abs(g - o.gates[:, o.mverts]).max())
scale = np.maximum.reduceat(np.abs(A.data), A.indptr[:-1])
rowscale = np.repeat(scale, np.diff(A.indptr))
print("A max rel diff (row scale)", np.abs(Av - A.data).max() / 1, (np.abs(Av - A.data) / rowscale).max())
print("b max rel diff", (np.abs(bv - b) / np.abs(b).max()).max(), np.abs(bv - b).max(), np.abs(b).max())
for cfg, conv in [("c1_square32_direct.yaml", "direct"), ("c2_square32_iterative.yaml", "gmres")]:
    p = make(cfg)
    s = kb.SolverKNPEMI(p, p.solver_config)
    t0 = time.time(); s.solve(); print("solve wall", time.time() - t0)
    li = p.l2_norm(p.wh[0][3], 1); le = p.l2_norm(p.wh[1][3], 2)
    print(cfg, li, le, getattr(s, "iterations", None), s.tot_its)
    gold = dict(direct=(2.6337161145147203e-08, 1.5258564901943312e-08), gmres=(3.510994056704844e-08, 6.369472309249516e-11))[conv]
    print("  vs golden", (li - gold[0]) / gold[0], (le - gold[1]) / gold[1])
