"""Short launch sequence for ncu: one gate step, one assembly (facet + rows), one SpMV on A, one AMG V-cycle,
on the bench workload at size N (default 2048)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgx_b200 as kb
import bench
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
p = kb.ProblemKNPEMI(bench.workload_yaml(kb, N), verbose=False)
p.set_initial_conditions(); p.init_ionic_models([kb.NeuronalCotransporters(p), kb.HodgkinHuxley(p), kb.ATPPump(p)])
p.setup_variational_form(); p.solver_config["view_ksp"] = False
s = kb.SolverKNPEMI(p, solver_config=p.solver_config); s.setup_solver(); p.setup_preconditioner(True)
ctx = s.ctx; ctx.pc_setup(s.opts)
x = torch.randn(ctx.n_cols, dtype=torch.float64, device="cuda"); y = torch.empty(ctx.n_rows, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for rep in range(2):
    ctx.gate_step(); ctx.assemble(1e-4); ctx.spmv(x.data_ptr(), y.data_ptr()); ctx.pc_apply(x.data_ptr(), y.data_ptr())
ctx.to_host(y.data_ptr(), 1)
print("probe done", ctx.n_rows, ctx.nnz)
