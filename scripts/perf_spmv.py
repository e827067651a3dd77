"""Development probe: time y = A x and one preconditioner application on the C3 workload (size argv[1])."""
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
ctx = s.ctx
if len(sys.argv) > 2: ctx.pc_setup(s.opts)
ctx.assemble(1e-4)
st = torch.cuda.Stream(); sp = st.cuda_stream
x = torch.randn(ctx.n_cols, dtype=torch.float64, device="cuda"); y = torch.empty(ctx.n_rows, dtype=torch.float64, device="cuda")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = timeit(lambda: ctx.spmv(x.data_ptr(), y.data_ptr(), stream=sp))
B = 12 * ctx.nnz + 20 * ctx.n_rows
print(f"spmv {t:.3f} ms -> {B / t / 1e6:.0f} GB/s ({B / t / 1e6 / 6538:.3f} of measured peak)", flush=True)
if len(sys.argv) > 2:
    print(f"pc_apply {timeit(lambda: ctx.pc_apply(x.data_ptr(), y.data_ptr(), stream=sp), 10):.3f} ms", flush=True)
