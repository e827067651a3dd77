"""Short launch sequence for ncu on a bench workload: one gate step, one assembly (facet + rows), one SpMV on A and one
preconditioner application (plain launches: run with KNP_PC_GRAPH=0), twice; prints the number of kernel launches before
the second repetition so that `ncu -s <that> -c <per repetition>` captures exactly one of each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgx_b200 as kb
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[wl][1]
p, s = bench.build_problem(kb, wl, N, 0)
ctx = s.ctx
x = torch.randn(ctx.n_cols, dtype=torch.float64, device="cuda"); y = torch.empty(ctx.n_rows, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
for rep in range(2):
    l0 = kb.lib.launch_count()
    ctx.gate_step(); ctx.assemble(1e-4); ctx.spmv(x.data_ptr(), y.data_ptr()); ctx.pc_apply(x.data_ptr(), y.data_ptr())
    ctx.to_host(y.data_ptr(), 1)
    if rep == 0:
        print("launches per repetition", kb.lib.launch_count() - l0)
        print("launches before the second repetition", kb.lib.launch_count(), "(torch adds its own randn kernel: +1)")
print("probe done", ctx.n_rows, ctx.nnz, "pc bytes", ctx.pc_bytes())
