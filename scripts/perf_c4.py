"""Development perf probe for the 3D configuration (C4-like): setup phases, per-step phases, kernel rooflines."""
import sys, os, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgx_b200 as kb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
models = sys.argv[3] if len(sys.argv) > 3 else "passive"
cfgdir = os.path.join(os.path.dirname(kb.__file__), "configs")
txt = open(os.path.join(cfgdir, "c4_cube120_cells64_passive.yaml")).read().replace("N: 120", f"N: {N}")
tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False); tmp.write(txt); tmp.close()
t0 = time.time(); p = kb.ProblemKNPEMI(tmp.name, verbose=False); print("problem (mesh) s", time.time() - t0, flush=True)
p.set_initial_conditions()
p.init_ionic_models([kb.PassiveModel(p)] if models == "passive" else [kb.NeuronalCotransporters(p), kb.HodgkinHuxley(p), kb.ATPPump(p)])
t0 = time.time(); p.setup_variational_form(); print("context s", time.time() - t0, flush=True)
ctx = p._ctx
print("rows", ctx.n_rows, "nnz", ctx.nnz, "mverts", ctx.n_mverts, "mfacets", ctx.sizes.n_mfacets, "maxdeg", ctx.sizes.max_deg, ctx.sizes.max_gdeg, flush=True)
p.solver_config['view_ksp'] = False
s = kb.SolverKNPEMI(p, p.solver_config)
t0 = time.time(); s.setup_solver(); p.setup_preconditioner(True); ctx.pc_setup(s.opts); print("P + AMG setup s", time.time() - t0, [(a.shape[0], a.nnz) for a in ctx.amg_levels()], flush=True)
ctx.set_time(0.0, 0)
for i in range(steps):
    info = ctx.step(s.opts); print("step", i, "its", info.iterations, ctx.last_timings(), flush=True)
st = torch.cuda.Stream(); sp = st.cuda_stream
x = torch.randn(ctx.n_cols, dtype=torch.float64, device="cuda"); y = torch.empty(ctx.n_rows, dtype=torch.float64, device="cuda")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t_asm = timeit(lambda: ctx.assemble(1e-4, stream=sp)); t_spmv = timeit(lambda: ctx.spmv(x.data_ptr(), y.data_ptr(), stream=sp)); t_pc = timeit(lambda: ctx.pc_apply(x.data_ptr(), y.data_ptr(), stream=sp))
m = p.mesh
B_spmv = 12 * ctx.nnz + 20 * ctx.n_rows
B_asm = 8 * ctx.nnz + 16 * ctx.n_rows + 8 * m.gdim * m.x.shape[0] + (4 * (m.gdim + 1) + 4) * m.cells.shape[0] + 32 * ctx.n_mverts + 16 * ctx.sizes.n_mfacets
print(f"assemble {t_asm:.3f} ms -> {B_asm / t_asm / 1e6:.0f} GB/s ; spmv {t_spmv:.3f} ms -> {B_spmv / t_spmv / 1e6:.0f} GB/s ; pc_apply {t_pc:.3f} ms", flush=True)
print("facet+rows timings of last assemble:", ctx.last_timings())
