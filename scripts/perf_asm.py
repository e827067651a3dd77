"""Development probe: time only the assembly kernels (facet + rows) on the 2D (C3) and 3D (C4) workloads."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cgx_b200 as kb
cfgdir = os.path.join(os.path.dirname(kb.__file__), "configs")
def run(cfgname, repl, models):
    txt = open(os.path.join(cfgdir, cfgname)).read()
    for a, b in repl: txt = txt.replace(a, b)
    tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False); tmp.write(txt); tmp.close()
    p = kb.ProblemKNPEMI(tmp.name, verbose=False); p.set_initial_conditions(); p.init_ionic_models(models(p)); p.setup_variational_form()
    ctx = p._ctx
    st = torch.cuda.Stream(); sp = st.cuda_stream
    for _ in range(3): ctx.assemble(1e-4, stream=sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10): ctx.assemble(1e-4, stream=sp)
    e1.record(st); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    m = p.mesh
    B = 8 * ctx.nnz + 16 * ctx.n_rows + 8 * m.gdim * m.x.shape[0] + (4 * (m.gdim + 1) + 4) * m.cells.shape[0] + 32 * ctx.n_mverts + 16 * ctx.sizes.n_mfacets
    print(f"{cfgname}: assemble {t:.3f} ms -> {B / t / 1e6:.0f} GB/s ({B / t / 1e6 / 6538:.3f} of measured peak)", flush=True)
    p._ctx.close() if hasattr(p._ctx, "close") else None
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("2d", "both"):
    run("c3_square2048_cells64.yaml", [], lambda p: [kb.NeuronalCotransporters(p), kb.HodgkinHuxley(p), kb.ATPPump(p)])
if which == "c5":      # BASELINE C5 per-GPU size: N = 128 (1/8 of the N = 256 mesh), plate-stack cells, HH + ATP + KCC2 with stimulus
    n = sys.argv[2] if len(sys.argv) > 2 else "128"
    run("c5_cube256_tissue512_hh.yaml", [("N: 256", f"N: {n}")], lambda p: [kb.NeuronalCotransporters(p), kb.HodgkinHuxley(p), kb.ATPPump(p)])
if which in ("3d", "both"):
    run("c4_cube120_cells64_passive.yaml", [], lambda p: [kb.PassiveModel(p)])
