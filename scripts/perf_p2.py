"""P2 element path (fem_order: 2) on the GPU: assembly time and time steps on a 2D (C3-like, HH + ATP + KCC2) and a 3D
(C4-like, passive) tissue block.  Prints one JSON line per case.   python scripts/perf_p2.py [2d|3d|both] [N2d] [N3d] [--asm-only]"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import cgx_b200 as kb  # noqa: E402

cfgdir = os.path.join(os.path.dirname(kb.__file__), "configs")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
which = args[0] if args else "both"
N2, N3 = (int(args[1]) if len(args) > 1 else 512), (int(args[2]) if len(args) > 2 else 32)
asm_only = "--asm-only" in sys.argv


def run(tag, cfgname, repl, models, steps=3):
    txt = open(os.path.join(cfgdir, cfgname)).read()
    for a, b in repl:
        txt = txt.replace(a, b)
    txt = txt.replace('problem_type: "KNP-EMI"', 'problem_type: "KNP-EMI"\nfem_order: 2')
    tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
    tmp.write(txt)
    tmp.close()
    t0 = time.time()
    p = kb.ProblemKNPEMI(tmp.name, verbose=False)
    p.set_initial_conditions()
    p.init_ionic_models(models(p))
    p.setup_variational_form()
    ctx = p._ctx
    t_setup = time.time() - t0
    st = torch.cuda.Stream()
    sp = st.cuda_stream
    for _ in range(2):
        ctx.assemble(1e-4, stream=sp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        ctx.assemble(1e-4, stream=sp)
    e1.record(st)
    torch.cuda.synchronize()
    asm_ms = e0.elapsed_time(e1) / 5
    out = dict(case=tag, fem_order=2, nodes=int(p.mesh.x.shape[0]), cells=int(p.mesh.cells.shape[0]), dofs=int(ctx.n_rows),
               nnz=int(ctx.nnz), membrane_facets=int(ctx.sizes.n_mfacets), assembly_ms=asm_ms,
               matrix_write_GBps=8 * ctx.nnz / asm_ms / 1e6, setup_context_s=t_setup)
    if not asm_only:
        p.solver_config["view_ksp"] = False
        s = kb.SolverKNPEMI(p, solver_config=p.solver_config)
        t0 = time.time()
        s.setup_solver(); p.setup_preconditioner(True); s.ctx.pc_setup(s.opts); s.ctx.set_time(0.0, 0)
        out["setup_pc_s"] = time.time() - t0
        its, ms, asm = [], [], []
        for i in range(steps):
            info = s.ctx.step(s.opts)
            tm = s.ctx.last_timings()
            its.append(int(info.iterations)); ms.append(float(tm["total"])); asm.append(float(tm["facet"] + tm["rows"]))
        out.update(iterations_per_step=its, ms_per_step=ms, assembly_in_step_ms=asm)
    print(json.dumps(out), flush=True)
    ctx.close()


if which in ("2d", "both"):
    run("2d", "c3_square2048_cells64.yaml", [("N: 2048", f"N: {N2}")],
        lambda p: [kb.NeuronalCotransporters(p), kb.HodgkinHuxley(p), kb.ATPPump(p)])
if which in ("3d", "both"):
    run("3d", "c4_cube120_cells64_passive.yaml", [("N: 120", f"N: {N3}")], lambda p: [kb.PassiveModel(p)])
