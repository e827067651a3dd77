"""Diagnostic: is the preconditioner application of the 3D P2 case a fixed linear operator?  Builds the N = 32 P2 tissue block,
sets the preconditioner up, prints checksums of the hierarchy operators and of z = B r for repeated applications (same r),
and the GMRES iterations of two steps.   python scripts/diag_p2_pc.py [N]"""
import hashlib
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import cgx_b200 as kb  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfgdir = os.path.join(os.path.dirname(kb.__file__), "configs")
txt = open(os.path.join(cfgdir, "c4_cube120_cells64_passive.yaml")).read().replace("N: 120", f"N: {N}")
if order == 2:
    txt = txt.replace('problem_type: "KNP-EMI"', 'problem_type: "KNP-EMI"\nfem_order: 2')
tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
tmp.write(txt)
tmp.close()
p = kb.ProblemKNPEMI(tmp.name, verbose=False)
p.set_initial_conditions()
p.init_ionic_models([kb.PassiveModel(p)])
p.setup_variational_form()
if os.environ.get("DIAG_PRE_ASSEMBLE"):
    # what scripts/perf_p2.py does before the solver exists: assemblies on a foreign stream
    stt = torch.cuda.Stream()
    for _ in range(int(os.environ["DIAG_PRE_ASSEMBLE"])):
        p._ctx.assemble(1e-4, stream=stt.cuda_stream)
    torch.cuda.synchronize()
    Av, bv, _ = p._ctx.values_host()
    print("pre-assembled A, b:", hashlib.md5(Av.tobytes()).hexdigest()[:8], hashlib.md5(bv.tobytes()).hexdigest()[:8], flush=True)
p.solver_config["view_ksp"] = False
s = kb.SolverKNPEMI(p, solver_config=p.solver_config)
s.setup_solver(); p.setup_preconditioner(True); s.ctx.pc_setup(s.opts); s.ctx.set_time(0.0, 0)
ctx = s.ctx
md5 = lambda a: hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()[:8]
_, _, Pv = ctx.values_host()
print("P~ values", md5(Pv), "u", md5(ctx.get_state()[0]), flush=True)
for part in (0, 1):
    print("part", part, [(a.shape[0], a.nnz, md5(a.data), md5(a.indices)) for a in ctx.amg_levels(part)], flush=True)
n = ctx.n_rows
r = torch.from_numpy(np.random.default_rng(0).standard_normal(n)).cuda()
zs = []
for k in range(8):
    z = torch.zeros(n, dtype=torch.float64, device="cuda")
    ctx.pc_apply(r.data_ptr(), z.data_ptr())
    torch.cuda.synchronize()
    zs.append(z.cpu().numpy())
print("B r:", [md5(z) for z in zs], "max rel dev from first", max(np.abs(z - zs[0]).max() for z in zs) / np.abs(zs[0]).max(), flush=True)
# same output buffer (this is what GMRES does: the application is replayed as a CUDA graph from the second call on)
z = torch.zeros(n, dtype=torch.float64, device="cuda")
hs = []
for k in range(8):
    ctx.pc_apply(r.data_ptr(), z.data_ptr())
    torch.cuda.synchronize()
    hs.append(md5(z.cpu().numpy()))
print("B r (one buffer):", hs, flush=True)
its = []
try:
    for i in range(2):
        its.append(int(ctx.step(s.opts).iterations))
except Exception as e:
    its.append(str(e)[-60:])
print("iterations", its, "env", {k: v for k, v in os.environ.items() if k.startswith("KNP_")}, flush=True)
