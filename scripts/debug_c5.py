"""Debug probe: where do non-finite values first appear for a workload (assembly, preconditioner, first solve)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, cgx_b200 as kb
wl, n = sys.argv[1], int(sys.argv[2])
p, s = bench.build_problem(kb, wl, n, 0)
ctx = s.ctx
print("rows", ctx.n_rows, "nnz", ctx.nnz, "mverts", ctx.n_mverts, "mfacets", ctx.sizes.n_mfacets, "max_deg", ctx.sizes.max_deg, ctx.sizes.max_gdeg)
print("stimulus area", getattr(p, "stimulus_area", None))
u, g = ctx.get_state()
print("state finite", np.isfinite(u).all(), np.isfinite(g).all(), "u min/max", u.min(), u.max())
ctx.assemble(p.dt.value)
A, b, P = ctx.values_host()
print("A finite", np.isfinite(A).all(), "b finite", np.isfinite(b).all(), "P finite", np.isfinite(P).all())
if not np.isfinite(b).all():
    bad = np.flatnonzero(~np.isfinite(b))
    n0, n1 = ctx.n_own
    print("bad b rows", bad.size, "fields", np.unique(np.where(bad < 4 * n0, bad // max(n0, 1), 4 + (bad - 4 * n0) // max(n1, 1))))
if not np.isfinite(A).all():
    ip, ix = ctx.csr()
    badr = np.unique(np.searchsorted(ip, np.flatnonzero(~np.isfinite(A)), side="right") - 1)
    print("bad A rows", badr.size, badr[:10])
x = torch.ones(ctx.n_cols, dtype=torch.float64, device="cuda"); y = torch.zeros(ctx.n_rows, dtype=torch.float64, device="cuda")
ctx.pc_apply(x.data_ptr(), y.data_ptr()); torch.cuda.synchronize()
print("pc(ones) finite", bool(torch.isfinite(y).all()), float(y.abs().max()))
n0 = ctx._lib.knp_amg_part_levels(ctx.h, 0)
lv = ctx.amg_levels()
print("ion hierarchy", [(a.shape[0], a.nnz) for a in lv[:n0]], "potential hierarchy", [(a.shape[0], a.nnz) for a in lv[n0:]])
for a in lv:
    d = a.diagonal()
    print("  level", a.shape[0], "diag min", d.min(), "finite", np.isfinite(a.data).all())
ctx.set_time(0.0, 0)
ctx.gate_step(); 
u, g = ctx.get_state()
print("after gate: gates finite", np.isfinite(g).all(), g.min(), g.max())
ctx.assemble(p.dt.value)
A, b, P = ctx.values_host()
print("after gate: A finite", np.isfinite(A).all(), "b finite", np.isfinite(b).all())
d = ctx.dev_ptrs()
bt = torch.zeros(ctx.n_rows, dtype=torch.float64, device="cuda")
ctx._lib.knp_copy(ctx.h, bt.data_ptr(), d["b"], ctx.n_rows * 8, 3)
ctx.pc_apply(d["b"], y.data_ptr()); torch.cuda.synchronize()
print("pc(b) finite", bool(torch.isfinite(y).all()), float(y.abs().max()), "b absmax", float(bt.abs().max()))
yh = y.cpu().numpy()
if not np.isfinite(yh).all():
    bad = np.flatnonzero(~np.isfinite(yh)); n0, n1 = ctx.n_own
    print("bad pc rows", bad.size, "fields", np.unique(np.where(bad < 4 * n0, bad // max(n0, 1), 4 + (bad - 4 * n0) // max(n1, 1))))
try:
    info = ctx.step(s.opts); print("step ok", info.iterations)
except Exception as e:
    print("step failed:", e)
