"""Minimal driver: build a workload and take a few timesteps (used under compute-sanitizer and for debugging)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, cgx_b200 as kb
wl, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 2
p, s = bench.build_problem(kb, wl, n, 0)
for i in range(steps):
    info = s.ctx.step(s.opts)
    print("step", i + 1, "iterations", info.iterations, "converged", info.converged, flush=True)
p._mark_device_newer()
print("norms", bench.field_norms(kb, p))
