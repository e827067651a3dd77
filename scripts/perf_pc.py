"""Development probe: preconditioner application time on a workload (c3 | c4) with the hierarchy sizes and the algorithmic
bytes.  Usage: python scripts/perf_pc.py c3 2048   (env: KNP_FUSE_NNZ, KNP_TAIL_GRID, KNP_PC_GRAPH ...)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import cgx_b200 as kb
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WORKLOADS[wl][1]
world, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 8) // world))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if local != 0:
        sys.stdout = open(os.devnull, "w")
t0 = time.time()
p, s = bench.build_problem(kb, wl, n, local)
ctx = s.ctx
print(f"ranks {world} setup {time.time() - t0:.1f} s rows {ctx.n_rows} nnz {ctx.nnz} transport {'peer' if ctx.peer_direct() else 'nccl'}")
n0 = ctx._lib.knp_amg_part_levels(ctx.h, 0)
lv = ctx.amg_levels()
print("ion hierarchy", [(a.shape[0], a.nnz) for a in lv[:n0]], "potential hierarchy", [(a.shape[0], a.nnz) for a in lv[n0:]])
st = torch.cuda.Stream(); sp = st.cuda_stream
x = torch.randn(ctx.n_cols, dtype=torch.float64, device="cuda"); y = torch.empty(ctx.n_rows, dtype=torch.float64, device="cuda")
def timeit(fn, reps=20):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
l0 = kb.lib.launch_count()
ctx.pc_apply(x.data_ptr(), y.data_ptr(), stream=sp); ctx.pc_apply(x.data_ptr(), y.data_ptr(), stream=sp)
l1 = kb.lib.launch_count()
ctx.pc_apply(x.data_ptr(), y.data_ptr(), stream=sp)
print("launches per application", kb.lib.launch_count() - l1)
t = timeit(lambda: ctx.pc_apply(x.data_ptr(), y.data_ptr(), stream=sp))
B = ctx.pc_bytes()
print(f"pc_apply {t:.3f} ms ; algorithmic bytes {B / 1e9:.3f} GB -> {B / t / 1e6:.0f} GB/s = {B / t / 1e6 / 6538:.3f} of 6538")
ts = timeit(lambda: ctx.spmv(x.data_ptr(), y.data_ptr(), stream=sp))
print(f"spmv A {ts:.3f} ms")
if not os.environ.get("KNP_HALO_SKIP"):
    its = [int(ctx.step(s.opts).iterations) for _ in range(6)]
    print("iterations", its, "last step", ctx.last_timings())
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
