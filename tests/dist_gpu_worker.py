"""Multi-GPU worker (torchrun, NCCL): C2 fixture split over WORLD_SIZE GPUs; per-field norms after 3 steps are
compared with the single-process CPU oracle on rank 0.  Used by tests/test_gpu_multi.py and by hand:
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_gpu_worker.py"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cgx_b200 as kb                                            # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = os.path.join(os.path.dirname(kb.__file__), "configs", "c2_square32_iterative.yaml")
p = kb.ProblemKNPEMI(cfg, verbose=False, device=local)
p.set_initial_conditions()
p.init_ionic_models([kb.NeuronalCotransporters(p), kb.HodgkinHuxley(p), kb.ATPPump(p)])
p.setup_variational_form()
p.solver_config["view_ksp"] = False
s = kb.SolverKNPEMI(p, solver_config=p.solver_config)
s.ksp_rtol = 1e-12
s.time_steps = 3
s.solve()
norms = [p.l2_norm(p.wh[sd][f], 1 if sd == 0 else 2) for sd in range(2) for f in range(4)]
if p.comm.rank == 0:
    from oracle.fixtures import unit_square
    from oracle.knpemi import KNPEMIOracle, OracleParams
    import scipy.sparse.linalg as spla
    o = KNPEMIOracle(unit_square(32), OracleParams(), [("NeuronalCT", None), ("HH", None), ("ATP", None)])
    lu = spla.splu(o.assemble_P().tocsc())
    x = o.pack()
    for i in range(3):
        _, _, x, _ = o.step("gmres", lambda v: lu.solve(v), 1e-13, x, first=(i == 0))
    ref = [o.l2_norm(o.c[sd][f] if f < 3 else o.phi[sd], 1 if sd == 0 else 2) for sd in range(2) for f in range(4)]
    # phi_e (index 7) is ~1e-3 of phi_i: relative to the potential scale, as in tests/test_gpu_parity.py
    scale = list(ref)
    scale[7] = max(ref[7], ref[3])
    err = max(abs(a - b) / sc for a, b, sc in zip(norms, ref, scale))
    print(f"ranks {p.comm.size} iterations {s.iterations} max rel norm err {err:.3e} transport "
          f"{'peer' if s.ctx.peer_direct() else 'nccl'}", flush=True)
    assert err < 1e-8, (norms, ref)
    print("MULTI_GPU_OK", flush=True)
dist.destroy_process_group()
