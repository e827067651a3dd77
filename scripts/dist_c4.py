"""3D (C4-like) time loop on WORLD_SIZE GPUs (torchrun) or one GPU: prints iterations and per-field L2 norms so that
runs with different GPU counts can be compared (the preconditioner is partition independent, the norms must agree)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cgx_b200 as kb
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfgdir = os.path.join(os.path.dirname(kb.__file__), "configs")
txt = open(os.path.join(cfgdir, "c4_cube120_cells64_passive.yaml")).read().replace("N: 120", f"N: {N}")
tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False); tmp.write(txt); tmp.close()
p = kb.ProblemKNPEMI(tmp.name, verbose=False, device=local)
p.set_initial_conditions(); p.init_ionic_models([kb.PassiveModel(p)]); p.setup_variational_form()
p.solver_config["view_ksp"] = False
s = kb.SolverKNPEMI(p, p.solver_config); s.time_steps = steps
s.solve()
tags_i = list(range(2, 66))
norms = [np.sqrt(p.comm.allreduce(p.l2_norm_squared(p.wh[sd][f], tags_i if sd == 0 else 1), op=kb.MPI.SUM)) for sd in range(2) for f in range(4)]
if p.comm.rank == 0:
    print(f"ranks {world} N {N} iterations {s.iterations} solve ms {[round(1e3 * t, 2) for t in s.solve_time]} asm ms {[round(1e3 * t, 2) for t in s.assembly_time]}")
    print("norms", " ".join(f"{v:.12e}" for v in norms), flush=True)
if world > 1:
    dist.destroy_process_group()
