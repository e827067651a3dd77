"""Worker for tests/test_partition_cpu.py::test_gloo_world_size_2 (launched with torch.distributed.run, gloo)."""
import importlib
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cgx_b200 as kb                                            # noqa: E402

part = importlib.import_module("knp-emi-cgx_b200.partition")
dist.init_process_group("gloo")
comm = kb.Comm()
rank, size = comm.rank, comm.size
assert size == 2
assert comm.allreduce(float(rank + 1), op=kb.MPI.SUM) == 3.0 and comm.allreduce(float(rank), op=kb.MPI.MAX) == 1.0
assert comm.bcast("id" if rank == 0 else None, root=0) == "id"
mesh = kb.mesh.cell_array_mesh(2, 24, 3)
local, info = part.partition_mesh(mesh, rank, size)
lay = part.Layout(part.local_dofmaps(local), local.n_owned)
allreq = part.gather_requests(comm, local, info, lay)
peers, sp, sc, rp, rc = part.build_halo_lists(local, info, lay, allreq)
x = np.full(lay.n_cols, np.nan)
for s in range(2):
    gv = info["l2g"][lay.node_vert[s]]
    for f in range(4):
        x[lay.col(s, f, np.arange(lay.n_own[s]))] = 1000.0 * (4 * s + f) + gv[:lay.n_own[s]]
reqs = []
recv_bufs = []
for i, pr in enumerate(peers.tolist()):
    sb = torch.from_numpy(x[sc[sp[i]:sp[i + 1]]].copy())
    rb = torch.empty(int(rp[i + 1] - rp[i]), dtype=torch.float64)
    reqs += [dist.isend(sb, pr), dist.irecv(rb, pr)]
    recv_bufs.append(rb)
for r in reqs:
    r.wait()
for i, rb in enumerate(recv_bufs):
    x[rc[rp[i]:rp[i + 1]]] = rb.numpy()
assert not np.isnan(x).any()
for s in range(2):
    gv = info["l2g"][lay.node_vert[s]]
    for f in range(4):
        assert np.array_equal(x[lay.col(s, f, np.arange(lay.n_loc[s]))], 1000.0 * (4 * s + f) + gv)
print("WORKER_OK", rank, flush=True)
dist.destroy_process_group()
