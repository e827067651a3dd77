"""Communicator shim with the mpi4py surface the reference drivers use (comm.rank, comm.size,
comm.allreduce(x, op=MPI.SUM|MAX|MIN), comm.bcast, comm.Barrier), backed by torch.distributed
(NCCL on GPUs, gloo on CPU) when a process group is initialised, trivial otherwise.
Replaces MPI.COMM_WORLD of src/CGx/utils/mixed_dim_problem.py:27."""
import numpy as np


class _Op:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"MPI.{self.name}"


class MPI:
    SUM = _Op("SUM")
    MAX = _Op("MAX")
    MIN = _Op("MIN")


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


class Comm:
    @property
    def rank(self):
        d = _dist()
        return d.get_rank() if d else 0

    @property
    def size(self):
        d = _dist()
        return d.get_world_size() if d else 1

    def allreduce(self, value, op=MPI.SUM):
        d = _dist()
        if d is None or d.get_world_size() == 1:
            return value
        import torch
        dev = "cuda" if d.get_backend() == "nccl" else "cpu"
        scalar = np.isscalar(value)
        t = torch.as_tensor(np.atleast_1d(np.asarray(value, dtype=np.float64)), device=dev).clone()
        ops = {"SUM": d.ReduceOp.SUM, "MAX": d.ReduceOp.MAX, "MIN": d.ReduceOp.MIN}
        d.all_reduce(t, op=ops[op.name])
        out = t.cpu().numpy()
        return float(out[0]) if scalar else out

    def bcast(self, obj, root=0):
        d = _dist()
        if d is None or d.get_world_size() == 1:
            return obj
        box = [obj]
        d.broadcast_object_list(box, src=root)
        return box[0]

    def allgather(self, obj):
        d = _dist()
        if d is None or d.get_world_size() == 1:
            return [obj]
        out = [None] * d.get_world_size()
        d.all_gather_object(out, obj)
        return out

    def Barrier(self):
        d = _dist()
        if d is not None and d.get_world_size() > 1:
            d.barrier()
