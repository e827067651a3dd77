"""Stand-in for `from mpi4py import MPI` in the reference's driver scripts: the reduction ops and a COMM_WORLD with the
mpi4py surface those scripts use, backed by torch.distributed (see knp-emi-cgx_b200/comm.py)."""
import importlib as _il

_comm = _il.import_module("knp-emi-cgx_b200.comm")


class _MPI:
    SUM, MAX, MIN = _comm.MPI.SUM, _comm.MPI.MAX, _comm.MPI.MIN
    COMM_WORLD = _comm.Comm()
    Comm = _comm.Comm


MPI = _MPI
