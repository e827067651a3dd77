"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repository root:
``python tests/golden/make_golden.py``).

These are ORACLE-derived regression pins (the reference itself cannot be run here: DOLFINx/PETSc are absent).
The only values that come from the reference are the four golden norms in tests/conftest.py."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.fixtures import unit_square, unit_cube          # noqa: E402
from oracle.knpemi import KNPEMIOracle, OracleParams        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = [("NeuronalCT", None), ("HH", None), ("ATP", None)]


def c1():
    o = KNPEMIOracle(unit_square(32), OracleParams(), MODELS)
    o.t += o.p.dt
    o.gate_update()
    A, b = o.assemble(o.t)
    P = o.assemble_P()
    out = dict(indptr=A.indptr.astype(np.int32), indices=A.indices.astype(np.int32), b=b,
               A_rowsum=np.asarray(A.sum(axis=1)).ravel(), A_absrowsum=np.asarray(abs(A).sum(axis=1)).ravel(),
               A_diag=A.diagonal(), P_diag=P.diagonal(), gates=o.gates[:, o.mverts], S_i=o.S[0], S_e=o.S[1])
    # 10 steps, direct convention: per-step functionals
    o = KNPEMIOracle(unit_square(32), OracleParams(), MODELS)
    phim_mean, norms = [], []
    x = o.pack()
    for i in range(10):
        o.step("direct", first=(i == 0))
        phim_mean.append(o.phi_m[o.mverts].mean())
        norms.append([o.l2_norm(o.phi[0], 1), o.l2_norm(o.phi[1], 2)] +
                     [o.l2_norm(o.c[s][k], 1 if s == 0 else 2) for s in range(2) for k in range(3)])
    out.update(phim_mean=np.array(phim_mean), norms=np.array(norms), gates_final=o.gates[:, o.mverts])
    np.savez_compressed(os.path.join(HERE, "c1_square32.npz"), **out)


def cube():
    o = KNPEMIOracle(unit_cube(6), OracleParams(), MODELS)
    o.t += o.p.dt
    o.gate_update()
    A, b = o.assemble(o.t)
    np.savez_compressed(os.path.join(HERE, "cube6.npz"), indptr=A.indptr.astype(np.int32),
                        indices=A.indices.astype(np.int32), b=b, A_data=A.data)


if __name__ == "__main__":
    c1()
    cube()
    print("wrote", os.listdir(HERE))
