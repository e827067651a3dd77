"""Regenerates tests/golden/parity_small.json from the CPU oracle (``python tests/golden/make_parity_golden.py``).

Two small fixed problems -- BASELINE C3 in miniature (2D tissue block, HH + ATP + KCC2, perturbed initial state) and C4 in
miniature (3D tissue block, passive membrane) -- are stepped three times with linear solves driven to 1e-13, and the L2
norms of the eight fields are stored.  bench.py replays the same problems on however many GPUs it was launched on and
compares (rank-count independence of the distributed path, checked in every driver-run bench record) WITHOUT importing
the oracle in the product arm.  ORACLE-derived pins, not reference outputs."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cgx_b200 as kb                                         # noqa: E402  (mesh generators only; no GPU needed)
from oracle.fixtures import from_arrays                      # noqa: E402
from oracle.knpemi import KNPEMIOracle, OracleParams         # noqa: E402
from oracle.amg import SchurPC                               # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STEPS = 3

CASES = {
    # name: (gdim, N, cells_per_dim, models, perturbed, stimulus tags)
    "c3_mini": (2, 32, 2, [("NeuronalCT", None), ("HH", None), ("ATP", None)], True, (2,)),
    "c4_mini": (3, 8, 2, [("Passive", None)], False, ()),
}


def run(name):
    gdim, n, m, models, perturbed, stim = CASES[name]
    mesh = kb.mesh.cell_array_mesh(gdim, n, m)
    om = from_arrays(gdim, mesh.x, mesh.cells, mesh.cell_tags, mesh.intra_tags)
    it = tuple(mesh.intra_tags)
    p = OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=stim if stim else it)
    o = KNPEMIOracle(om, p, models)
    if perturbed:                                             # configs/c3_*.yaml initial_perturbation
        X = om.x / 1e-6
        fac = 1 + 0.01 * np.sin(2 * np.pi * X[:, 0]) * np.sin(2 * np.pi * X[:, 1])
        for s in range(2):
            o.c[s] *= fac[None, :]
        dphi = 0.005 * np.cos(2 * np.pi * X[:, 0])
        o.phi_m += dphi
        o.phi[0] += dphi
    pc = SchurPC(o, exact=True)
    x = o.pack()
    for i in range(STEPS):
        _, _, x, _ = o.step("gmres", pc, 1e-13, x, first=(i == 0))
    norms = []
    for s in range(2):
        tags = list(it) if s == 0 else [1]
        for f in range(4):
            norms.append(float(o.l2_norm(o.c[s][f] if f < 3 else o.phi[s], tags)))
    return dict(gdim=gdim, N=n, cells_per_dim=m, steps=STEPS, rows=int(o.n), norms=norms)


if __name__ == "__main__":
    out = {k: run(k) for k in CASES}
    with open(os.path.join(HERE, "parity_small.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))
