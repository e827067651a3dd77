"""CPU tier: the C++/OpenMP CPU baseline (oracle/cpu_baseline/knpemi_cpu.cpp, what bench.py times as `cpu_baseline` and in the
`--impl reference` arm) against the numpy oracle it restates: assembled matrix / vector entries to 1e-12, the Schur
preconditioner application, GMRES iteration counts and per-step solutions of the time loop."""
import numpy as np
import pytest

from oracle.cpu import CpuBaseline
from oracle.fixtures import from_arrays, unit_square
from oracle.knpemi import KNPEMIOracle, OracleParams
from oracle.amg import SchurPC
from conftest import MODELS_TEST


def _case(kb, name):
    if name == "square32":
        om, p, models = unit_square(32), OracleParams(), MODELS_TEST
    else:
        gdim, n, m = (2, 24, 3) if name == "cells2d" else (3, 8, 2)
        mesh = kb.mesh.cell_array_mesh(gdim, n, m)
        om = from_arrays(gdim, mesh.x, mesh.cells, mesh.cell_tags, mesh.intra_tags)
        it = tuple(mesh.intra_tags)
        if name == "cells2d":
            p = OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=(2,), stimulus_region=(0, 0.1e-6, 0.3e-6))
            models = [("NeuronalCT", None), ("HH", None), ("ATP", None)]
        elif name == "glia3d":
            p = OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=it, glia_tags=it[4:])
            models = [("HH", it[:4]), ("ATP", it[:4]), ("KirNa", it[4:]), ("GlialCT", it[4:])]
        else:
            p = OracleParams(intra_tags=it, extra_tag=1, membrane_tags=it, stimulus_tags=it)
            models = [("Passive", None)]
    o = KNPEMIOracle(om, p, models)
    rng = np.random.default_rng(1)
    for s in range(2):
        o.c[s] *= 1 + 0.03 * rng.random(o.c[s].shape)
    o.phi[0] += 0.003 * rng.standard_normal(o.phi[0].shape)
    o.phi_m = o.phi[0] - o.phi[1]
    return om, p, models, o


def _baseline(om, p, models, o):
    A, _ = o.assemble(o.t + p.dt)
    # the library builds the CSR pattern itself; passing the oracle's makes construction fail on any difference
    pat = (A.indptr.astype(np.int32), A.indices.astype(np.int32))
    cb = CpuBaseline(om.gdim, om.x, om.cells, om.cell_tags, om.mf_verts, om.mf_tags, p, models, pat)
    assert cb.nnz == A.nnz and np.array_equal(cb.S[0], o.S[0]) and np.array_equal(cb.S[1], o.S[1])
    cb.set_state(o.pack(), o.gates[:, o.mverts])
    assert np.array_equal(cb.mverts, o.mverts)
    return cb


@pytest.mark.parametrize("name", ["square32", "cells2d", "passive3d", "glia3d"])
def test_assembly_matches_numpy_oracle(kb, name):
    om, p, models, o = _case(kb, name)
    cb = _baseline(om, p, models, o)
    t = 3 * p.dt
    A, b = o.assemble(t)
    vals, bb = cb.assemble(t)
    scale = np.repeat(np.maximum.reduceat(np.abs(A.data), A.indptr[:-1]), np.diff(A.indptr))
    assert (np.abs(vals - A.data) / scale).max() < 1e-12
    assert np.abs(bb - b).max() <= 1e-12 * np.abs(b).max()
    cb.close()


@pytest.mark.parametrize("name", ["square32", "cells2d", "passive3d"])
def test_time_loop_matches_numpy_oracle(kb, name):
    """Same algorithm, same iteration counts: GMRES(30) + Schur preconditioner (SA-AMG W-cycles), three steps."""
    om, p, models, o = _case(kb, name)
    cb = _baseline(om, p, models, o)
    pc = SchurPC(o)
    cb.pc_setup()
    r = np.random.default_rng(2).standard_normal(o.n)
    z_ref, z = pc(r), cb.pc_apply(r)
    assert np.abs(z - z_ref).max() <= 1e-9 * np.abs(z_ref).max()
    x = o.pack()
    for i in range(3):
        _, _, x, its_ref = o.step("gmres", pc, 1e-9, x, first=(i == 0))
        its, ms = cb.step(1e-9)
        u, g = cb.get_state()
        assert abs(its - its_ref) <= 1, (i, its, its_ref)
        for s in range(2):
            for f in range(4):
                sl = slice(o.base[s] + f * o.ns[s], o.base[s] + (f + 1) * o.ns[s])
                ref = x[sl]
                scale = np.abs(ref).max() if f < 3 else max(np.abs(x[o.base[0] + 3 * o.ns[0]: o.base[0] + 4 * o.ns[0]]).max(), 1e-3)
                assert np.abs(u[sl] - ref).max() <= 2e-8 * scale, (i, s, f)
        assert np.abs(g - o.gates[:, o.mverts]).max() < 1e-9
    assert cb.threads >= 1
    cb.close()
