import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def kb():
    import cgx_b200
    return cgx_b200


@pytest.fixture(scope="session")
def cfgdir(kb):
    return os.path.join(os.path.dirname(kb.__file__), "configs")


MODELS_TEST = [("NeuronalCT", None), ("HH", None), ("ATP", None)]

# golden L2 norms held by the reference's own tests
GOLD_DIRECT = (2.6337161145147203e-08, 1.5258564901943312e-08)      # tests/KNPEMI/electric_potential_norms_direct_solver.py:55-56
GOLD_ITERATIVE = (3.510994056704844e-08, 6.369472309249516e-11)      # tests/KNPEMI/electric_potential_norms_iterative_solver.py:58-59
GOLD_ITERATIONS = 3.0                                                # ...iterative_solver.py:81


FLAGS = {"NeuronalCT": 8, "HH": 32, "ATP": 16, "Passive": 1, "GlialCT": 4, "KirNa": 2}


def params_struct(kb, p, models, stim_area=0.0):
    """knp_params + the (tag, model flags, stimulated) table for an oracle parameter set (direct C-ABI use)."""
    P = kb.lib.Params()
    P.dt, P.F, P.R, P.T, P.C_M, P.phi_rest = p.dt, p.F, p.R, p.T, p.C_M, p.phi_rest
    for k in range(3):
        P.z[k], P.D[k], P.g_leak[k], P.g_leak_g[k] = p.z[k], p.D[k], p.g_leak[k], p.g_leak_g[k]
    P.g_Na_bar, P.g_K_bar, P.g_syn_bar, P.a_syn, P.T_stim = p.g_Na_bar, p.g_K_bar, p.g_syn_bar, p.a_syn, p.T_stim
    P.scale_stimulus = int(p.scale_stimulus)
    for i in range(3):
        P.stim_dir[i] = -1
    if p.stimulus_region is not None:
        regions = p.stimulus_region if isinstance(p.stimulus_region[0], (tuple, list)) else [p.stimulus_region]
        for i, (d, lo, hi) in enumerate(regions):
            P.stim_dir[i], P.stim_lo[i], P.stim_hi[i] = d, lo, hi
    P.K_e_init, P.K_i_g_init = p.c_e_init[1], p.c_i_g_init[1]
    P.ode_substeps, P.rush_larsen, P.stim_area = p.ode_substeps, int(p.rush_larsen), stim_area
    table = {}
    for name, tags in models:
        for t in (p.membrane_tags if tags is None else tags):
            table[t] = table.get(t, 0) | FLAGS[name]
    return P, [(t, fl, t in p.stimulus_tags) for t, fl in sorted(table.items())]


def perturb(o, seed=0):
    """Random perturbation of every field of an oracle instance (the same for P1 and P2 instances)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    for s in range(2):
        o.c[s] *= 1 + 0.05 * rng.random(o.c[s].shape)
    o.phi[0] += 0.004 * rng.standard_normal(o.phi[0].shape)
    o.phi[1] += 0.001 * rng.standard_normal(o.phi[1].shape)
    o.phi_m = o.phi[0] - o.phi[1]
    o.gates *= 1 + 0.1 * rng.random(o.gates.shape)
    return o
