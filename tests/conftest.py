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
