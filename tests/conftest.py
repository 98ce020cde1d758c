import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def water():
    from jqmc_b200.trexio_lite import load_golden_system

    return load_golden_system(os.path.join(GOLDEN, "water_ccecp_ccpvqz.npz"))


def load_system(name):
    from jqmc_b200.trexio_lite import load_golden_system

    return load_golden_system(os.path.join(GOLDEN, name + ".npz"))


def random_walkers(H, nw, seed, scale=0.8):
    import numpy as np

    rng = np.random.default_rng(seed)
    R = np.asarray(H.structure_data.positions, dtype=np.float64)
    gem = H.wavefunction_data.geminal_data
    n_up, n_dn = gem.num_electron_up, gem.num_electron_dn
    own_u = rng.integers(0, len(R), size=(nw, n_up))
    own_d = rng.integers(0, len(R), size=(nw, n_dn))
    r_up = R[own_u] + rng.normal(scale=scale, size=(nw, n_up, 3))
    r_dn = R[own_d] + rng.normal(scale=scale, size=(nw, n_dn, 3))
    return r_up, r_dn
