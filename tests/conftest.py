import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def g_ops():
    return load_golden("ops.npz")


@pytest.fixture(scope="session")
def g_coords():
    return load_golden("coords.npz")


@pytest.fixture(scope="session")
def g_gen():
    return load_golden("g_small.npz")


@pytest.fixture(scope="session")
def g_disc():
    return load_golden("d_small.npz")


@pytest.fixture(scope="session")
def g_ada():
    return load_golden("ada.npz")


@pytest.fixture(scope="session")
def g_vanilla():
    return load_golden("vanilla_small.npz")


@pytest.fixture(scope="session")
def g_inv():
    return load_golden("inversion.npz")


@pytest.fixture(scope="session")
def g_invloop():
    return load_golden("inversion_loop.npz")


@pytest.fixture(scope="session")
def g_step():
    return load_golden("trainer_step.npz")


@pytest.fixture(scope="session")
def g_step_mid():
    return load_golden("trainer_step_mid.npz")


@pytest.fixture(scope="session")
def g_kitti():
    return load_golden("kitti_scan.npz")
