import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from __graft_entry__ import load_package  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = ["NGC6440E", "sim_sw_wb", "sim3_gp", "sim_dmgp_wb", "sim2"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def vb():
    return load_package()


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.lib()
    return O


class Golden:
    def __init__(self, vb, name):
        path = os.path.join(GOLDEN, name + ".npz")
        self.name = name
        self.model, self.toas, self.extra = vb.readwrite.load_pulsar_npz(path)
        z = np.load(path)
        self.y, self.w = z["golden_y"], z["golden_w"]
        self.walkers, self.lnlike = z["golden_walkers"], z["golden_lnlike"]
        self.phiinv = z["golden_phiinv"] if "golden_phiinv" in z else None
        self.md = vb.ModelDescriptor(self.model, self.toas)
        self.x0 = self.model.read_param_values_to_vector()


@pytest.fixture(scope="session")
def golden(vb):
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(vb, name)
        return cache[name]

    return get


def has_gpu() -> bool:
    try:
        return load_package().lib.lib().vela_b200_device_count() > 0
    except Exception:
        return False
