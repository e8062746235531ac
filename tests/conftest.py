import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly if selected without a device; they are only skipped when the
    # whole run did not ask for them (`-m "not gpu"` deselects them before this hook).
    pass


ALAN = dict(time_step=1 / 60., neighbor_dist=5.0, max_neighbors=10, time_horizon=1.5, time_horizon_obst=1.5,
            radius=0.5, max_speed=1.0)
ENV = dict(time_step=1 / 60., neighbor_dist=1.5, max_neighbors=5, time_horizon=1.5, time_horizon_obst=1.5,
           radius=0.5, max_speed=1.0)


@pytest.fixture(scope="session")
def alan_params():
    return dict(ALAN)


@pytest.fixture(scope="session")
def env_params():
    return dict(ENV)
