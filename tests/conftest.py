import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, 'tests'), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The C oracle (oracle/nmpc_oracle.c), built on demand — test infrastructure only."""
    from oracle import oracle_c
    oracle_c.build()
    return oracle_c


@pytest.fixture(scope="session")
def gpu_solver_factory():
    import mpc_trajectory_generator_b200 as pkg
    made = []

    def make(cfg=None, device=0):
        s = pkg.NmpcSolver(cfg, device)
        made.append(s)
        return s

    yield make
    for s in made:
        s.close()
