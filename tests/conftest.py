import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def readme_Y():
    from oracle.r_rng import readme_demo_Y
    return readme_demo_Y()


@pytest.fixture(scope="session")
def readme_oracle_fit(readme_Y):
    from oracle.scgbm_oracle import gbm_sc
    return gbm_sc(readme_Y, 10)
