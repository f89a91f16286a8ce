import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built_lib():
    from cafe_mpc_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def cm(built_lib):
    import cafe_mpc_b200
    return cafe_mpc_b200


@pytest.fixture(scope="session")
def data_dir():
    return os.path.join(REPO, "data")


@pytest.fixture(scope="session")
def hkd_problem(cm, data_dir):
    return cm.HKDProblem(os.path.join(data_dir, "Reference/Data/trot/heuristic/quad_reference.csv"))


@pytest.fixture(scope="session")
def hkd_options(cm, data_dir):
    return cm.load_hsddp_setting(os.path.join(data_dir, "HKDMPC/settings/ddp_setting.info"))
