import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def refdata():
    with open(os.path.join(ROOT, "carnd-mpc-project_b200", "data", "reference_data.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def stable_cd(refdata, po):
    return po.load_config_dict(refdata["configs"]["stable"])


@pytest.fixture(scope="session")
def mpc():
    """The product binding.  Builds the CUDA library if it is missing (nvcc cross-compiles)."""
    import mpc_b200
    if not os.path.exists(mpc_b200.LIB_PATH):
        mpc_b200.build()
    return mpc_b200


@pytest.fixture(scope="session")
def stable_cfg(mpc, refdata):
    return mpc.config_from_json_text(json.dumps(refdata["configs"]["stable"]))


@pytest.fixture(scope="session", params=[2, 3], ids=["lane-kernel", "coop-kernel"])
def kernel_kind(request):
    """Every CUDA kernel is held to the same parity bar: 2 = one problem per lane (throughput path; problems that need
    a rare branch of the algorithm finish in the coop kernel), 3 = one problem per lane group with the rows in shared
    memory (latency / small-batch / tail path, every branch of the algorithm)."""
    return request.param


@pytest.fixture(scope="session")
def solver(mpc, stable_cfg, kernel_kind):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    s = mpc.Solver(stable_cfg, 0)
    s.set_kernel(kernel_kind)
    yield s
    s.close()
