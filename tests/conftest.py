import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# GPU tests run the tcgen05 kernels with BOUNDED pipeline waits (a protocol bug traps with a site code instead of hanging the
# box); production launches spin without a bound (tc_common.cuh).  Must be set before the library reads it (first launch).
os.environ.setdefault("FNSSL_TC_WAIT_TIMEOUT", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_fnssl():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "fnssl_golden.npz"))


@pytest.fixture(scope="session")
def golden_ipdnet():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ipdnet_golden.npz"))


@pytest.fixture(scope="session")
def golden_ipdnet2():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "ipdnet2_golden.npz"))
