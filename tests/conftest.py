import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pp_golden():
    with open(os.path.join(GOLDEN, "postprocessing_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def mpc_golden():
    with open(os.path.join(GOLDEN, "mpc_oracle_golden.json")) as f:
        return json.load(f)


def infra_from_json(d):
    """JSON round trip turns arrays into lists; restore what the shim expects."""
    out = dict(d)
    for k in ("constraint_matrix", "constraint_limits", "phases", "voltages", "max_pilot", "min_pilot"):
        out[k] = np.array(d[k], dtype=float)
    out["allowable_pilots"] = [np.array(a, dtype=float) for a in d["allowable_pilots"]]
    return out


@pytest.fixture(scope="session")
def require_gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from adacharge_b200 import _cabi

    _cabi.lib()  # raises loudly if the native library is missing
