import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_scipy.npz"))


def pytest_sessionfinish(session, exitstatus):
    """Observed max error / tolerance per parity test (tests/helpers.py: OBSERVED), next to the GPU run's other output."""
    import json
    try:
        from tests import helpers
    except Exception:
        return
    if not helpers.OBSERVED:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    rows = {k: {"err_over_tol": round(v[0], 4), "rel_err": float(f"{v[1]:.3e}")} for k, v in sorted(helpers.OBSERVED.items())}
    with open(os.path.join(out_dir, "parity_observed.json"), "w") as f:
        json.dump(rows, f, indent=0)
