import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "hotpath_golden.npz"))


@pytest.fixture(scope="session")
def golden_cases(golden):
    return sorted({k.split("/")[0] for k in golden.files if "/" in k})


def pytest_sessionfinish(session, exitstatus):
    """Worst raw ratios against the bare north-star tolerances (tests/parity.py REPORT), printed and saved."""
    try:
        from tests import parity
    except Exception:
        return
    if not parity.REPORT:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(parity.REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass
    worst = {}
    for slot in parity.REPORT.values():
        for k, v in slot.items():
            worst[k] = max(worst.get(k, 0.0), v)
    print("\n[parity] worst raw ratios vs bare north-star tolerances:", json.dumps(worst))
