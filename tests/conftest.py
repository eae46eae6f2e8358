import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """The parity suites leave the errors they measured in gpurun_out/ (copied to profiles/ by hand)."""
    try:
        from tests.test_engine_emulated import MEASURED
    except Exception:
        return
    if not MEASURED:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    rows = [dict(method=k[0], dataset=k[1], device=k[2], precision=k[3], steps=k[4], **v) for k, v in MEASURED.items()]
    with open(os.path.join(out, "parity_fixture_errors.json"), "w") as f:
        json.dump(rows, f, indent=1)
