import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-gaussian-splatting-for-novel-view-synthesis_b200")
for p in (ROOT, PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


# ---- measured parity figures (max abs image error, flip pixels, per-tensor gradient errors, integer-stage counts) ----
# The -m gpu parity tests record what they MEASURE, not only whether it passed; at session end the numbers are written
# to $B200GS_PARITY_OUT (default gpurun_out/PARITY_r02.json when that directory exists) and a copy is committed under
# profiles/.
_PARITY_CASES = {}


@pytest.fixture(scope="session")
def parity_log():
    return _PARITY_CASES


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY_CASES:
        return
    import json
    out = os.environ.get("B200GS_PARITY_OUT")
    if out is None:
        d = os.path.join(ROOT, "gpurun_out")
        if not os.path.isdir(d):
            return
        out = os.path.join(d, "PARITY_r02.json")
    try:
        import torch
        gpu = torch.cuda.get_device_name(0) if torch.cuda.is_available() else None
    except Exception:  # noqa: BLE001
        gpu = None
    doc = {"tolerances": {"image_abs": 1e-4, "grad_rel_maxnorm": 1e-3, "integer_stages": "exact"},
           "reference": "unmodified reference via the oracle restatement (bit-identical, oracle/make_golden.py) and the "
                        "golden vectors generated from it", "gpu": gpu, "exit_status": int(exitstatus),
           "cases": _PARITY_CASES}
    try:
        with open(out, "w") as f:
            json.dump(doc, f, indent=1, sort_keys=True, default=float)
    except OSError:
        pass
