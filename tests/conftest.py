import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "v2x-sim_b200")
for p in (ROOT, SRC):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
LIB = os.path.join(SRC, "v2x_b200", "libv2x_b200.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are SKIPPED (not failed) on a host without a CUDA device or without the built library, so a
    plain ``pytest tests`` is clean on a CPU box; on a GPU box a missing library still fails loudly (test_cabi)."""
    import torch
    reason = None
    if not torch.cuda.is_available():
        reason = "no CUDA device"
    elif not os.path.exists(LIB):
        reason = "libv2x_b200.so not built"
    if reason is None:
        return
    skip = pytest.mark.skip(reason="gpu test: " + reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


# ---- measured parity errors, on record -----------------------------------------------------------------------------
# GPU parity tests call ``parity_log(case, mode, **numbers)``; at session end everything measured is written to
# gpurun_out/parity.json (merged back from the GPU box by gpurun) -- the committed copy is profiles/r02_parity.json.
_PARITY = []


@pytest.fixture(scope="session")
def parity_log():
    def log(case, mode, **numbers):
        _PARITY.append(dict(case=case, mode=mode, **{k: (float(v) if isinstance(v, float) else v) for k, v in numbers.items()}))
    return log


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, os.environ.get("V2X_PARITY_FILE", "parity.json"))   # child runs write their own file
    prev = []
    if os.path.exists(path) and os.environ.get("V2X_PARITY_APPEND"):
        prev = json.load(open(path))["records"]
    with open(path, "w") as f:
        json.dump({"metric": "max-abs error / max-abs value vs the fp32 CPU oracle (restatement of the reference); "
                             "golden = same vs the live-reference fixture subsample", "exit": int(exitstatus),
                   "records": prev + _PARITY}, f, indent=1)
