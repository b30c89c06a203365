import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch
    return {n: torch.load(os.path.join(GOLDEN, n + ".pt"), weights_only=False)
            for n in ("swin", "rs_gcn", "roberta", "graph", "swin_train", "roberta_train", "fusion_classes")}


# measured parity errors of the -m gpu run: every model-level assert records (name, measured, tolerance) here and the
# session writes gpurun_out/parity_errors.json (copied to profiles/ as the round's evidence)
PARITY = {}


def record_parity(name: str, measured: float, tol: float):
    PARITY[name] = {"measured": float(measured), "tolerance": float(tol)}
    print(f"[parity] {name}: measured {measured:.3e} (tolerance {tol:.0e})")
    return measured


def pytest_sessionfinish(session, exitstatus):
    if not PARITY:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_errors.json"), "w") as fh:
        json.dump(PARITY, fh, indent=1, sort_keys=True)
