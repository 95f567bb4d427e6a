import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


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


def pytest_sessionfinish(session, exitstatus):
    """Write the parity report (how many slices needed the fp32 noise-floor rule, and by how much) next to the
    other GPU-run artefacts; `gpurun_out/` is merged back from the GPU box and copied to profiles/r02_parity.json."""
    try:
        import json

        import common

        if not common.PARITY_LOG:
            return
        import torch

        out_dir = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        name = "parity_gpu.json" if torch.cuda.is_available() else "parity_cpu.json"
        rep = common.parity_summary()
        rep["records"] = common.PARITY_LOG
        with open(os.path.join(out_dir, name), "w") as f:
            json.dump(rep, f, indent=1)
    except Exception as e:   # a report must never fail the suite
        print("[parity report] not written:", e)
