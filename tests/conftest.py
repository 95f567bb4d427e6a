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


GUARD_LOG = []


@pytest.fixture(autouse=True)
def _red_zone_check(request):
    """With GL_GUARD=1 (the memcheck stand-in: compute-sanitizer is closed on the GPU pool) every device allocation of the
    library carries 64 KB red zones; after each GPU test they must be intact.  Results go to gpurun_out/guard_check.json."""
    yield
    if os.environ.get("GL_GUARD", "0") in ("", "0") or "gpu" not in request.keywords:
        return
    from gigalens_b200 import _cabi

    lib = _cabi.load()
    n = lib.gl_guard_check()
    GUARD_LOG.append(dict(test=request.node.nodeid, live_guarded_allocations=int(n),
                          error=lib.gl_last_error().decode() if n < 0 else ""))
    assert n >= 0, lib.gl_last_error().decode()


def pytest_sessionfinish(session, exitstatus):
    """Write the parity report (how many slices needed the fp32 noise-floor rule, and by how much) next to the
    other GPU-run artefacts; `gpurun_out/` is merged back from the GPU box and copied to profiles/r02_parity.json."""
    if GUARD_LOG:
        import json

        out_dir = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "guard_check.json"), "w") as f:
            json.dump(dict(what="GL_GUARD=1: 64 KB red zones (0xFF) around every device allocation of libgigalens_b200.so, verified after "
                                "each GPU test; out-of-bounds reads of those zones would surface as NaN / -1 in the parity checks",
                           tests=len(GUARD_LOG), corrupted=sum(1 for r in GUARD_LOG if r["live_guarded_allocations"] < 0),
                           max_live_allocations=max(r["live_guarded_allocations"] for r in GUARD_LOG), records=GUARD_LOG), f, indent=1)
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
