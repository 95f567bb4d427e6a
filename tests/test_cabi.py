"""CPU tests of the C-ABI boundary: the library builds, loads, exports every symbol the header
declares, and refuses to compute without a GPU (no silent fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from gigalens_b200 import _cabi, workloads
from gigalens_b200.simulator import CompiledModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    __graft_entry__.build()
    return _cabi.load()


def test_header_and_binding_agree(lib):
    header = open(os.path.join(ROOT, "include", "gigalens_b200.h")).read()
    declared = set(re.findall(r"\b(gl_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTED_SYMBOLS), declared ^ set(_cabi.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.gl_abi_version() == 2


def test_struct_layout_matches_header():
    # field order/sizes of the ctypes mirrors (a mismatch would silently corrupt plans)
    assert C.sizeof(_cabi.ProfileDesc) == 4 + 4 + 40 + 40 + 4 + 4 + 8 + 4 + 4 + 8   # GL_MAX_PROFILE_PARAMS = 10 (ABI version 2)
    assert C.sizeof(_cabi.PriorLeaf) == 24
    assert C.sizeof(_cabi.LikeConfig) == 24


def test_model_compiles_to_descriptor():
    cm = CompiledModel(workloads.demo_phys_model())
    assert cm.n_params == 22 and cm.desc.n_lens == 2 and cm.desc.n_lens_light == 1 and cm.desc.n_source_light == 1
    assert cm.slot_keys[0] == ("lens_mass", 0, "theta_E") and cm.slot_keys[-1] == ("source_light", 0, "Ie")
    assert cm.desc.lens[0].type == _cabi.GL_EPL and cm.desc.lens[0].niter == 50


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    cm = CompiledModel(workloads.demo_phys_model())
    sc = _cabi.SimConfig()
    g = np.zeros(16, np.float32)
    sc.num_pix, sc.supersample = 4, 1
    sc.grid_x = sc.grid_y = g.ctypes.data_as(C.POINTER(C.c_float))
    plan = C.c_void_p()
    assert lib.gl_plan_create(C.byref(cm.desc), C.byref(sc), 2, 0, C.byref(plan)) != 0
    assert b"no CUDA device" in lib.gl_last_error()
    from gigalens_b200.simulator import LensSimulator

    with pytest.raises(RuntimeError):
        LensSimulator(workloads.demo_phys_model(), workloads.demo_sim_config(), bs=2)


def test_bad_descriptors_are_rejected(lib):
    plan = C.c_void_p()
    assert lib.gl_plan_create(None, None, 1, 0, C.byref(plan)) != 0
    assert lib.gl_last_error()
