"""The C-ABI shared library loads on a box without a GPU, exports every symbol the header declares,
and refuses to run (loudly) when no CUDA device is visible.  No compute calls here."""
import ctypes
import os
import re

import pytest

from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.engine import ABI_SYMBOLS, CConfig, load_library
from fastenhancer_b200.schema import canonical_size

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "fastenhancer_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fe_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = load_library()
    for sym in _declared():
        assert hasattr(lib, sym), sym


@pytest.mark.parametrize("name", sorted(PRESETS))
def test_sizes_match_host_schema(name):
    lib, cfg = load_library(), PRESETS[name]
    c = CConfig.from_cfg(cfg)
    assert lib.fe_weight_count(ctypes.byref(c)) == canonical_size(cfg)
    assert lib.fe_state_floats(ctypes.byref(c)) == cfg.state_floats


def test_no_device_is_a_loud_error():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    lib, cfg = load_library(), PRESETS["16k_t"]
    c = CConfig.from_cfg(cfg)
    w = np.zeros(canonical_size(cfg), np.float32)
    h = ctypes.c_void_p()
    rc = lib.fe_create(ctypes.byref(c), w.ctypes.data, w.size, 0, ctypes.byref(h))
    assert rc == -4 and not h.value                      # FE_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.fe_last_error()
    from fastenhancer_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(cfg, w)


def test_unsupported_shape_is_rejected():
    import dataclasses
    cfg = dataclasses.replace(PRESETS["16k_b"], activation="ReLU")
    with pytest.raises(ValueError):
        cfg.validate()
