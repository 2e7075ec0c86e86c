"""ctypes front-end of the CPU emulation build of the fused kernel (tests/emu/fe_emu.cpp).
TEST INFRASTRUCTURE ONLY: validates packer / tile / layout logic without a GPU."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "fe_emu.cpp")
_SO = os.path.join(_HERE, "_build", "libfe_emu.so")
_CSRC = os.path.join(_HERE, "..", "..", "fastenhancer_b200", "csrc")
_lib = None
MODE_STREAM, MODE_SPEC, MODE_OFFLINE = 0, 1, 2


def build(force=False):
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("fe_plan.h", "fe_pack.h", "fe_kernel.cuh", "fe_configs.h", "fe_half.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-o", _SO, _SRC], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.fee_run.argtypes = [ctypes.c_int] * 10 + [fp, ctypes.c_int, fp, fp, fp, fp] + [ctypes.c_int] * 3 + \
            [ctypes.c_longlong] * 2 + [fp, ctypes.c_int, ctypes.c_float]
        _lib.fee_tap_total.argtypes = [ctypes.c_int] * 8
    return _lib


def _p(a):
    if a is None:
        return ctypes.cast(None, ctypes.POINTER(ctypes.c_float))
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _shape(cfg):
    return (cfg.n_fft, cfg.hop_size, cfg.channels, cfg.n_enc, cfg.rf_channels, cfg.rf_freq, cfg.rf_blocks, cfg.rf_heads)


def to_native(cfg, state):
    """canonical state [B, 2*CL + K*F2*C2] (h as [K][F2][C2]) -> native (h as [K][C2][F2])."""
    B = state.shape[0]
    cl2 = 2 * cfg.cache_len
    h = state[:, cl2:].reshape(B, cfg.rf_blocks, cfg.rf_freq, cfg.rf_channels).transpose(0, 1, 3, 2)
    return np.ascontiguousarray(np.concatenate([state[:, :cl2], h.reshape(B, -1)], axis=1), np.float32)


def to_canonical(cfg, state):
    B = state.shape[0]
    cl2 = 2 * cfg.cache_len
    h = state[:, cl2:].reshape(B, cfg.rf_blocks, cfg.rf_channels, cfg.rf_freq).transpose(0, 1, 3, 2)
    return np.ascontiguousarray(np.concatenate([state[:, :cl2], h.reshape(B, -1)], axis=1), np.float32)


def run(cfg, S, canonical, mode, state_native, inp, out, spec_out=None, n_streams=1, n_hops=1, L=0, ld_in=0, ld_out=0,
        dbg=None, dbg_hop=-1, tc=False):
    lib = _load()
    rc = lib.fee_run(*_shape(cfg), S, int(tc), _p(np.ascontiguousarray(canonical, np.float32)), mode, _p(state_native), _p(inp), _p(out),
                     _p(spec_out), n_streams, n_hops, L, ld_in, ld_out, _p(dbg), dbg_hop, cfg.input_compression)
    if rc != 0:
        raise RuntimeError(f"fee_run failed rc={rc}")


def tap_total(cfg):
    return _load().fee_tap_total(*_shape(cfg))
