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
    """g++ the emulation: ten translation units (one per model configuration) in parallel + the dispatcher, linked into one .so."""
    import concurrent.futures as cf
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("fe_plan.h", "fe_pack.h", "fe_kernel.cuh", "fe_configs.h", "fe_half.h")]
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return _SO
    bdir = os.path.dirname(_SO)
    os.makedirs(bdir, exist_ok=True)
    jobs = [(f"-DFE_EMU_GROUP={g}", os.path.join(bdir, f"fe_emu_g{g}.o")) for g in range(10)] + [("-DFE_EMU_MAIN", os.path.join(bdir, "fe_emu_main.o"))]

    def cc(job):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", job[0], "-c", _SRC, "-o", job[1]], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"emu build failed ({job[0]}):\n{r.stderr[-4000:]}")
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(cc, jobs))
    subprocess.run(["g++", "-shared", "-o", _SO] + [j[1] for j in jobs], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.fee_run.argtypes = [ctypes.c_int] * 10 + [fp, ctypes.c_int, fp, fp, fp, fp] + [ctypes.c_int] * 3 + \
            [ctypes.c_longlong] * 2 + [fp, ctypes.c_int, ctypes.c_float, ctypes.c_int]
        _lib.fee_tap_total.argtypes = [ctypes.c_int] * 8
        _lib.fee_offline_tp.argtypes = [ctypes.c_int] * 9 + [fp, fp, fp, fp] + [ctypes.c_int] * 3 + [ctypes.c_float]
    return _lib


def _p(a):
    if a is None:
        return ctypes.cast(None, ctypes.POINTER(ctypes.c_float))
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _shape(cfg):
    return (cfg.n_fft, cfg.hop_size, cfg.channels, cfg.n_enc, cfg.rf_channels, cfg.rf_freq, cfg.rf_blocks, cfg.rf_heads)


def to_native(cfg, state):
    """per-stream rows [B, 2*CL + K*F2*C2] -> the kernels' planes (flat): cache_stft [B][CL] | cache_istft [B][CL] | h_k [B][F2*C2]."""
    B, cl, hf = state.shape[0], cfg.cache_len, cfg.rf_freq * cfg.rf_channels
    parts = [state[:, :cl], state[:, cl:2 * cl]] + [state[:, 2 * cl + k * hf: 2 * cl + (k + 1) * hf] for k in range(cfg.rf_blocks)]
    return np.ascontiguousarray(np.concatenate([p.reshape(-1) for p in parts]), np.float32)


def to_canonical(cfg, planes):
    cl, hf, K = cfg.cache_len, cfg.rf_freq * cfg.rf_channels, cfg.rf_blocks
    B = planes.size // (2 * cl + K * hf)
    off, parts = 0, []
    for n in [cl, cl] + [hf] * K:
        parts.append(planes[off:off + B * n].reshape(B, n))
        off += B * n
    return np.ascontiguousarray(np.concatenate(parts, axis=1), np.float32)


def run(cfg, S, canonical, mode, state_native, inp, out, spec_out=None, n_streams=1, n_hops=1, L=0, ld_in=0, ld_out=0,
        dbg=None, dbg_hop=-1, tc=False, slice_hops=0):
    lib = _load()
    rc = lib.fee_run(*_shape(cfg), S, int(tc), _p(np.ascontiguousarray(canonical, np.float32)), mode, _p(state_native), _p(inp), _p(out),
                     _p(spec_out), n_streams, n_hops, L, ld_in, ld_out, _p(dbg), dbg_hop, cfg.input_compression, slice_hops)
    if rc != 0:
        raise RuntimeError(f"fee_run failed rc={rc}")


def tap_total(cfg):
    return _load().fee_tap_total(*_shape(cfg))


def offline_tp(cfg, S, canonical, wav, out, spec_out=None, grid=3):
    """Model.forward on wav [B, L] through the frame-parallel offline schedule (fp32 family): staged launches of the fused kernel body on
    `grid` emulated CTAs, the GRU scan and the overlap-add restated on the host."""
    lib = _load()
    B, L = wav.shape
    rc = lib.fee_offline_tp(*_shape(cfg), S, _p(np.ascontiguousarray(canonical, np.float32)), _p(wav), _p(out), _p(spec_out), B, L, grid,
                            cfg.input_compression)
    if rc != 0:
        raise RuntimeError(f"fee_offline_tp failed rc={rc}")
