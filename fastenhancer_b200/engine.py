"""ctypes binding of the C ABI (include/fastenhancer_b200.h) -- the only route to the CUDA engine.

PyTorch is used for device memory and streams only (tensors are passed as raw device pointers).
There is no CPU path: if the shared library cannot be loaded, or no B200 is visible, construction
raises.
"""
from __future__ import annotations

import ctypes
import os
import typing as tp

import numpy as np

from .config import FEConfig

# FE_LIB overrides the library path (experiments with alternative builds of the same C ABI)
_LIB_PATH = os.environ.get("FE_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfastenhancer_b200.so")
_lib = None

#: every symbol include/fastenhancer_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = (
    "fe_last_error", "fe_weight_count", "fe_state_floats", "fe_create", "fe_destroy", "fe_state_create",
    "fe_state_destroy", "fe_state_reset", "fe_state_export", "fe_state_import", "fe_stream", "fe_stream_host",
    "fe_spec", "fe_stft", "fe_istft", "fe_offline", "fe_streams_per_cta", "fe_set_streams_per_cta", "fe_kernel_launches", "fe_tap_floats",
    "fe_stream_taps", "fe_profile_slots", "fe_set_profile", "fe_set_precision", "fe_get_precision", "fe_state_reserve_host",
    "fe_state_create_on", "fe_state_planes", "fe_microbench_fma", "fe_fold_device", "fe_create_from_device",
    "fe_pcm16_to_float", "fe_resample_poly", "fe_float_to_pcm16", "fe_set_offline_mode", "fe_stft_gemm", "fe_set_hop_slicing", "fe_plan_hop_slices",
)

#: precision name -> fe_set_precision mode (include/fastenhancer_b200.h)
PRECISION_MODES = {"tf32": 0, "fp32": 1, "f16": 2, "bf16": 3, "fp32x3": 4}


class CConfig(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("n_fft", "hop", "c1", "n_enc", "c2", "f2", "n_blocks", "n_heads")] + \
        [("compression", ctypes.c_float)]

    @classmethod
    def from_cfg(cls, cfg: FEConfig) -> "CConfig":
        return cls(cfg.n_fft, cfg.hop_size, cfg.channels, cfg.n_enc, cfg.rf_channels, cfg.rf_freq, cfg.rf_blocks,
                   cfg.rf_heads, cfg.input_compression)


def library_path() -> str:
    return _LIB_PATH


def load_library(build_if_missing: bool = True):
    """dlopen the in-tree engine library and declare the C ABI's signatures."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{_LIB_PATH} is missing: run `python -m fastenhancer_b200.build`")
        from .build import build
        build()
    lib = ctypes.CDLL(_LIB_PATH)
    vp, ip, ll, fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p   # fp: raw float* (host or device)
    cfgp = ctypes.POINTER(CConfig)
    lib.fe_last_error.restype = ctypes.c_char_p
    lib.fe_weight_count.restype = ctypes.c_size_t
    lib.fe_weight_count.argtypes = [cfgp]
    lib.fe_state_floats.restype = ctypes.c_size_t
    lib.fe_state_floats.argtypes = [cfgp]
    lib.fe_create.argtypes = [cfgp, fp, ctypes.c_size_t, ip, ctypes.POINTER(vp)]
    lib.fe_create_from_device.argtypes = [cfgp, fp, ctypes.c_size_t, ip, ctypes.POINTER(vp)]
    lib.fe_destroy.argtypes = [vp]
    lib.fe_destroy.restype = None
    lib.fe_state_create.argtypes = [vp, ip, ctypes.POINTER(vp)]
    lib.fe_state_create_on.argtypes = [vp, ip, fp, ctypes.POINTER(vp)]
    lib.fe_state_planes.argtypes = [vp]
    lib.fe_state_planes.restype = ctypes.c_void_p
    lib.fe_state_destroy.argtypes = [vp]
    lib.fe_state_destroy.restype = None
    lib.fe_state_reset.argtypes = [vp, vp]
    lib.fe_state_export.argtypes = [vp, fp, vp]
    lib.fe_state_import.argtypes = [vp, fp, vp]
    lib.fe_stream.argtypes = [vp, vp, fp, fp, ip, ll, ll, vp]
    lib.fe_stream_host.argtypes = [vp, vp, fp, fp, ip, ll, ll, ip, vp]
    lib.fe_state_reserve_host.argtypes = [vp, ip]
    lib.fe_spec.argtypes = [vp, vp, fp, fp, ip, vp]
    lib.fe_offline.argtypes = [vp, fp, ip, ip, fp, fp, vp]
    lib.fe_stft.argtypes = [vp, vp, fp, fp, ip, ll, vp]
    lib.fe_istft.argtypes = [vp, vp, fp, fp, ip, ll, vp]
    lib.fe_streams_per_cta.argtypes = [vp, ip]
    lib.fe_set_streams_per_cta.argtypes = [vp, ip]
    lib.fe_kernel_launches.argtypes = [vp]
    lib.fe_kernel_launches.restype = ll
    lib.fe_tap_floats.argtypes = [vp]
    lib.fe_stream_taps.argtypes = [vp, vp, fp, fp, ip, ll, ll, fp, ip, vp]
    lib.fe_profile_slots.argtypes = []
    lib.fe_set_profile.argtypes = [vp, vp]
    lib.fe_pcm16_to_float.argtypes = [fp, ll, ip, fp, vp]
    lib.fe_resample_poly.argtypes = [fp, ll, ip, ip, fp, ip, fp, ll, vp]
    lib.fe_float_to_pcm16.argtypes = [fp, ll, fp, vp]
    lib.fe_microbench_fma.argtypes = [ip, ctypes.POINTER(ctypes.c_double)]
    lib.fe_set_precision.argtypes = [vp, ip]
    lib.fe_get_precision.argtypes = [vp]
    lib.fe_set_offline_mode.argtypes = [vp, ip]
    lib.fe_set_hop_slicing.argtypes = [vp, ip]
    lib.fe_plan_hop_slices.argtypes = [ip, ip, ip]
    lib.fe_stft_gemm.argtypes = [vp, fp, ip, ll, ip, fp, ip, vp]
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().fe_last_error()
        raise RuntimeError(f"{what} failed ({rc}): {msg.decode() if msg else 'unknown error'}")


def measured_fma_tflops(device_index: int = 0) -> float:
    """fp32 FMA-pipe throughput of the device, measured (fe_microbench_fma)."""
    v = ctypes.c_double()
    _check(load_library().fe_microbench_fma(int(device_index), ctypes.byref(v)), "fe_microbench_fma")
    return float(v.value)


def _stream_ptr(device=None) -> int:
    """raw cudaStream_t of torch's current stream ON THE ENGINE'S DEVICE (not of whatever device is current)."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


class State:
    """Recurrent (GRU h per block) + overlap (STFT / iSTFT caches) state of ``n_streams`` streams, on device."""

    def __init__(self, engine: "Engine", n_streams: int):
        """The state lives in a torch allocation handed to the engine (fe_state_create_on): the kernels keep it as planes in
        the reference's own cache shapes, so :meth:`caches` returns zero-copy views the reference-shaped wrappers hand out."""
        import torch
        self.engine, self.n_streams = engine, int(n_streams)
        self.buf = torch.zeros(self.n_streams * engine.state_floats, dtype=torch.float32, device=engine.device)
        h = ctypes.c_void_p()
        _check(engine._lib.fe_state_create_on(engine._h, self.n_streams, self.buf.data_ptr(), ctypes.byref(h)), "fe_state_create_on")
        self._h = h
        cfg, B = engine.cfg, self.n_streams
        cl, hf = cfg.cache_len, cfg.rf_freq * cfg.rf_channels
        self._caches = [self.buf[:B * cl].view(B, cl), self.buf[B * cl:2 * B * cl].view(B, cl)] + \
            [self.buf[2 * B * cl + k * B * hf: 2 * B * cl + (k + 1) * B * hf].view(1, B * cfg.rf_freq, cfg.rf_channels)
             for k in range(cfg.rf_blocks)]

        self._live = [None] * len(self._caches)

    # Reference-shaped cache tensors without copies.  Plane i (0 = cache_stft [B, N-H], 1 = cache_istft [B, N-H], 2 + k = h_k
    # [1, B*F2, C2]: the shapes of ONNXSTFT.initialize_cache + ONNXModel.initialize_cache, functional/audio_modules.py:238-241,
    # model.py:614-618) is handed out by emit(i) as a fresh view object of the device state and taken back by adopt(i, t):
    #   * t is the object emit(i) returned last  -> the state already holds it: nothing to do (the per-hop fast path);
    #   * t is some other tensor                 -> its contents are copied in (explicit caches of another session, zeros, ...);
    #   * t is an OLDER view of the same memory  -> its contents are gone (the state has moved on): a loud error, never silence.
    def emit(self, i: int):
        v = self._caches[i].view(self._caches[i].shape)
        self._live[i] = v
        return v

    def adopt(self, i: int, t) -> None:
        c = self._caches[i]
        if t is None:
            c.zero_()
        elif t is self._live[i]:
            return
        elif t.device == c.device and t.data_ptr() == c.data_ptr():
            raise RuntimeError("stale cache tensor: the caches this engine returns are views of its device state and are overwritten "
                               "by the next step; clone() one if you need to keep an older generation")
        else:
            c.copy_(t.reshape(c.shape))
        self._live[i] = None

    def __del__(self):
        if getattr(self, "_h", None):
            self.engine._lib.fe_state_destroy(self._h)
            self._h = None

    def reserve_host(self, hops_per_chunk: int = 64) -> None:
        """pre-allocate the staging buffers / streams of :meth:`Engine.stream_host` (keeps allocations out of the hot call)."""
        _check(self.engine._lib.fe_state_reserve_host(self._h, int(hops_per_chunk)), "fe_state_reserve_host")

    def reset(self) -> None:
        _check(self.engine._lib.fe_state_reset(self._h, _stream_ptr(self.engine.device)), "fe_state_reset")

    def export(self):
        """-> float32 cuda tensor [n_streams, state_floats] in the reference cache layout
        [cache_stft | cache_istft | h_0 [F2][C2] | ...]."""
        import torch
        out = torch.empty(self.n_streams, self.engine.state_floats, dtype=torch.float32, device=self.engine.device)
        _check(self.engine._lib.fe_state_export(self._h, out.data_ptr(), _stream_ptr(self.engine.device)), "fe_state_export")
        return out

    def load(self, t) -> None:
        t = self.engine._dev(t, (self.n_streams, self.engine.state_floats))
        _check(self.engine._lib.fe_state_import(self._h, t.data_ptr(), _stream_ptr(self.engine.device)), "fe_state_import")


class Engine:
    """One folded FastEnhancer model resident on one B200."""

    def __init__(self, cfg: FEConfig, canonical: np.ndarray, device: tp.Union[int, str, None] = None,
                 precision: tp.Optional[str] = None):
        """``precision``: None (default) = results identical to the fp32 reference -- "fp32x3" (fp32-accurate tensor-core
        contractions: split-fp16 operands, three MMAs per product) where the model has such kernels, else "fp32" (everything on
        the fp32 FMA pipe).  Faster, reduced-precision opt-ins: "tf32" (TF32 operands, fp32 accumulate), "f16" (as tf32 with the
        conv section's operands stored as fp16), "bf16" (bfloat16 conv section, TF32 RNNFormer: BASELINE config 3's arithmetic).
        ``FE_PRECISION`` in the environment overrides the default."""
        import torch
        cfg.validate()
        if not torch.cuda.is_available():
            raise RuntimeError("fastenhancer_b200: no CUDA device visible; the engine has no CPU fallback")
        self._lib = load_library()
        self.cfg = cfg
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"fastenhancer_b200: device must be a CUDA device, got {dev}")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self._c = CConfig.from_cfg(cfg)
        need = self._lib.fe_weight_count(ctypes.byref(self._c))
        h = ctypes.c_void_p()
        if isinstance(canonical, torch.Tensor) and canonical.is_cuda:      # already on the device (checkpoint.fold_on_device)
            canonical = canonical.to(device=self.device, dtype=torch.float32).contiguous()
            if canonical.numel() != need:
                raise ValueError(f"canonical weights: got {canonical.numel()} floats, need {need}")
            _check(self._lib.fe_create_from_device(ctypes.byref(self._c), canonical.data_ptr(), canonical.numel(), self.device.index,
                                                   ctypes.byref(h)), "fe_create_from_device")
        else:
            canonical = np.ascontiguousarray(canonical, dtype=np.float32)
            if canonical.size != need:
                raise ValueError(f"canonical weights: got {canonical.size} floats, need {need}")
            _check(self._lib.fe_create(ctypes.byref(self._c), canonical.ctypes.data, canonical.size, self.device.index,
                                       ctypes.byref(h)), "fe_create")
        self._h = h
        self.state_floats = int(self._lib.fe_state_floats(ctypes.byref(self._c)))
        if precision is not None:
            self.set_precision(precision)

    @classmethod
    def from_checkpoint(cls, path: str, device: tp.Union[int, str, None] = None, precision: tp.Optional[str] = None) -> "Engine":
        """``logs/<name>`` (or one of its ``NNNNN.pth`` files) -> engine, with the reference's pre-fold parameters folded on the
        device (fastenhancer_b200.checkpoint; replaces wrapper.load() + remove_weight_reparameterizations(),
        /root/reference/wrappers/ns.py:308-321, models/fastenhancer/default/model.py:532-608)."""
        import torch
        from .checkpoint import fold_on_device, load_checkpoint
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cfg, sd = load_checkpoint(path, device=dev)
        return cls(cfg, fold_on_device(cfg, sd, dev), dev, precision=precision)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.fe_destroy(self._h)
            self._h = None

    # ---- helpers ----
    def _dev(self, t, shape=None):
        import torch
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t))
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def _out(self, out, like):
        """validate a caller-provided output tensor: same device / dtype / shape as ``like``, contiguous rows."""
        import torch
        if out is None:
            return torch.empty_like(like)
        if not isinstance(out, torch.Tensor) or out.device != like.device or out.dtype != torch.float32 or \
                tuple(out.shape) != tuple(like.shape) or (out.dim() > 1 and out.stride(-1) != 1) or \
                (out.dim() > 2 and not out.is_contiguous()):
            raise ValueError(f"out must be a float32 tensor of shape {tuple(like.shape)} on {like.device} with contiguous rows")
        return out

    def set_precision(self, precision: str) -> None:
        if precision not in PRECISION_MODES:
            raise ValueError(f"precision must be one of {sorted(PRECISION_MODES)}")
        _check(self._lib.fe_set_precision(self._h, PRECISION_MODES[precision]), "fe_set_precision")

    def set_hop_slicing(self, on: bool) -> None:
        """Multi-round streaming launches cut into hop ranges on a persistent grid (default on; results are bit-identical)."""
        _check(self._lib.fe_set_hop_slicing(self._h, 1 if on else 0), "fe_set_hop_slicing")

    def set_offline_mode(self, mode: str) -> None:
        """Schedule of ``offline`` (``Model.forward``): 'auto', 'walk' (one CTA per group of utterances steps through the frames) or
        'frame_parallel' (CTAs take groups of frames, the GRU recurrence runs as a scan between the launches; fp32-accurate modes)."""
        modes = {"auto": 0, "walk": 1, "frame_parallel": 2}
        if mode not in modes:
            raise ValueError(f"offline mode must be one of {sorted(modes)}")
        _check(self._lib.fe_set_offline_mode(self._h, modes[mode]), "fe_set_offline_mode")

    @property
    def precision(self) -> str:
        return {v: k for k, v in PRECISION_MODES.items()}[int(self._lib.fe_get_precision(self._h))]

    def new_state(self, n_streams: int) -> State:
        return State(self, n_streams)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.fe_kernel_launches(self._h))

    def streams_per_cta(self, n_streams: int) -> int:
        return int(self._lib.fe_streams_per_cta(self._h, int(n_streams)))

    def set_streams_per_cta(self, s: int) -> None:
        _check(self._lib.fe_set_streams_per_cta(self._h, int(s)), "fe_set_streams_per_cta")

    PHASES = ("init", "load", "window", "fft", "compress", "enc_pre", "enc", "lin_pre", "rf_pre", "hload", "gru", "rnn_fc", "qkv",
              "attn", "attn_fc", "lin_post", "rf_post", "skip_load", "pwcat", "dec", "convt", "mask", "pretw", "ifft", "ola",
              "dbg", "state", "tc:wait_weights", "tc:issue", "tc:mma_done", "tc:tmem_ld", "tc:epi_math")

    N_SUB = 5       # sub-timers (the "tc:*" entries), also recorded per phase

    def enable_profile(self, on: bool = True):
        """Per-phase SM-cycle counters of CTA 0 (int64 cuda tensor, accumulated over launches) or None."""
        import torch
        if on:
            n = int(self._lib.fe_profile_slots())
            assert n == len(self.PHASES) * (1 + self.N_SUB)      # totals, then [phase][sub-timer]
            self._prof = torch.zeros(n, dtype=torch.int64, device=self.device)
            _check(self._lib.fe_set_profile(self._h, self._prof.data_ptr()), "fe_set_profile")
            return self._prof
        _check(self._lib.fe_set_profile(self._h, None), "fe_set_profile")
        self._prof = None
        return None

    # ---- the hot path ----
    def stream(self, state: State, wav_in, out=None):
        """``n_hops`` streaming steps for every stream: wav_in [B, n_hops*hop] (cuda) -> wav_out, one launch."""
        import torch
        x = self._dev(wav_in)
        B, L = x.shape
        H = self.cfg.hop_size
        if B != state.n_streams or L % H:
            raise ValueError(f"wav_in must be [{state.n_streams}, k*{H}], got {tuple(x.shape)}")
        out = self._out(out, x)
        _check(self._lib.fe_stream(self._h, state._h, x.data_ptr(), out.data_ptr(), L // H, x.stride(0), out.stride(0),
                                   _stream_ptr(self.device)), "fe_stream")
        return out

    def stream_taps(self, state: State, wav_in, tap_hop: int):
        import torch
        x = self._dev(wav_in)
        B, L = x.shape
        H = self.cfg.hop_size
        out = torch.empty_like(x)
        taps = torch.zeros(int(self._lib.fe_tap_floats(self._h)), dtype=torch.float32, device=self.device)
        _check(self._lib.fe_stream_taps(self._h, state._h, x.data_ptr(), out.data_ptr(), L // H, x.stride(0), out.stride(0),
                                        taps.data_ptr(), int(tap_hop), _stream_ptr(self.device)), "fe_stream_taps")
        return out, taps

    def stream_host(self, state: State, wav_in, out=None, hops_per_chunk: int = 0):
        """Same as :meth:`stream` for HOST tensors (pinned preferred); copies are pipelined with the kernel."""
        import torch
        if wav_in.device.type != "cpu" or wav_in.dtype != torch.float32 or wav_in.stride(1) != 1:
            raise ValueError("stream_host takes a float32 CPU tensor with contiguous rows")
        B, L = wav_in.shape
        H = self.cfg.hop_size
        if B != state.n_streams or L % H:
            raise ValueError(f"wav_in must be [{state.n_streams}, k*{H}], got {tuple(wav_in.shape)}")
        if out is None:
            out = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
        elif out.device.type != "cpu" or out.dtype != torch.float32 or tuple(out.shape) != (B, L) or out.stride(1) != 1:
            raise ValueError(f"out must be a float32 CPU tensor of shape {(B, L)} with contiguous rows")
        _check(self._lib.fe_stream_host(self._h, state._h, wav_in.data_ptr(), out.data_ptr(), L // H, wav_in.stride(0),
                                        out.stride(0), int(hops_per_chunk), _stream_ptr(self.device)), "fe_stream_host")
        return out

    def spec(self, state: State, spec_in, out=None):
        """ONNXModel.forward on [B, n_fft/2+1, T, 2] spectra (GRU state in ``state``)."""
        import torch
        x = self._dev(spec_in)
        B, NB, T, two = x.shape
        if B != state.n_streams or NB != self.cfg.n_fft // 2 + 1 or two != 2:
            raise ValueError(f"bad spectrum shape {tuple(x.shape)}")
        out = self._out(out, x)
        _check(self._lib.fe_spec(self._h, state._h, x.data_ptr(), out.data_ptr(), T, _stream_ptr(self.device)), "fe_spec")
        return out

    def stft(self, state: State, wav_in):
        """ONNXSTFT.forward for every hop of wav_in [B, n_hops*hop] -> spectrum [B, n_fft/2+1, n_hops, 2]; uses / updates
        the cache_stft part of ``state``."""
        import torch
        x = self._dev(wav_in)
        B, L = x.shape
        H = self.cfg.hop_size
        if B != state.n_streams or L % H:
            raise ValueError(f"wav_in must be [{state.n_streams}, k*{H}], got {tuple(x.shape)}")
        out = torch.empty((B, self.cfg.n_fft // 2 + 1, L // H, 2), dtype=torch.float32, device=self.device)
        _check(self._lib.fe_stft(self._h, state._h, x.data_ptr(), out.data_ptr(), L // H, x.stride(0), _stream_ptr(self.device)), "fe_stft")
        return out

    def istft(self, state: State, spec_in):
        """ONNXSTFT.inverse: spectrum [B, n_fft/2+1, T, 2] -> wav [B, T*hop]; uses / updates the cache_istft part of ``state``."""
        import torch
        x = self._dev(spec_in)
        B, NB, T, two = x.shape
        if B != state.n_streams or NB != self.cfg.n_fft // 2 + 1 or two != 2:
            raise ValueError(f"bad spectrum shape {tuple(x.shape)}")
        out = torch.empty((B, T * self.cfg.hop_size), dtype=torch.float32, device=self.device)
        _check(self._lib.fe_istft(self._h, state._h, x.data_ptr(), out.data_ptr(), T, out.stride(0), _stream_ptr(self.device)), "fe_istft")
        return out

    def stft_gemm(self, wav, n_frames: tp.Optional[int] = None, accurate: bool = True):
        """STFT of wav [B, L] as a tensor-core GEMM (ConvSTFT.forward over a whole signal): frame t = wav[:, t*hop : t*hop + n_fft];
        returns spec [B, n_fft/2+1, T, 2] with T = 1 + (L - n_fft) // hop (or ``n_frames``)."""
        import torch
        x = self._dev(wav)
        B, L = x.shape
        N, H = self.cfg.n_fft, self.cfg.hop_size
        T = 1 + (L - N) // H if n_frames is None else n_frames
        if T < 1 or (T - 1) * H + N > L:
            raise ValueError(f"signal of {L} samples holds fewer than {T} frames")
        if x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0:          # TMA tensor map: 16-byte aligned rows
            pad = (-L) % 4
            x = torch.nn.functional.pad(x, (0, pad)).contiguous()
        spec = torch.empty((B, N // 2 + 1, T, 2), dtype=torch.float32, device=self.device)
        _check(self._lib.fe_stft_gemm(self._h, x.data_ptr(), B, x.stride(0), T, spec.data_ptr(), 1 if accurate else 0, _stream_ptr(self.device)),
               "fe_stft_gemm")
        return spec

    def offline(self, wav, want_spec: bool = True):
        """Model.forward: wav [B, L] -> (wav_hat [B, hop*(L//hop)], spec_hat [B, n_fft/2, 1+L//hop, 2] or None)."""
        import torch
        x = self._dev(wav)
        B, L = x.shape
        H = self.cfg.hop_size
        T = 1 + L // H
        out = torch.empty((B, H * (T - 1)), dtype=torch.float32, device=self.device)
        spec = torch.empty((B, self.cfg.f_in, T, 2), dtype=torch.float32, device=self.device) if want_spec else None
        _check(self._lib.fe_offline(self._h, x.data_ptr(), B, L, out.data_ptr(), spec.data_ptr() if want_spec else None,
                                    _stream_ptr(self.device)), "fe_offline")
        return out, spec
