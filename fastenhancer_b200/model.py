"""Reference-compatible ``Model`` / ``ONNXModel`` whose forward passes run on the fused CUDA engine.

This is the host-side mirror of the reference's operator interface for the hot path
(/root/reference/models/fastenhancer/default/model.py):

* ``Model(**model_kwargs).forward(noisy [B, L]) -> (wav_hat [B, hop*(L//hop)], spec_hat [B, n_fft/2, T, 2])``
  -- model.py:728-735, what scripts/test_pytorch.py:34 and wrappers/ns.py:240 call;
* ``ONNXModel(**model_kwargs).forward(spec [B, n_fft/2+1, T, 2], *h) -> (spec_hat, *h_out)`` with
  ``h_k [1, B*F2, C2]`` -- model.py:677-710, the spec2spec export graph (scripts/export_onnx_spec.py);
* ``ONNXModel.stft`` with ``forward(x, cache)``, ``inverse(spec, cache)``, ``initialize_cache(x)`` --
  functional/audio_modules.py:238-303, and ``ONNXModel.initialize_cache(x)`` -- model.py:614-618;
* :class:`StreamingModel` -- the wav2wav streaming graph of scripts/export_onnx.py:37-58
  ``(wav_in, cache_stft, cache_istft, *h) -> (wav_out, cache_stft, cache_istft, *h)``.

The modules hold the reference's *pre-fold* parameters under the reference's own names, so
``load_state_dict(ckpt['model'], strict=True)`` works on a reference checkpoint
(wrappers/ns.py:308-321).  Parameters are folded (fastenhancer_b200.fold) and packed on first use -- a snapshot: ``load_state_dict`` / ``.to()`` re-pack,
in-place edits of a parameter afterwards need ``remove_weight_reparameterizations()`` to be picked up;
every forward then is a launch of the fused kernel through the C ABI.  There is no PyTorch compute path.
Arithmetic: by default results identical to the fp32 reference (fp32-accurate tensor-core kernels where the model has them,
else the fp32 FMA pipe); ``model.precision = "f16" | "tf32" | "bf16"`` opts into the faster reduced-precision kernels.
``ONNXModel.stft(x, cache)`` / ``.stft.inverse(spec, cache)`` run the front / back end of the same kernel on their own.
"""
from __future__ import annotations

import typing as tp

import numpy as np
import torch
from torch import Tensor, nn

from .config import FEConfig
from .engine import Engine, State
from .fold import fold_to_canonical
from .schema import state_dict_schema, synthetic_state_dict


def _register(root: nn.Module, name: str, value: Tensor, kind: str) -> None:
    """Create the nested container modules of a dotted reference parameter name and register the leaf."""
    parts = name.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, nn.Module())
        mod = mod._modules[p]
    if kind == "param":
        mod.register_parameter(parts[-1], nn.Parameter(value, requires_grad=False))
    else:
        mod.register_buffer(parts[-1], value, persistent=True)


class _StftShim(nn.Module):
    """``model.stft`` of the reference: exposes n_fft / hop_size / window and, for ONNXModel, the per-hop
    forward / inverse with explicit caches (functional/audio_modules.py:182-303)."""

    def __init__(self, owner: "ONNXModel", streaming: bool):
        super().__init__()
        cfg = owner.cfg
        self.n_fft, self.hop_size, self.win_size = cfg.n_fft, cfg.hop_size, cfg.win_size
        self.cache_len = cfg.n_fft - cfg.hop_size
        self.normalized = False
        self.register_buffer("window", torch.hann_window(cfg.win_size), persistent=False)
        self._owner = [owner]            # list: do not register the owner as a sub-module
        self._streaming = streaming

    def initialize_cache(self, x: Tensor) -> tp.List[Tensor]:
        z = lambda: torch.zeros(x.size(0), self.cache_len, dtype=x.dtype, device=x.device)  # noqa: E731
        return [z(), z()]

    def forward(self, x: Tensor, cache: tp.Optional[Tensor] = None):
        if not self._streaming:
            raise RuntimeError("Model.stft is fused into Model.forward; use ONNXModel.stft for per-hop STFT")
        return self._owner[0]._stft_forward(x, cache)

    def inverse(self, spec: Tensor, cache: tp.Optional[Tensor] = None):
        if not self._streaming:
            raise RuntimeError("Model.stft.inverse is fused into Model.forward; use ONNXModel.stft for per-hop iSTFT")
        return self._owner[0]._stft_inverse(spec, cache)


class ONNXModel(nn.Module):
    """Streaming (spec -> spec with GRU caches) FastEnhancer on the fused CUDA engine."""

    _streaming_stft = True

    def __init__(self, **model_kwargs):
        super().__init__()
        sample_rate = int(model_kwargs.pop("sample_rate", 48_000 if model_kwargs.get("n_fft", 512) >= 1024 else 16_000))
        self.cfg = FEConfig.from_model_kwargs(model_kwargs, sample_rate=sample_rate)
        self.cfg.validate()
        self.input_compression = self.cfg.input_compression
        self.rf_ch, self.rf_freq = self.cfg.rf_channels, self.cfg.rf_freq
        self.weight_norm, self.resnet = self.cfg.weight_norm, self.cfg.resnet
        # same names / shapes as the reference's state_dict; values = seeded synthetic checkpoint (no checkpoint
        # ships with the reference), replaced by load_state_dict()
        init = synthetic_state_dict(self.cfg, seed=0)
        for name, shape, kind in state_dict_schema(self.cfg):
            _register(self, name, torch.from_numpy(np.array(init[name])).reshape(shape), kind if kind == "param" else "buffer")
        self.stft = _StftShim(self, self._streaming_stft)
        self._engine: tp.Optional[Engine] = None
        self._states: tp.Dict[int, State] = {}
        self.precision: tp.Optional[str] = None      # None = identical-to-reference default (Engine picks fp32x3 or fp32)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    # ---- engine lifetime ----
    def _invalidate(self) -> None:
        self._engine = None
        self._states = {}

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._invalidate()
        return out

    def canonical_weights(self) -> np.ndarray:
        sd = {k: v.detach().cpu().numpy() for k, v in self.state_dict().items()}
        return fold_to_canonical(self.cfg, sd)

    @property
    def engine(self) -> Engine:
        if self._engine is None:
            dev = self.stft.window.device
            if dev.type != "cuda":
                if not torch.cuda.is_available():
                    raise RuntimeError("fastenhancer_b200: no CUDA device visible; the engine has no CPU fallback")
                dev = torch.device("cuda", torch.cuda.current_device())
            self._engine = Engine(self.cfg, self.canonical_weights(), dev, precision=self.precision)
        elif self.precision is not None and self._engine.precision != self.precision:
            self._engine.set_precision(self.precision)
        return self._engine

    _MAX_STATES = 4      # batch sizes kept alive (least recently used goes first)

    def _state(self, n_streams: int) -> State:
        """the device state of this module's streaming session for ``n_streams`` streams (one per batch size)."""
        st = self._states.pop(n_streams, None)
        if st is None:
            st = self.engine.new_state(n_streams)
            while len(self._states) >= self._MAX_STATES:
                self._states.pop(next(iter(self._states)))
        self._states[n_streams] = st
        return st

    # ---- reference API ----
    def remove_weight_reparameterizations(self) -> None:
        """Reference: folds weight-norm / BatchNorm in place (model.py:532-608).  Here folding happens when the
        engine packs its weights; calling this just forces that to happen now."""
        self._invalidate()
        _ = self.engine

    def flatten_parameters(self) -> None:   # model.py:610-612 (cuDNN GRU detail; nothing to do)
        return None

    def initialize_cache(self, x: Tensor) -> tp.List[Tensor]:
        """GRU caches, one per RNNFormer block: [1, B*F2, C2] zeros (model.py:263-264, 614-618; B = x.size(0))."""
        return [x.new_zeros(1, x.size(0) * self.cfg.rf_freq, self.cfg.rf_channels) for _ in range(self.cfg.rf_blocks)]

    @torch.no_grad()
    def forward(self, spec_noisy: Tensor, *args):
        """[B, n_fft/2+1, T, 2] (+ GRU caches) -> (spec_hat, *caches_out); no caches = zero state (model.py:623-626).
        The returned caches are views of the engine's device state (see fastenhancer_b200.engine.State.emit): feeding them back
        into the next call -- what scripts/test_onnx_spec.py:55-62 does -- costs nothing; foreign tensors are copied in."""
        B, K = spec_noisy.size(0), self.cfg.rf_blocks
        if len(args) not in (0, K):
            raise ValueError(f"expected 0 or {K} cache tensors, got {len(args)}")
        st = self._state(B)
        for k in range(K):
            st.adopt(2 + k, args[k] if args else None)
        out = self.engine.spec(st, spec_noisy)
        return (out.to(spec_noisy.device), *[st.emit(2 + k) for k in range(K)])

    # per-hop STFT / iSTFT shims (functional/audio_modules.py:243-303); one fused-kernel launch each
    @torch.no_grad()
    def _stft_forward(self, x: Tensor, cache: tp.Optional[Tensor]):
        """ONNXSTFT.forward: x [B, k*hop], cache [B, N-H] -> (spec [B, N/2+1, k, 2], cache)."""
        st = self._state(x.size(0))
        st.adopt(0, cache)
        spec = self.engine.stft(st, x)
        return spec.to(x.device), st.emit(0)

    @torch.no_grad()
    def _stft_inverse(self, spec: Tensor, cache: tp.Optional[Tensor]):
        """ONNXSTFT.inverse: spec [B, N/2+1, T, 2], cache [B, N-H] -> (wav [B, T*hop], cache)."""
        st = self._state(spec.size(0))
        st.adopt(1, cache)
        wav = self.engine.istft(st, spec)
        return wav.to(spec.device), st.emit(1)


class Model(ONNXModel):
    """Offline wav -> wav (model.py:713-735)."""

    _streaming_stft = False

    @torch.no_grad()
    def forward(self, noisy: Tensor):
        wav, spec = self.engine.offline(noisy, want_spec=True)
        return wav.to(noisy.device), spec.to(noisy.device)


class StreamingModel(nn.Module):
    """The wav2wav streaming graph of scripts/export_onnx.py:37-58, one fused launch per call:
    ``(wav_in [B, hop], cache_stft [B, N-H], cache_istft [B, N-H], *h [1, B*F2, C2]) -> (wav_out, caches...)``.

    ``forward`` keeps the reference's explicit-cache calling convention (the loop of scripts/test_onnx.py:44-49): the caches it
    returns are views of the engine's device state, so feeding them back costs nothing and a hop is ONE kernel launch; any other
    cache tensors are copied in first.  ``run(wav [B, n_hops*hop])`` processes many hops in one launch on a state the module
    keeps between calls (``reset()`` zeroes it)."""

    def __init__(self, model: ONNXModel):
        super().__init__()
        self.model = model
        self._run_states: tp.Dict[int, State] = {}

    def initialize_cache(self, x: Tensor) -> tp.List[Tensor]:
        return self.model.stft.initialize_cache(x) + self.model.initialize_cache(x)

    @torch.no_grad()
    def forward(self, wav_in: Tensor, cache_stft: Tensor, cache_istft: Tensor, *cache_model):
        m = self.model
        n = 2 + m.cfg.rf_blocks
        if len(cache_model) != m.cfg.rf_blocks:
            raise ValueError(f"expected {m.cfg.rf_blocks} GRU cache tensors, got {len(cache_model)}")
        st = m._state(wav_in.size(0))
        for i, t in enumerate((cache_stft, cache_istft, *cache_model)):
            st.adopt(i, t)
        out = m.engine.stream(st, wav_in)
        return (out.to(wav_in.device), *[st.emit(i) for i in range(n)])

    def reset(self) -> None:
        """zero the state(s) ``run`` keeps between calls."""
        for st in self._run_states.values():
            st.reset()

    @torch.no_grad()
    def run(self, wav: Tensor, state: tp.Optional[State] = None) -> Tensor:
        """many hops in one launch.  Without ``state`` the module's own persistent state for this batch size is used and
        carries over from call to call (chunked callers keep their GRU / overlap state); call ``reset()`` to start over."""
        m = self.model
        if state is None:
            B = wav.size(0)
            if B not in self._run_states or self._run_states[B].engine is not m.engine:
                self._run_states[B] = m.engine.new_state(B)
            state = self._run_states[B]
        return m.engine.stream(state, wav).to(wav.device)
