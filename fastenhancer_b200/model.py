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
(wrappers/ns.py:308-321).  Parameters are folded (fastenhancer_b200.fold) and packed on first use;
every forward then is a launch of the fused kernel through the C ABI.  There is no PyTorch compute path.
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
            self._engine = Engine(self.cfg, self.canonical_weights(), dev)
        return self._engine

    def _state(self, n_streams: int) -> State:
        st = self._states.get(n_streams)
        if st is None:
            st = self._states[n_streams] = self.engine.new_state(n_streams)
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

    def _pack_state(self, B: int, cache_stft, cache_istft, hs) -> Tensor:
        """reference cache tensors -> [B, state_floats] in the C ABI's export layout."""
        cfg, dev = self.cfg, self.engine.device
        z = torch.zeros(B, cfg.cache_len, device=dev)
        parts = [z if cache_stft is None else cache_stft.to(dev, torch.float32).reshape(B, cfg.cache_len),
                 z if cache_istft is None else cache_istft.to(dev, torch.float32).reshape(B, cfg.cache_len)]
        for k in range(cfg.rf_blocks):
            if hs is None or len(hs) == 0:
                parts.append(torch.zeros(B, cfg.rf_freq * cfg.rf_channels, device=dev))
            else:
                parts.append(hs[k].to(dev, torch.float32).reshape(B, cfg.rf_freq * cfg.rf_channels))
        return torch.cat(parts, dim=1)

    def _unpack_h(self, B: int, flat: Tensor, like: Tensor) -> tp.List[Tensor]:
        cfg = self.cfg
        n = cfg.rf_freq * cfg.rf_channels
        off = 2 * cfg.cache_len
        return [flat[:, off + k * n: off + (k + 1) * n].reshape(1, B * cfg.rf_freq, cfg.rf_channels).to(like.device)
                for k in range(cfg.rf_blocks)]

    @torch.no_grad()
    def forward(self, spec_noisy: Tensor, *args):
        """[B, n_fft/2+1, T, 2] (+ GRU caches) -> (spec_hat, *caches_out); no caches = zero state (model.py:623-626)."""
        B = spec_noisy.size(0)
        if len(args) not in (0, self.cfg.rf_blocks):
            raise ValueError(f"expected 0 or {self.cfg.rf_blocks} cache tensors, got {len(args)}")
        st = self._state(B)
        st.load(self._pack_state(B, None, None, args))
        out = self.engine.spec(st, spec_noisy)
        hs = self._unpack_h(B, st.export(), spec_noisy)
        return (out.to(spec_noisy.device), *hs)

    # per-hop STFT / iSTFT shims (functional/audio_modules.py:243-303); one fused-kernel launch each
    @torch.no_grad()
    def _stft_forward(self, x: Tensor, cache: tp.Optional[Tensor]):
        """ONNXSTFT.forward: x [B, k*hop], cache [B, N-H] -> (spec [B, N/2+1, k, 2], cache)."""
        B = x.size(0)
        st = self._state(B)
        st.load(self._pack_state(B, cache, None, None))
        spec = self.engine.stft(st, x)
        return spec.to(x.device), st.export()[:, :self.cfg.cache_len].to(x.device)

    @torch.no_grad()
    def _stft_inverse(self, spec: Tensor, cache: tp.Optional[Tensor]):
        """ONNXSTFT.inverse: spec [B, N/2+1, T, 2], cache [B, N-H] -> (wav [B, T*hop], cache)."""
        B = spec.size(0)
        st = self._state(B)
        st.load(self._pack_state(B, None, cache, None))
        wav = self.engine.istft(st, spec)
        cl = self.cfg.cache_len
        return wav.to(spec.device), st.export()[:, cl:2 * cl].to(spec.device)


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

    ``run(wav [B, n_hops*hop])`` keeps the state on the device between hops (no cache round trip), which is
    what the engine is built for; ``forward`` keeps the reference's explicit-cache calling convention."""

    def __init__(self, model: ONNXModel):
        super().__init__()
        self.model = model

    def initialize_cache(self, x: Tensor) -> tp.List[Tensor]:
        return self.model.stft.initialize_cache(x) + self.model.initialize_cache(x)

    @torch.no_grad()
    def forward(self, wav_in: Tensor, cache_stft: Tensor, cache_istft: Tensor, *cache_model):
        m = self.model
        B = wav_in.size(0)
        st = m._state(B)
        st.load(m._pack_state(B, cache_stft, cache_istft, cache_model))
        out = m.engine.stream(st, wav_in)
        flat = st.export()
        cl = m.cfg.cache_len
        return (out.to(wav_in.device), flat[:, :cl].to(wav_in.device), flat[:, cl:2 * cl].to(wav_in.device),
                *m._unpack_h(B, flat, wav_in))

    @torch.no_grad()
    def run(self, wav: Tensor, state: tp.Optional[State] = None) -> Tensor:
        m = self.model
        if state is None:
            state = m.engine.new_state(wav.size(0))
        return m.engine.stream(state, wav).to(wav.device)
