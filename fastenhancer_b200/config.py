"""Shape/config description of the FastEnhancer per-frame hot path.

Mirrors the ``model_kwargs`` block of the reference YAMLs
(/root/reference/configs/fastenhancer/{t,b,s,m,l}.yaml:1-29 and
/root/reference/configs/fastenhancer_48khz/{t,b,s,m,l}.yaml:1-29) and the constructor
signature of ``ONNXModel`` (/root/reference/models/fastenhancer/default/model.py:384-403).

Only the options the shipped configs use are accepted by the CUDA engine; anything else is
rejected loudly in :meth:`FEConfig.validate` (no silent fallback).
"""
from __future__ import annotations

import dataclasses
import typing as tp


@dataclasses.dataclass(frozen=True)
class FEConfig:
    # STFT front-end (functional/audio_modules.py:182-236)
    n_fft: int = 512
    hop_size: int = 256
    win_size: int = 512
    # encoder / decoder (model.py:433-521)
    channels: int = 48                     # C1
    kernel_size: tp.Tuple[int, ...] = (8, 3, 3)
    stride: int = 4
    # RNNFormer (model.py:294-305, 467-483)
    rf_blocks: int = 3                     # K
    rf_channels: int = 36                  # C2
    rf_freq: int = 24                      # F2
    rf_heads: int = 4                      # NH
    rf_eps: float = 1e-5
    positional_embedding: tp.Optional[str] = "train"
    attn_bias: bool = False
    post_act: bool = False
    pre_norm: bool = False
    # misc
    pre_post_init: tp.Optional[str] = "linear_fixed"
    window: tp.Optional[str] = "hann"
    stft_normalized: bool = False
    mask: tp.Optional[str] = None
    activation: str = "SiLU"
    input_compression: float = 0.3
    weight_norm: bool = True
    normalize_final_conv: bool = True
    resnet: bool = False
    bn_eps: float = 1e-5                   # nn.BatchNorm1d default used by enc/dec (model.py:441)
    sample_rate: int = 16_000

    # ---- derived shapes (SURVEY.md section 8 notation) ----
    @property
    def f_in(self) -> int:                 # Fin = N/2 (Nyquist dropped, model.py:684)
        return self.n_fft // 2

    @property
    def f1(self) -> int:                   # F1 = Fin / stride (model.py:459)
        return self.f_in // self.stride

    @property
    def n_enc(self) -> int:                # E = len(kernel_size) - 1 (model.py:447)
        return len(self.kernel_size) - 1

    @property
    def head_dim(self) -> int:
        return self.rf_channels // self.rf_heads

    @property
    def cache_len(self) -> int:            # N - H (audio_modules.py:197)
        return self.n_fft - self.hop_size

    @property
    def state_floats(self) -> int:
        """floats of recurrent + overlap state per stream."""
        return 2 * self.cache_len + self.rf_blocks * self.rf_freq * self.rf_channels

    def macs_per_frame(self) -> int:
        """Closed-form MACs per frame; same terms as the reference's MAC counter
        (/root/reference/models/fastenhancer/default/macs.py:17-87) with T=1."""
        C1, C2, F1, F2, K = self.channels, self.rf_channels, self.f1, self.rf_freq, self.rf_blocks
        k0 = self.kernel_size[0]
        macs = 2 * C1 * k0 * F1
        for k in self.kernel_size[1:]:
            macs += C1 * C1 * k * F1
        macs += F1 * F2 * C1 + C1 * C2 * F2
        per_block = C2 * C2 * 6 * F2 + C2 * C2 * F2
        per_block += C2 * C2 * 3 * F2 + F2 * C2 * F2 + F2 * F2 * C2 + C2 * C2 * F2
        macs += K * per_block
        macs += F2 * F1 * C2 + C2 * C1 * F1
        for k in reversed(self.kernel_size[1:]):
            macs += 2 * C1 * C1 * F1 + C1 * C1 * k * F1
        macs += 2 * C1 * C1 * F1 + C1 * 2 * k0 * F1
        return macs

    def flops_per_frame(self) -> float:
        """Algorithmic FLOP per frame used by the roofline:
        2*MAC + 2*(2.5*N*log2 N) for the forward + inverse FFT."""
        import math
        return 2.0 * self.macs_per_frame() + 2.0 * (2.5 * self.n_fft * math.log2(self.n_fft))

    def frames_per_second(self) -> float:
        return self.sample_rate / self.hop_size

    def validate(self) -> None:
        def req(cond: bool, msg: str) -> None:
            if not cond:
                raise ValueError(f"fastenhancer_b200: unsupported model_kwargs: {msg}")
        req(self.n_fft % 2 == 0, "n_fft must be even")
        req(self.win_size == self.n_fft, "win_size must equal n_fft")
        req(self.window == "hann", "window must be 'hann'")
        req(not self.stft_normalized, "stft_normalized must be False")
        req(self.stride == 4 and self.kernel_size[0] == 8, "stride must be 4 and kernel_size[0] 8")
        req(all(k == 3 for k in self.kernel_size[1:]), "kernel_size[1:] must all be 3")
        req(self.activation == "SiLU", "activation must be SiLU")
        req(self.mask is None, "mask must be null")
        req(not self.resnet, "resnet must be False")
        req(not self.post_act, "rnnformer post_act must be False")
        req(self.rf_channels % self.rf_heads == 0, "rf channels must be divisible by heads")
        req(0 < self.hop_size <= self.n_fft, "hop_size must be in (0, n_fft]")

    # ---- construction from the reference's model_kwargs dict ----
    @classmethod
    def from_model_kwargs(cls, kw: tp.Mapping[str, tp.Any], sample_rate: int = 16_000) -> "FEConfig":
        kw = dict(kw)
        rf = dict(kw.pop("rnnformer_kwargs", {}) or {})
        kw.pop("activation_kwargs", None)
        rf.pop("p_dropout", None)          # inference: dropout is identity (model.py:198)
        cfg = cls(
            n_fft=int(kw.pop("n_fft", 512)),
            hop_size=int(kw.pop("hop_size", 256)),
            win_size=int(kw.pop("win_size", 512)),
            channels=int(kw.pop("channels", 64)),
            kernel_size=tuple(int(k) for k in kw.pop("kernel_size", (8, 3, 3))),
            stride=int(kw.pop("stride", 4)),
            rf_blocks=int(rf.pop("num_blocks", 3)),
            rf_channels=int(rf.pop("channels", 32)),
            rf_freq=int(rf.pop("freq", 32)),
            rf_heads=int(rf.pop("num_heads", 4)),
            rf_eps=float(rf.pop("eps", 1e-8)),
            positional_embedding=rf.pop("positional_embedding", "train"),
            attn_bias=bool(rf.pop("attn_bias", False)),
            post_act=bool(rf.pop("post_act", False)),
            pre_norm=bool(rf.pop("pre_norm", False)),
            pre_post_init=kw.pop("pre_post_init", None),
            window=kw.pop("window", "hann"),
            stft_normalized=bool(kw.pop("stft_normalized", False)),
            mask=kw.pop("mask", None),
            activation=kw.pop("activation", "ReLU"),
            input_compression=float(kw.pop("input_compression", 0.3)),
            weight_norm=bool(kw.pop("weight_norm", False)),
            normalize_final_conv=bool(kw.pop("normalize_final_conv", False)),
            resnet=bool(kw.pop("resnet", False)),
            sample_rate=sample_rate,
        )
        if kw or rf:
            raise TypeError(f"unexpected model_kwargs: {sorted(kw) + sorted(rf)}")
        return cfg

    def to_model_kwargs(self) -> tp.Dict[str, tp.Any]:
        """The dict the reference's ``Model(**model_kwargs)`` takes for this config."""
        return dict(
            channels=self.channels, kernel_size=list(self.kernel_size), stride=self.stride,
            rnnformer_kwargs=dict(
                num_blocks=self.rf_blocks, channels=self.rf_channels, freq=self.rf_freq,
                num_heads=self.rf_heads, eps=self.rf_eps,
                positional_embedding=self.positional_embedding, attn_bias=self.attn_bias,
                post_act=self.post_act, pre_norm=self.pre_norm),
            pre_post_init=self.pre_post_init, n_fft=self.n_fft, hop_size=self.hop_size,
            win_size=self.win_size, window=self.window, stft_normalized=self.stft_normalized,
            mask=self.mask, activation=self.activation, activation_kwargs=dict(inplace=True),
            input_compression=self.input_compression,
            normalize_final_conv=self.normalize_final_conv, weight_norm=self.weight_norm,
            resnet=self.resnet,
        )


def _mk(sr, n_fft, hop, c1, n_enc, c2, f2, k, init) -> FEConfig:
    return FEConfig(n_fft=n_fft, hop_size=hop, win_size=n_fft, channels=c1,
                    kernel_size=(8,) + (3,) * n_enc, rf_blocks=k, rf_channels=c2, rf_freq=f2,
                    pre_post_init=init, sample_rate=sr)


#: The ten shipped configurations.
PRESETS: tp.Dict[str, FEConfig] = {
    "16k_t": _mk(16_000, 512, 256, 24, 2, 20, 16, 2, "linear_fixed"),
    "16k_b": _mk(16_000, 512, 256, 48, 2, 36, 24, 3, "linear_fixed"),
    "16k_s": _mk(16_000, 512, 256, 64, 3, 48, 36, 3, "linear_fixed"),
    "16k_m": _mk(16_000, 512, 160, 96, 3, 72, 48, 4, "linear_fixed"),
    "16k_l": _mk(16_000, 512, 100, 128, 4, 96, 64, 5, "linear_fixed"),
    "48k_t": _mk(48_000, 1024, 512, 24, 2, 20, 24, 2, "linear"),
    "48k_b": _mk(48_000, 1024, 512, 48, 2, 36, 36, 3, "linear"),
    "48k_s": _mk(48_000, 1024, 512, 64, 3, 48, 48, 3, "linear"),
    "48k_m": _mk(48_000, 1024, 320, 96, 3, 72, 72, 4, "linear"),
    "48k_l": _mk(48_000, 1024, 200, 128, 4, 96, 96, 5, "linear"),
}


def preset(name: str) -> FEConfig:
    try:
        return PRESETS[name]
    except KeyError:
        raise KeyError(f"unknown preset {name!r}; available: {sorted(PRESETS)}") from None
