"""One-time host-side folding of a reference checkpoint into the canonical weight array.

What the reference does in ``ONNXModel.remove_weight_reparameterizations``
(/root/reference/models/fastenhancer/default/model.py:532-608, block part :215-258, final
transposed conv :74-81), restated on numpy arrays:

1. weight-norm removal: ``W = g * v / ||v||_row`` for GRU ``weight_ih_l0 / weight_hh_l0`` and the
   attention ``qkv`` projection (torch ``_weight_norm`` with dim=0);
2. final ConvTranspose1d: ``W = scale * W / max(||W||_F, 1e-12)`` (``normalize_final_conv``) or
   ``scale * W``;
3. eval-mode BatchNorm folded into the preceding bias-free conv / linear:
   ``W' = W * gamma/std``, ``b' = beta - mean*gamma/std``, ``std = sqrt(var + eps)``;
4. optional ``pre_norm`` (affine-free BN before GRU / attention) folded into ``W_ih``/``b_ih`` and
   ``qkv`` weight/bias.

The arithmetic is done in float64 and rounded once to float32.
"""
from __future__ import annotations

import typing as tp

import numpy as np

from .config import FEConfig
from .schema import flatten_canonical


def _np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=np.float64)


def _bn_scale_shift(sd, prefix: str, eps: float) -> tp.Tuple[np.ndarray, np.ndarray]:
    std = np.sqrt(_np(sd[f"{prefix}.running_var"]) + eps)
    scale = _np(sd[f"{prefix}.weight"]) / std
    shift = _np(sd[f"{prefix}.bias"]) - _np(sd[f"{prefix}.running_mean"]) * scale
    return scale, shift


def _weight_normed(sd, base: str, leaf: str, enabled: bool) -> np.ndarray:
    if enabled and f"{base}.parametrizations.{leaf}.original0" in sd:
        g = _np(sd[f"{base}.parametrizations.{leaf}.original0"])
        v = _np(sd[f"{base}.parametrizations.{leaf}.original1"])
        return g * v / np.linalg.norm(v.reshape(v.shape[0], -1), axis=1).reshape(g.shape)
    return _np(sd[f"{base}.{leaf}"])


def fold_state_dict(cfg: FEConfig, sd: tp.Mapping[str, tp.Any]) -> tp.Dict[str, np.ndarray]:
    """pre-fold reference state_dict -> dict of canonical folded tensors (float32)."""
    C1, C2 = cfg.channels, cfg.rf_channels
    out: tp.Dict[str, np.ndarray] = {}

    def conv_bn(w_key: str, bn_prefix: str, eps: float):
        w = _np(sd[w_key])
        s, b = _bn_scale_shift(sd, bn_prefix, eps)
        return w * s.reshape((-1,) + (1,) * (w.ndim - 1)), b

    out["enc_pre.w"], out["enc_pre.b"] = conv_bn("enc_pre.0.weight", "enc_pre.1", cfg.bn_eps)
    for i in range(cfg.n_enc):
        out[f"enc.{i}.w"], out[f"enc.{i}.b"] = conv_bn(f"encoder.{i}.0.weight", f"encoder.{i}.1", cfg.bn_eps)
    out["rf_pre.lin"] = _np(sd["rf_pre.0.weight"])
    w, b = conv_bn("rf_pre.1.weight", "rf_pre.2", cfg.bn_eps)
    out["rf_pre.w"], out["rf_pre.b"] = w[:, :, 0], b

    for k in range(cfg.rf_blocks):
        p = f"rf_block.{k}"
        w_ih = _weight_normed(sd, f"{p}.rnn", "weight_ih_l0", cfg.weight_norm)
        w_hh = _weight_normed(sd, f"{p}.rnn", "weight_hh_l0", cfg.weight_norm)
        b_ih = _np(sd[f"{p}.rnn.bias_ih_l0"])
        b_hh = _np(sd[f"{p}.rnn.bias_hh_l0"])
        qkv_w = _weight_normed(sd, f"{p}.attn.qkv", "weight", cfg.weight_norm)
        qkv_b = _np(sd[f"{p}.attn.qkv.bias"]) if f"{p}.attn.qkv.bias" in sd else np.zeros(3 * C2)
        if cfg.pre_norm:                                   # model.py:233-258
            for norm, which in ((f"{p}.attn_pre_norm", "attn"), (f"{p}.rnn_pre_norm", "rnn")):
                std = np.sqrt(_np(sd[f"{norm}.running_var"]) + cfg.rf_eps)
                beta = -_np(sd[f"{norm}.running_mean"]) / std
                if which == "attn":
                    qkv_b = qkv_b + qkv_w @ beta
                    qkv_w = qkv_w / std[None, :]
                else:
                    b_ih = b_ih + w_ih @ beta
                    w_ih = w_ih / std[None, :]
        out[f"blk.{k}.w_ih"], out[f"blk.{k}.w_hh"] = w_ih, w_hh
        out[f"blk.{k}.b_ih"], out[f"blk.{k}.b_hh"] = b_ih, b_hh
        out[f"blk.{k}.rnn_fc.w"], out[f"blk.{k}.rnn_fc.b"] = conv_bn(f"{p}.rnn_fc.weight", f"{p}.rnn_post_norm", cfg.rf_eps)
        if k == 0:
            out["blk.0.pe"] = _np(sd[f"{p}.pe"]) if f"{p}.pe" in sd else np.zeros((cfg.rf_freq, C2))
        out[f"blk.{k}.qkv.w"], out[f"blk.{k}.qkv.b"] = qkv_w, qkv_b
        out[f"blk.{k}.attn_fc.w"], out[f"blk.{k}.attn_fc.b"] = conv_bn(f"{p}.attn_fc.weight", f"{p}.attn_post_norm", cfg.rf_eps)

    out["rf_post.lin"] = _np(sd["rf_post.0.weight"])
    w, b = conv_bn("rf_post.1.weight", "rf_post.2", cfg.bn_eps)
    out["rf_post.w"], out["rf_post.b"] = w[:, :, 0], b
    for i in range(cfg.n_enc):
        w, b = conv_bn(f"decoder.{i}.0.weight", f"decoder.{i}.1", cfg.bn_eps)
        out[f"dec.{i}.w1"], out[f"dec.{i}.b1"] = w[:, :, 0], b
        out[f"dec.{i}.w2"], out[f"dec.{i}.b2"] = conv_bn(f"decoder.{i}.3.weight", f"decoder.{i}.4", cfg.bn_eps)
    w, b = conv_bn("dec_post.0.weight", "dec_post.1", cfg.bn_eps)
    out["dec_post.w"], out["dec_post.b"] = w[:, :, 0], b
    wt = _np(sd["dec_post.3.weight"])
    if "dec_post.3.scale" in sd and sd["dec_post.3.scale"] is not None:
        scale = _np(sd["dec_post.3.scale"]).reshape(())
        if cfg.normalize_final_conv:
            wt = wt / max(float(np.sqrt((wt * wt).sum())), 1e-12)
        wt = wt * scale
    out["dec_post.wt"], out["dec_post.bt"] = wt, _np(sd["dec_post.3.bias"])
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


def fold_to_canonical(cfg: FEConfig, sd: tp.Mapping[str, tp.Any]) -> np.ndarray:
    return flatten_canonical(cfg, fold_state_dict(cfg, sd))
