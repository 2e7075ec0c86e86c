"""Parameter schema of the reference checkpoint format and the canonical folded order.

Two name spaces live here:

* the **pre-fold state_dict** of the reference ``Model``/``ONNXModel``
  (/root/reference/models/fastenhancer/default/model.py:436-521 for the module tree,
  :187-213 for the RNNFormer block; weight-norm parametrisations appear as
  ``parametrizations.<w>.original0`` (g) / ``original1`` (v)); ``state_dict_schema`` lists every
  key with its shape so the engine's nn.Module can ``load_state_dict(strict=True)`` a reference
  checkpoint without instantiating any reference class;
* the **canonical folded order** -- the flat float32 array the C-ABI ``fe_create`` and the CPU
  oracle take (declared in include/fastenhancer_b200.h).  ``canonical_schema`` lists its
  tensors in order.
"""
from __future__ import annotations

import typing as tp

import numpy as np

from .config import FEConfig

Shape = tp.Tuple[int, ...]


def _bn(prefix: str, c: int) -> tp.List[tp.Tuple[str, Shape, str]]:
    return [
        (f"{prefix}.weight", (c,), "param"),
        (f"{prefix}.bias", (c,), "param"),
        (f"{prefix}.running_mean", (c,), "buffer"),
        (f"{prefix}.running_var", (c,), "buffer"),
        (f"{prefix}.num_batches_tracked", (), "buffer_long"),
    ]


def state_dict_schema(cfg: FEConfig) -> tp.List[tp.Tuple[str, Shape, str]]:
    """(name, shape, kind) for every entry of the reference's pre-fold state_dict, in the
    reference's registration order.  kind is 'param' | 'buffer' | 'buffer_long'."""
    C1, C2, F1, F2 = cfg.channels, cfg.rf_channels, cfg.f1, cfg.rf_freq
    s = cfg.stride
    out: tp.List[tp.Tuple[str, Shape, str]] = []
    out.append(("enc_pre.0.weight", (C1, 2 * s, cfg.kernel_size[0] // s), "param"))
    out += _bn("enc_pre.1", C1)
    for i in range(cfg.n_enc):
        out.append((f"encoder.{i}.0.weight", (C1, C1, cfg.kernel_size[i + 1]), "param"))
        out += _bn(f"encoder.{i}.1", C1)
    lin_kind = "buffer" if (cfg.pre_post_init or "").endswith("_fixed") else "param"
    out.append(("rf_pre.0.weight", (F2, F1), lin_kind))
    out.append(("rf_pre.1.weight", (C2, C1, 1), "param"))
    out += _bn("rf_pre.2", C2)
    for k in range(cfg.rf_blocks):
        p = f"rf_block.{k}"
        if k == 0 and cfg.positional_embedding is not None:
            out.append((f"{p}.pe", (F2, C2), "param" if cfg.positional_embedding == "train" else "buffer"))
        if cfg.pre_norm:
            out += [(f"{p}.rnn_pre_norm.running_mean", (C2,), "buffer"),
                    (f"{p}.rnn_pre_norm.running_var", (C2,), "buffer"),
                    (f"{p}.rnn_pre_norm.num_batches_tracked", (), "buffer_long")]
        if cfg.weight_norm:
            out += [(f"{p}.rnn.bias_ih_l0", (3 * C2,), "param"),
                    (f"{p}.rnn.bias_hh_l0", (3 * C2,), "param"),
                    (f"{p}.rnn.parametrizations.weight_ih_l0.original0", (3 * C2, 1), "param"),
                    (f"{p}.rnn.parametrizations.weight_ih_l0.original1", (3 * C2, C2), "param"),
                    (f"{p}.rnn.parametrizations.weight_hh_l0.original0", (3 * C2, 1), "param"),
                    (f"{p}.rnn.parametrizations.weight_hh_l0.original1", (3 * C2, C2), "param")]
        else:
            out += [(f"{p}.rnn.weight_ih_l0", (3 * C2, C2), "param"),
                    (f"{p}.rnn.weight_hh_l0", (3 * C2, C2), "param"),
                    (f"{p}.rnn.bias_ih_l0", (3 * C2,), "param"),
                    (f"{p}.rnn.bias_hh_l0", (3 * C2,), "param")]
        out.append((f"{p}.rnn_fc.weight", (C2, C2), "param"))
        out += _bn(f"{p}.rnn_post_norm", C2)
        if cfg.pre_norm:
            out += [(f"{p}.attn_pre_norm.running_mean", (C2,), "buffer"),
                    (f"{p}.attn_pre_norm.running_var", (C2,), "buffer"),
                    (f"{p}.attn_pre_norm.num_batches_tracked", (), "buffer_long")]
        if cfg.attn_bias:
            out.append((f"{p}.attn.qkv.bias", (3 * C2,), "param"))
        if cfg.weight_norm:
            out += [(f"{p}.attn.qkv.parametrizations.weight.original0", (3 * C2, 1), "param"),
                    (f"{p}.attn.qkv.parametrizations.weight.original1", (3 * C2, C2), "param")]
        else:
            out.append((f"{p}.attn.qkv.weight", (3 * C2, C2), "param"))
        out.append((f"{p}.attn_fc.weight", (C2, C2), "param"))
        out += _bn(f"{p}.attn_post_norm", C2)
    out.append(("rf_post.0.weight", (F1, F2), lin_kind))
    out.append(("rf_post.1.weight", (C1, C2, 1), "param"))
    out += _bn("rf_post.2", C1)
    for i in range(cfg.n_enc):
        ks = cfg.kernel_size[cfg.n_enc - i]
        out.append((f"decoder.{i}.0.weight", (C1, 2 * C1, 1), "param"))
        out += _bn(f"decoder.{i}.1", C1)
        out.append((f"decoder.{i}.3.weight", (C1, C1, ks), "param"))
        out += _bn(f"decoder.{i}.4", C1)
    out.append(("dec_post.0.weight", (C1, 2 * C1, 1), "param"))
    out += _bn("dec_post.1", C1)
    out.append(("dec_post.3.weight", (C1, 2, cfg.kernel_size[0]), "param"))
    out.append(("dec_post.3.bias", (2,), "param"))
    out.append(("dec_post.3.scale", (1,), "param"))
    return out


def canonical_schema(cfg: FEConfig) -> tp.List[tp.Tuple[str, Shape]]:
    """Tensors of the canonical folded weight array, in order (all float32, C-contiguous)."""
    C1, C2, F1, F2 = cfg.channels, cfg.rf_channels, cfg.f1, cfg.rf_freq
    out: tp.List[tp.Tuple[str, Shape]] = [("enc_pre.w", (C1, 8, 2)), ("enc_pre.b", (C1,))]
    for i in range(cfg.n_enc):
        out += [(f"enc.{i}.w", (C1, C1, 3)), (f"enc.{i}.b", (C1,))]
    out += [("rf_pre.lin", (F2, F1)), ("rf_pre.w", (C2, C1)), ("rf_pre.b", (C2,))]
    for k in range(cfg.rf_blocks):
        out += [(f"blk.{k}.w_ih", (3 * C2, C2)), (f"blk.{k}.w_hh", (3 * C2, C2)),
                (f"blk.{k}.b_ih", (3 * C2,)), (f"blk.{k}.b_hh", (3 * C2,)),
                (f"blk.{k}.rnn_fc.w", (C2, C2)), (f"blk.{k}.rnn_fc.b", (C2,))]
        if k == 0:
            out.append(("blk.0.pe", (F2, C2)))
        out += [(f"blk.{k}.qkv.w", (3 * C2, C2)), (f"blk.{k}.qkv.b", (3 * C2,)),
                (f"blk.{k}.attn_fc.w", (C2, C2)), (f"blk.{k}.attn_fc.b", (C2,))]
    out += [("rf_post.lin", (F1, F2)), ("rf_post.w", (C1, C2)), ("rf_post.b", (C1,))]
    for i in range(cfg.n_enc):
        out += [(f"dec.{i}.w1", (C1, 2 * C1)), (f"dec.{i}.b1", (C1,)),
                (f"dec.{i}.w2", (C1, C1, 3)), (f"dec.{i}.b2", (C1,))]
    out += [("dec_post.w", (C1, 2 * C1)), ("dec_post.b", (C1,)),
            ("dec_post.wt", (C1, 2, 8)), ("dec_post.bt", (2,))]
    return out


def canonical_size(cfg: FEConfig) -> int:
    return int(sum(int(np.prod(s)) for _, s in canonical_schema(cfg)))


def flatten_canonical(cfg: FEConfig, tensors: tp.Mapping[str, np.ndarray]) -> np.ndarray:
    parts = []
    for name, shape in canonical_schema(cfg):
        a = np.asarray(tensors[name], dtype=np.float32)
        if a.shape != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
        parts.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts))


def split_canonical(cfg: FEConfig, flat: np.ndarray) -> tp.Dict[str, np.ndarray]:
    out, off = {}, 0
    for name, shape in canonical_schema(cfg):
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n
    if off != flat.size:
        raise ValueError(f"canonical array has {flat.size} floats, schema needs {off}")
    return out


# ---------------------------------------------------------------------------------------------
# Seeded synthetic checkpoint (SURVEY.md section 8(c) "suggested synthetic weights"): default-init-like
# weights, non-trivial BN statistics and weight-norm gains so that folding is exercised, and
# dec_post bias = [1, 0] so the mask is ~identity and output RMS ~ input RMS.
# numpy RandomState is used because its stream is frozen across numpy versions.
# ---------------------------------------------------------------------------------------------
def linear_interp_filterbank(n_freq: int, n_filter: int) -> tp.Tuple[np.ndarray, np.ndarray]:
    """Triangular interpolation filterbank of ``pre_post_init='linear*'``
    (/root/reference/models/fastenhancer/default/model.py:322-325,360-369): filter centres on a
    uniform grid, rows of ``pre`` normalised to sum 1, ``post`` = row-normalised transpose."""
    delta = (n_freq - 1) / (n_filter - 1)
    centres = np.linspace(0.0, n_freq - 1, n_filter, dtype=np.float32)
    bins = np.linspace(0.0, n_freq - 1, n_freq, dtype=np.float32)
    down = np.ones((n_filter, n_freq), np.float32)
    up = np.ones((n_filter, n_freq), np.float32)
    down[:-1] = (centres[1:, None] - bins[None, :]) / np.float32(delta)
    up[1:] = (bins[None, :] - centres[:-1, None]) / np.float32(delta)
    pre = np.maximum(np.float32(0), np.minimum(down, up))
    pre = pre / pre.sum(axis=1, keepdims=True)
    post = pre.T / pre.T.sum(axis=1, keepdims=True)
    return np.ascontiguousarray(pre, np.float32), np.ascontiguousarray(post, np.float32)


def sinusoid_pe(channels: int, freq: int) -> np.ndarray:
    """Initial value of the learned positional embedding (model.py:98-110)."""
    f = np.arange(1, freq + 1, dtype=np.float32) * np.float32(np.pi / freq)
    c = np.exp(np.linspace(np.log(1.0), np.log(freq - 1.0), channels // 2, dtype=np.float32))
    grid = f[:, None] * c[None, :]
    return np.concatenate([np.sin(grid), np.cos(grid)], axis=1).astype(np.float32)


def synthetic_state_dict(cfg: FEConfig, seed: int = 0) -> tp.Dict[str, np.ndarray]:
    rs = np.random.RandomState(seed)
    pre_w, post_w = linear_interp_filterbank(cfg.f1, cfg.rf_freq)
    sd: tp.Dict[str, np.ndarray] = {}
    for name, shape, kind in state_dict_schema(cfg):
        leaf = name.rsplit(".", 1)[-1]
        if kind == "buffer_long":
            v = np.asarray(100, dtype=np.int64)
        elif leaf == "running_mean":
            v = rs.normal(0.0, 0.1, shape)
        elif leaf == "running_var":
            v = rs.uniform(0.5, 1.5, shape)
        elif name == "rf_pre.0.weight":
            v = pre_w if kind == "buffer" else pre_w + rs.normal(0.0, 0.01, shape)
        elif name == "rf_post.0.weight":
            v = post_w if kind == "buffer" else post_w + rs.normal(0.0, 0.01, shape)
        elif leaf == "pe":
            v = sinusoid_pe(cfg.rf_channels, cfg.rf_freq) + rs.normal(0.0, 0.05, shape)
        elif name == "dec_post.3.bias":
            v = np.array([1.0, 0.0])
        elif name == "dec_post.3.scale":
            v = np.array([1.3])
        elif leaf == "original0":                       # weight-norm gain g
            v = None                                    # filled after its 'original1'
        elif leaf in ("weight", "original1", "weight_ih_l0", "weight_hh_l0") and len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            if name == "dec_post.3.weight":             # ConvTranspose1d: fan_in = out_ch * k
                fan_in = shape[1] * shape[2]
            b = 1.0 / np.sqrt(fan_in)
            v = rs.uniform(-b, b, shape)
        elif leaf == "weight":                          # BN gamma
            v = rs.uniform(0.5, 1.5, shape)
        elif leaf in ("bias", "bias_ih_l0", "bias_hh_l0"):
            v = rs.normal(0.0, 0.1, shape)
        else:
            raise AssertionError(f"no init rule for {name}")
        sd[name] = None if v is None else np.ascontiguousarray(v, dtype=np.float32 if kind != "buffer_long" else np.int64)
    for name in list(sd):
        if name.endswith("original0"):
            v = sd[name[:-1] + "1"]
            g = np.linalg.norm(v.astype(np.float64), axis=1, keepdims=True) * rs.uniform(0.8, 1.2, (v.shape[0], 1))
            sd[name] = g.astype(np.float32)
    return sd
