"""Checkpoint ingestion without the reference classes, folded ON THE DEVICE.

Replaces, for the hot path, ``ModelWrapper.load()`` + ``remove_weight_reparameterizations()``
(/root/reference/wrappers/ns.py:308-321, /root/reference/models/fastenhancer/default/model.py:532-608): reads
``logs/<name>/NNNNN.pth`` (``ckpt['model']`` holds the pre-fold parameters under the reference's names) and
``logs/<name>/config.yaml`` (``model_kwargs``), moves the tensors to the GPU and applies one fold rule per canonical tensor with
``fe_fold_device`` (C ABI, include/fastenhancer_b200.h).  ``fastenhancer_b200.fold`` (numpy, float64) is the host oracle of the same
rules (tests/test_checkpoint.py).  PyTorch is used for file I/O and device memory only.
"""
from __future__ import annotations

import ctypes
import os
import re
import typing as tp

import numpy as np

from .config import FEConfig
from .schema import canonical_schema

FOLD_COPY, FOLD_WEIGHT_NORM, FOLD_BATCH_NORM, FOLD_FINAL_CONV = 0, 1, 2, 3


class CFoldOp(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("rows", ctypes.c_int), ("cols", ctypes.c_int), ("flag", ctypes.c_int)] + \
        [(n, ctypes.c_void_p) for n in ("w", "a", "b", "c", "d")] + \
        [("dst", ctypes.c_longlong), ("bias_dst", ctypes.c_longlong), ("eps", ctypes.c_float)]


def fold_rules(cfg: FEConfig, names: tp.Container[str]) -> tp.List[tp.Dict[str, tp.Any]]:
    """One rule per weight tensor of the canonical array: ``{'kind', 'dst', 'w', ['a', 'b', 'c', 'd', 'bias', 'eps', 'flag']}`` with
    state_dict names as sources and canonical tensor names as destinations."""
    if cfg.pre_norm:
        raise ValueError("pre_norm checkpoints are folded on the host (fastenhancer_b200.fold); no shipped config uses them")
    rules: tp.List[tp.Dict[str, tp.Any]] = []

    def bn(dst_w, dst_b, w, prefix, eps):
        rules.append(dict(kind=FOLD_BATCH_NORM, dst=dst_w, bias=dst_b, w=w, a=f"{prefix}.weight", b=f"{prefix}.bias",
                          c=f"{prefix}.running_mean", d=f"{prefix}.running_var", eps=eps))

    def wn(dst, base, leaf):
        if cfg.weight_norm and f"{base}.parametrizations.{leaf}.original0" in names:
            rules.append(dict(kind=FOLD_WEIGHT_NORM, dst=dst, w=f"{base}.parametrizations.{leaf}.original1",
                              a=f"{base}.parametrizations.{leaf}.original0"))
        else:
            rules.append(dict(kind=FOLD_COPY, dst=dst, w=f"{base}.{leaf}"))

    def copy(dst, src, zero_if_missing=False):
        rules.append(dict(kind=FOLD_COPY, dst=dst, w=src, optional=zero_if_missing))

    bn("enc_pre.w", "enc_pre.b", "enc_pre.0.weight", "enc_pre.1", cfg.bn_eps)
    for i in range(cfg.n_enc):
        bn(f"enc.{i}.w", f"enc.{i}.b", f"encoder.{i}.0.weight", f"encoder.{i}.1", cfg.bn_eps)
    copy("rf_pre.lin", "rf_pre.0.weight")
    bn("rf_pre.w", "rf_pre.b", "rf_pre.1.weight", "rf_pre.2", cfg.bn_eps)
    for k in range(cfg.rf_blocks):
        p = f"rf_block.{k}"
        wn(f"blk.{k}.w_ih", f"{p}.rnn", "weight_ih_l0")
        wn(f"blk.{k}.w_hh", f"{p}.rnn", "weight_hh_l0")
        copy(f"blk.{k}.b_ih", f"{p}.rnn.bias_ih_l0")
        copy(f"blk.{k}.b_hh", f"{p}.rnn.bias_hh_l0")
        bn(f"blk.{k}.rnn_fc.w", f"blk.{k}.rnn_fc.b", f"{p}.rnn_fc.weight", f"{p}.rnn_post_norm", cfg.rf_eps)
        if k == 0:
            copy("blk.0.pe", f"{p}.pe", zero_if_missing=True)
        wn(f"blk.{k}.qkv.w", f"{p}.attn.qkv", "weight")
        copy(f"blk.{k}.qkv.b", f"{p}.attn.qkv.bias", zero_if_missing=True)
        bn(f"blk.{k}.attn_fc.w", f"blk.{k}.attn_fc.b", f"{p}.attn_fc.weight", f"{p}.attn_post_norm", cfg.rf_eps)
    copy("rf_post.lin", "rf_post.0.weight")
    bn("rf_post.w", "rf_post.b", "rf_post.1.weight", "rf_post.2", cfg.bn_eps)
    for i in range(cfg.n_enc):
        bn(f"dec.{i}.w1", f"dec.{i}.b1", f"decoder.{i}.0.weight", f"decoder.{i}.1", cfg.bn_eps)
        bn(f"dec.{i}.w2", f"dec.{i}.b2", f"decoder.{i}.3.weight", f"decoder.{i}.4", cfg.bn_eps)
    bn("dec_post.w", "dec_post.b", "dec_post.0.weight", "dec_post.1", cfg.bn_eps)
    if "dec_post.3.scale" in names:
        rules.append(dict(kind=FOLD_FINAL_CONV, dst="dec_post.wt", w="dec_post.3.weight", a="dec_post.3.scale",
                          flag=int(cfg.normalize_final_conv)))
    else:
        copy("dec_post.wt", "dec_post.3.weight")
    copy("dec_post.bt", "dec_post.3.bias")
    return rules


def fold_on_device(cfg: FEConfig, state_dict: tp.Mapping[str, tp.Any], device) -> "tp.Any":
    """pre-fold reference state_dict -> canonical folded weight array, a float32 CUDA tensor, computed by fe_fold_device."""
    import torch
    from .engine import _check, _stream_ptr, load_library
    lib = load_library()
    device = torch.device(device)
    offsets, off = {}, 0
    for name, shape in canonical_schema(cfg):
        offsets[name] = (off, shape)
        off += int(np.prod(shape))
    canon = torch.zeros(off, dtype=torch.float32, device=device)       # optional tensors that are absent stay zero
    keep = []                                                           # device copies of the sources, alive until the launches ran

    def dev(name):
        t = state_dict[name]
        t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t))
        t = t.detach().to(device=device, dtype=torch.float32).contiguous()
        keep.append(t)
        return t

    ops = []
    for r in fold_rules(cfg, state_dict):
        if r.get("optional") and r["w"] not in state_dict:
            continue
        dst_off, shape = offsets[r["dst"]]
        w = dev(r["w"])
        if w.numel() != int(np.prod(shape)):
            raise ValueError(f"{r['w']}: {tuple(w.shape)} does not fold into {r['dst']} {shape}")
        rows = shape[0] if r["kind"] in (FOLD_WEIGHT_NORM, FOLD_BATCH_NORM) else 1
        op = CFoldOp(kind=r["kind"], rows=rows, cols=w.numel() // rows, flag=int(r.get("flag", 0)), w=w.data_ptr(),
                     dst=dst_off, bias_dst=-1, eps=float(r.get("eps", 0.0)))
        for key in ("a", "b", "c", "d"):
            if key in r:
                t = dev(r[key])
                if r["kind"] != FOLD_FINAL_CONV and t.numel() != rows:
                    raise ValueError(f"{r[key]}: {t.numel()} values for {rows} rows of {r['dst']}")
                setattr(op, key, t.data_ptr())
        if "bias" in r:
            op.bias_dst = offsets[r["bias"]][0]
        ops.append(op)
    arr = (CFoldOp * len(ops))(*ops)
    lib.fe_fold_device.argtypes = [ctypes.POINTER(CFoldOp), ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    with torch.cuda.device(device):
        _check(lib.fe_fold_device(arr, len(ops), canon.data_ptr(), _stream_ptr(device)), "fe_fold_device")
        torch.cuda.current_stream(device).synchronize()
    del keep
    return canon


def latest_checkpoint(base_dir: str) -> str:
    """``logs/<name>`` -> path of its newest ``NNNNN.pth`` (wrappers/ns.py:296-304)."""
    files = sorted(int(f[:-4]) for f in os.listdir(base_dir) if re.match(r"[0-9]{5,}\.pth$", f))
    if not files:
        raise FileNotFoundError(f"no NNNNN.pth checkpoint in {base_dir}")
    return os.path.join(base_dir, f"{files[-1]:05d}.pth")


def load_checkpoint(path: str, device="cpu"):
    """-> (FEConfig, state_dict of pre-fold tensors).  ``path`` is a ``.pth`` file or a ``logs/<name>`` directory; the model shape
    comes from the ``config.yaml`` beside it (utils/hparams.py:137-146 copies it there)."""
    import torch
    import yaml
    if os.path.isdir(path):
        path = latest_checkpoint(path)
    hps = yaml.safe_load(open(os.path.join(os.path.dirname(path), "config.yaml")))
    kw = dict(hps["model_kwargs"])
    sr = int(hps.get("data", {}).get("sampling_rate", 48_000 if kw.get("n_fft", 512) >= 1024 else 16_000))
    cfg = FEConfig.from_model_kwargs(kw, sample_rate=sr)
    cfg.validate()
    ckpt = torch.load(path, map_location=device)
    return cfg, ckpt["model"]
