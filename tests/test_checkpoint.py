"""Checkpoint ingestion (fastenhancer_b200.checkpoint): `logs/<name>/NNNNN.pth` + `config.yaml` -> canonical weights, folded on the
device by fe_fold_device.  Host oracle of the fold rules: fastenhancer_b200.fold (itself pinned against the reference's own
remove_weight_reparameterizations in tests/test_oracle.py::test_fold_matches_reference)."""
import os

import numpy as np
import pytest
import torch

from fastenhancer_b200.checkpoint import (FOLD_BATCH_NORM, FOLD_COPY, FOLD_FINAL_CONV, FOLD_WEIGHT_NORM, fold_rules, latest_checkpoint,
                                          load_checkpoint)
from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.fold import fold_state_dict, fold_to_canonical
from fastenhancer_b200.schema import canonical_schema, synthetic_state_dict
from fastenhancer_b200.synth import synthetic_noisy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_logs(tmp_path, name):
    import yaml
    cfg = PRESETS[name]
    d = tmp_path / "logs" / name
    d.mkdir(parents=True)
    kw = cfg.to_model_kwargs()
    yaml.safe_dump({"model": "fastenhancer.default", "model_kwargs": kw, "data": {"sampling_rate": cfg.sample_rate}}, open(d / "config.yaml", "w"))
    sd = {k: torch.from_numpy(np.array(v)) for k, v in synthetic_state_dict(cfg, seed=3).items()}
    torch.save({"model": sd, "epoch": 7}, d / "00007.pth")
    torch.save({"model": {k: v * 0 for k, v in sd.items()}, "epoch": 2}, d / "00002.pth")       # an older one that must not be picked
    return str(d), sd


@pytest.mark.parametrize("name", sorted(PRESETS))
def test_rule_table_matches_host_fold(name):
    """every canonical tensor has exactly one rule, and the rules -- evaluated with numpy the way fold_kernel evaluates them --
    reproduce fastenhancer_b200.fold"""
    cfg = PRESETS[name]
    sd = synthetic_state_dict(cfg, seed=1)
    rules = fold_rules(cfg, sd)
    want = fold_state_dict(cfg, sd)
    produced = set()
    for r in rules:
        if r.get("optional") and r["w"] not in sd:              # absent optional tensors (qkv bias with attn_bias: False) fold to zeros
            assert not np.any(want[r["dst"]])
            produced.add(r["dst"])
            continue
        w = np.asarray(sd[r["w"]], np.float64)
        rows = w.shape[0]
        if r["kind"] == FOLD_COPY:
            out = w
        elif r["kind"] == FOLD_WEIGHT_NORM:
            g = np.asarray(sd[r["a"]], np.float64).reshape(rows)
            out = w * (g / np.sqrt((w.reshape(rows, -1) ** 2).sum(1))).reshape((rows,) + (1,) * (w.ndim - 1))
        elif r["kind"] == FOLD_BATCH_NORM:
            f = np.asarray(sd[r["a"]], np.float64) / np.sqrt(np.asarray(sd[r["d"]], np.float64) + r["eps"])
            out = w * f.reshape((rows,) + (1,) * (w.ndim - 1))
            np.testing.assert_allclose(np.asarray(sd[r["b"]], np.float64) - np.asarray(sd[r["c"]], np.float64) * f, want[r["bias"]], atol=1e-6)
            produced.add(r["bias"])
        else:
            assert r["kind"] == FOLD_FINAL_CONV
            f = float(np.asarray(sd[r["a"]]).reshape(()))
            out = w * (f / max(np.sqrt((w ** 2).sum()), 1e-12) if r["flag"] else f)
        np.testing.assert_allclose(out.reshape(want[r["dst"]].shape), want[r["dst"]], atol=1e-6, err_msg=r["dst"])
        assert r["dst"] not in produced
        produced.add(r["dst"])
    assert produced == {n for n, _ in canonical_schema(cfg)}


def test_logs_dir_layout(tmp_path):
    d, sd = _write_logs(tmp_path, "16k_t")
    assert latest_checkpoint(d).endswith("00007.pth")                         # newest numeric file (wrappers/ns.py:296-304)
    cfg, loaded = load_checkpoint(d)
    assert cfg == PRESETS["16k_t"] and set(loaded) == set(sd)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["16k_t", "16k_b", "16k_m", "48k_l"])
def test_device_fold_matches_host_fold(name):
    from fastenhancer_b200.checkpoint import fold_on_device
    cfg = PRESETS[name]
    sd = synthetic_state_dict(cfg, seed=2)
    got = fold_on_device(cfg, sd, "cuda:0").cpu().numpy()
    want = fold_to_canonical(cfg, sd)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-7


@pytest.mark.gpu
def test_engine_from_checkpoint(tmp_path):
    """Engine.from_checkpoint(logs/<name>) == Engine(host-folded weights): same kernels, same blob up to the last fold bit"""
    from fastenhancer_b200.engine import Engine
    d, sd = _write_logs(tmp_path, "16k_b")
    cfg = PRESETS["16k_b"]
    a = Engine.from_checkpoint(d, "cuda:0")
    b = Engine(cfg, fold_to_canonical(cfg, {k: v.numpy() for k, v in sd.items()}), "cuda:0")
    x = torch.from_numpy(synthetic_noisy(3, 10 * cfg.hop_size, cfg.sample_rate)).cuda()
    ya, yb = a.stream(a.new_state(3), x), b.stream(b.new_state(3), x)
    assert a.precision == b.precision == "fp32x3"
    assert float((ya - yb).abs().max()) < 1e-6 and float(ya.abs().max()) > 0.01
