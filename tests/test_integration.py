"""The overlay package plugs into the reference's own model lookup (only where the reference tree is mounted:
the build container; skipped on the GPU box)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from fastenhancer_b200.config import PRESETS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not mounted")
def test_overlay_resolves_like_the_reference_does(tmp_path, monkeypatch):
    import yaml
    # `models` must resolve to our overlay package (a maintainer would drop it inside <reference>/models/)
    monkeypatch.syspath_prepend(os.path.join(ROOT, "integration"))
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        monkeypatch.delitem(sys.modules, k)
    hps = yaml.safe_load(open(os.path.join(REF, "configs", "fastenhancer", "b.yaml")))
    hps["model"] = "fastenhancer_b200"                                     # the only config change
    module = importlib.import_module(f"models.{hps['model']}.model")       # wrappers/ns.py:29-32
    model = module.Model(**hps["model_kwargs"])
    onnx_model = module.ONNXModel(**hps["model_kwargs"])                   # scripts/export_onnx.py:32-35
    assert model.cfg == PRESETS["16k_b"] and onnx_model.cfg == PRESETS["16k_b"]
    assert (model.stft.n_fft, model.stft.hop_size) == (512, 256)           # wrappers/ns.py:85-86


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not mounted")
def test_reference_checkpoint_loads_strict(tmp_path, monkeypatch):
    """state_dict of the reference's own Model -> our Model.load_state_dict(strict=True); folded weights agree."""
    stub = tmp_path / "librosa"
    stub.mkdir()
    (stub / "__init__.py").write_text("from . import filters\n")
    (stub / "filters.py").write_text("def mel(*a, **k):\n    raise NotImplementedError\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.syspath_prepend(REF)
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "functional" or k.startswith("functional.")]:
        monkeypatch.delitem(sys.modules, k)
    ref_mod = importlib.import_module("models.fastenhancer.default.model")
    cfg = PRESETS["16k_t"]
    torch.manual_seed(0)
    ref = ref_mod.ONNXModel(**cfg.to_model_kwargs()).eval()
    ck = {k: v.clone() for k, v in ref.state_dict().items()}
    from fastenhancer_b200.model import Model
    ours = Model(**cfg.to_model_kwargs())
    ours.load_state_dict(ck, strict=True)
    ref.remove_weight_reparameterizations()
    folded = ours.canonical_weights()
    # spot-check: first conv and the GRU of block 0, as folded by the reference itself
    sd = ref.state_dict()
    np.testing.assert_allclose(folded[:cfg.channels * 16], sd["enc_pre.0.weight"].numpy().reshape(-1), atol=2e-7)
    from fastenhancer_b200.schema import canonical_schema, split_canonical
    parts = split_canonical(cfg, folded)
    np.testing.assert_allclose(parts["blk.0.w_hh"], sd["rf_block.0.rnn.weight_hh_l0"].numpy(), atol=2e-7)
    np.testing.assert_allclose(parts["dec_post.wt"], sd["dec_post.2.weight"].numpy(), atol=2e-7)
    assert [n for n, _ in canonical_schema(cfg)][0] == "enc_pre.w"
