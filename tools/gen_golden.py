#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE itself (imported from /root/reference).

Runs only on the build container (the reference does not travel to the GPU box).  For every
shipped preset it loads fastenhancer_b200.schema.synthetic_state_dict(seed=0) into the reference
``Model`` / ``ONNXModel`` (strict), feeds fastenhancer_b200.synth.synthetic_noisy and stores

* ``stream_out``  : streaming wav->wav, the composition of scripts/export_onnx.py:48-58 on the
                    folded ONNXModel, hop by hop (raw concatenated hops, not delay-trimmed);
* ``stream_state``: final [cache_stft | cache_istft | h_0..h_{K-1}] per stream;
* ``offline_wav`` / ``offline_spec``: ``Model.forward`` on an unfolded model (model.py:728-735);
* ``spec_in`` / ``spec_out`` / ``spec_h``: ``ONNXModel.forward`` with T=3 frames per call, two calls;
* ``tap.*``       : per-module outputs (forward hooks) of one streaming frame, stream 0;
* ``canonical_sha``/``canonical_head``: the reference's own folded weights in canonical order.

Usage: python tools/gen_golden.py [preset ...]
"""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastenhancer_b200.config import PRESETS  # noqa: E402
from fastenhancer_b200.schema import synthetic_state_dict, canonical_schema  # noqa: E402
from fastenhancer_b200.synth import synthetic_noisy  # noqa: E402
from fastenhancer_b200.fold import fold_to_canonical  # noqa: E402

N_STREAMS, N_HOPS, TAP_HOP = 2, 24, 5
FULL = ("16k_t", "16k_b")          # presets that keep every per-layer tap


def import_reference():
    stub = tempfile.mkdtemp(prefix="librosa_stub_")
    os.makedirs(os.path.join(stub, "librosa"))
    open(os.path.join(stub, "librosa", "__init__.py"), "w").write("from . import filters\n")
    open(os.path.join(stub, "librosa", "filters.py"), "w").write("def mel(*a, **k):\n    raise NotImplementedError\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, "/root/reference")
    from models.fastenhancer.default.model import Model, ONNXModel
    return Model, ONNXModel


def reference_canonical(cfg, m) -> np.ndarray:
    """folded reference module -> canonical flat array (names after folding: model.py:559-608)."""
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    t = {}
    t["enc_pre.w"], t["enc_pre.b"] = sd["enc_pre.0.weight"], sd["enc_pre.0.bias"]
    for i in range(cfg.n_enc):
        t[f"enc.{i}.w"], t[f"enc.{i}.b"] = sd[f"encoder.{i}.0.weight"], sd[f"encoder.{i}.0.bias"]
    t["rf_pre.lin"], t["rf_pre.w"], t["rf_pre.b"] = sd["rf_pre.0.weight"], sd["rf_pre.1.weight"][:, :, 0], sd["rf_pre.1.bias"]
    for k in range(cfg.rf_blocks):
        p = f"rf_block.{k}"
        t[f"blk.{k}.w_ih"], t[f"blk.{k}.w_hh"] = sd[f"{p}.rnn.weight_ih_l0"], sd[f"{p}.rnn.weight_hh_l0"]
        t[f"blk.{k}.b_ih"], t[f"blk.{k}.b_hh"] = sd[f"{p}.rnn.bias_ih_l0"], sd[f"{p}.rnn.bias_hh_l0"]
        t[f"blk.{k}.rnn_fc.w"], t[f"blk.{k}.rnn_fc.b"] = sd[f"{p}.rnn_fc.weight"], sd[f"{p}.rnn_fc.bias"]
        if k == 0:
            t["blk.0.pe"] = sd[f"{p}.pe"]
        t[f"blk.{k}.qkv.w"] = sd[f"{p}.attn.qkv.weight"]
        t[f"blk.{k}.qkv.b"] = sd.get(f"{p}.attn.qkv.bias", np.zeros(3 * cfg.rf_channels, np.float32))
        t[f"blk.{k}.attn_fc.w"], t[f"blk.{k}.attn_fc.b"] = sd[f"{p}.attn_fc.weight"], sd[f"{p}.attn_fc.bias"]
    t["rf_post.lin"], t["rf_post.w"], t["rf_post.b"] = sd["rf_post.0.weight"], sd["rf_post.1.weight"][:, :, 0], sd["rf_post.1.bias"]
    for i in range(cfg.n_enc):
        t[f"dec.{i}.w1"], t[f"dec.{i}.b1"] = sd[f"decoder.{i}.0.weight"][:, :, 0], sd[f"decoder.{i}.0.bias"]
        t[f"dec.{i}.w2"], t[f"dec.{i}.b2"] = sd[f"decoder.{i}.2.weight"], sd[f"decoder.{i}.2.bias"]
    t["dec_post.w"], t["dec_post.b"] = sd["dec_post.0.weight"][:, :, 0], sd["dec_post.0.bias"]
    t["dec_post.wt"], t["dec_post.bt"] = sd["dec_post.2.weight"], sd["dec_post.2.bias"]
    return np.concatenate([np.asarray(t[n], np.float32).reshape(-1) for n, _ in canonical_schema(cfg)])


LONG_HOPS = {"16k_b": 626, "16k_m": 1003}      # BASELINE configs 2 / 3 at their real horizon (10 s utterances)
LONG_KEEP = 50                                  # hops of the tail kept in the fixture


def gen_long(ONNXModel, name):
    """tests/golden/<name>_long.npz: the reference's streaming graph over the whole 10 s utterance, one stream; keeps the last
    LONG_KEEP hops of the waveform and the final caches (a long-recurrence pin for the oracle and the kernels)."""
    cfg = PRESETS[name]
    n_hops = LONG_HOPS[name]
    H, N, K, F2, C2 = cfg.hop_size, cfg.n_fft, cfg.rf_blocks, cfg.rf_freq, cfg.rf_channels
    sd = {k: torch.from_numpy(np.array(v)) for k, v in synthetic_state_dict(cfg, seed=0).items()}
    with torch.no_grad():
        m = ONNXModel(**cfg.to_model_kwargs()).eval()
        m.load_state_dict(sd, strict=True)
        m.remove_weight_reparameterizations()
        x = torch.from_numpy(synthetic_noisy(1, n_hops * H, cfg.sample_rate, first_stream=5))
        c_stft, c_istft = torch.zeros(1, N - H), torch.zeros(1, N - H)
        hs = [torch.zeros(1, F2, C2) for _ in range(K)]
        hops = []
        for i in range(n_hops):
            spec_in, c_stft = m.stft(x[:, i * H:(i + 1) * H], c_stft)
            spec_out, *hs = m(spec_in, *hs)
            y, c_istft = m.stft.inverse(spec_out, c_istft)
            hops.append(y.clone())
    out = {"n_hops": np.int64(n_hops), "keep_hops": np.int64(LONG_KEEP),
           "stream_tail": torch.cat(hops[-LONG_KEEP:], dim=1).numpy(),
           "stream_state": torch.cat([c_stft, c_istft] + [h.view(1, F2 * C2) for h in hs], dim=1).numpy()}
    path = os.path.join(ROOT, "tests", "golden", f"{name}_long.npz")
    np.savez_compressed(path, torch_version=np.array(torch.__version__), **out)
    print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), {n_hops} hops, tail rms {np.sqrt((out['stream_tail'] ** 2).mean()):.4f}")


def main():
    Model, ONNXModel = import_reference()
    torch.set_num_threads(1)
    if "--long" in sys.argv:
        for name in [a for a in sys.argv[1:] if a != "--long"] or sorted(LONG_HOPS):
            gen_long(ONNXModel, name)
        return
    names = sys.argv[1:] or sorted(PRESETS)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in names:
        cfg = PRESETS[name]
        kw = cfg.to_model_kwargs()
        sd = {k: torch.from_numpy(np.array(v)) for k, v in synthetic_state_dict(cfg, seed=0).items()}
        H, N, K, F2, C2 = cfg.hop_size, cfg.n_fft, cfg.rf_blocks, cfg.rf_freq, cfg.rf_channels
        out = {}
        with torch.no_grad():
            # ---------------- offline, unfolded (what scripts/test_pytorch.py runs) -----------
            off = Model(**kw).eval()
            off.load_state_dict(sd, strict=True)
            L = 20 * H + 37
            wav = torch.from_numpy(synthetic_noisy(N_STREAMS, L, cfg.sample_rate))
            wav_hat, spec_hat = off(wav)
            out["offline_len"] = np.int64(L)
            out["offline_wav"] = wav_hat.numpy()
            if name in FULL:
                out["offline_spec"] = spec_hat.numpy()
            else:                                   # keep the fixture small: 3 frames only
                out["offline_spec_frames"] = np.array([0, 10, 20])
                out["offline_spec"] = spec_hat[:, :, [0, 10, 20]].numpy()
            # ---------------- streaming, folded ----------------------------------------------
            m = ONNXModel(**kw).eval()
            m.load_state_dict(sd, strict=True)
            m.remove_weight_reparameterizations()
            canon_ref = reference_canonical(cfg, m)
            canon_mine = fold_to_canonical(cfg, synthetic_state_dict(cfg, seed=0))
            err = np.abs(canon_ref - canon_mine).max()
            assert err < 1e-6, f"{name}: fold mismatch {err}"
            out["canonical_sha"] = np.frombuffer(hashlib.sha256(canon_ref.tobytes()).digest(), np.uint8)
            out["canonical_head"] = canon_ref[:4096].copy()
            x = torch.from_numpy(synthetic_noisy(N_STREAMS, N_HOPS * H, cfg.sample_rate))
            c_stft = torch.zeros(N_STREAMS, N - H)
            c_istft = torch.zeros(N_STREAMS, N - H)
            hs = [torch.zeros(1, N_STREAMS * F2, C2) for _ in range(K)]
            taps = {}
            hooks = []

            def grab(key, fn=lambda o: o):
                def hook(_mod, _inp, o):
                    taps[key] = fn(o).detach().clone()
                return hook
            hops = []
            for i in range(N_HOPS):
                if i == TAP_HOP:
                    hooks.append(m.enc_pre.register_forward_hook(grab("enc_pre")))
                    for j, mod in enumerate(m.encoder):
                        hooks.append(mod.register_forward_hook(grab(f"enc.{j}")))
                    hooks.append(m.rf_pre.register_forward_hook(grab("rf_pre")))
                    for j, mod in enumerate(m.rf_block):
                        hooks.append(mod.register_forward_hook(grab(f"blk.{j}.out", lambda o: o[0])))
                    hooks.append(m.rf_post.register_forward_hook(grab("rf_post")))
                    for j, mod in enumerate(m.decoder):
                        hooks.append(mod.register_forward_hook(grab(f"dec.{j}")))
                    hooks.append(m.dec_post.register_forward_hook(grab("mask")))
                spec_in, c_stft = m.stft(x[:, i * H:(i + 1) * H], c_stft)
                spec_out, *hs = m(spec_in, *hs)
                y, c_istft = m.stft.inverse(spec_out, c_istft)
                hops.append(y.clone())
                if i == TAP_HOP:
                    for hk in hooks:
                        hk.remove()
                    taps["spec_in"] = spec_in.clone()
                    taps["spec_out"] = spec_out.clone()
            out["stream_out"] = torch.cat(hops, dim=1).numpy()
            out["stream_state"] = torch.cat(
                [c_stft, c_istft] + [h.view(N_STREAMS, F2 * C2) for h in hs], dim=1).numpy()
            # taps, stream 0 only.  conv outputs are [B, C, F]; rf_pre is [B, C2, F2] (pre-permute),
            # block outputs are [T=1, B, F2, C2].
            if name in FULL:
                out["tap.enc_pre"] = taps["enc_pre"][0].numpy()
                out["tap.rf_post"] = taps["rf_post"][0].numpy()
                for j in range(cfg.n_enc):
                    out[f"tap.enc.{j}"] = taps[f"enc.{j}"][0].numpy()
                    out[f"tap.dec.{j}"] = taps[f"dec.{j}"][0].numpy()
            out["tap.rf_pre"] = taps["rf_pre"][0].t().contiguous().numpy()
            for j in range(K):
                out[f"tap.blk.{j}.out"] = taps[f"blk.{j}.out"][0, 0].numpy()
            out["tap.mask"] = taps["mask"][0].numpy()
            out["tap.spec_in"] = taps["spec_in"][0, :, 0].numpy()
            out["tap.spec_out"] = taps["spec_out"][0, :, 0].numpy()
            # ---------------- spec2spec, T=3 frames per call, two calls ----------------------
            frames = torch.stft(torch.from_numpy(synthetic_noisy(N_STREAMS, 8 * H, cfg.sample_rate)), N, hop_length=H,
                                win_length=N, window=torch.hann_window(N), center=True, return_complex=True)
            spec = torch.view_as_real(frames)[:, :, :6].contiguous()
            hs = [torch.zeros(1, N_STREAMS * F2, C2) for _ in range(K)]
            o1, *hs = m(spec[:, :, :3], *hs)
            o2, *hs = m(spec[:, :, 3:6], *hs)
            out["spec_in"] = spec.numpy()
            out["spec_out"] = torch.cat([o1, o2], dim=2).numpy()
            out["spec_h"] = torch.stack([h.view(N_STREAMS, F2, C2) for h in hs], dim=1).numpy()
        path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, torch_version=np.array(torch.__version__), **out)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB), fold max|d|={err:.2e}, "
              f"stream rms={np.sqrt((out['stream_out'] ** 2).mean()):.4f}")


if __name__ == "__main__":
    main()
