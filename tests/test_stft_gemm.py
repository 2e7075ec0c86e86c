"""The tensor-core (DFT-as-GEMM) STFT operator, fe_stft_gemm, against numpy's rfft -- the reference's ConvSTFT front end
(models/fastenhancer/conv_stft/model.py:55-63, 110-114: F.conv1d with the windowed DFT basis, stride = hop) -- and against the
engine's own FFT operator (fe_stft).  Tolerances, relative to the largest bin magnitude: fp32-accurate mode (3xTF32) 1e-5 max / 3e-6 RMS
-- the operands carry 22 significand bits, what is left is the tensor core's truncating fp32 accumulation over the 192 MMAs of a
512-sample contraction (measured 5.6e-6 max) --, single-pass TF32 2e-3."""
import numpy as np
import pytest
import torch

from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.synth import synthetic_noisy

pytestmark = pytest.mark.gpu


def _ref(cfg, x, T):
    N, H = cfg.n_fft, cfg.hop_size
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N) / N)
    fr = np.stack([x[:, t * H:t * H + N].astype(np.float64) * w for t in range(T)], axis=2)       # [B, N, T]
    return np.fft.rfft(fr, axis=1)                                                                 # [B, N/2+1, T]


@pytest.mark.parametrize("name,B,T", [("16k_b", 3, 300), ("16k_m", 2, 129), ("16k_l", 1, 128), ("48k_l", 2, 77), ("48k_t", 1, 1), ("16k_t", 5, 626)])
@pytest.mark.parametrize("accurate", [True, False])
def test_stft_gemm_matches_rfft(name, B, T, accurate, canonical):
    from fastenhancer_b200.engine import Engine
    cfg = PRESETS[name]
    eng = Engine(cfg, canonical(name), "cuda:0")
    N, H = cfg.n_fft, cfg.hop_size
    L = (T - 1) * H + N + 3                                  # ragged tail: not every sample belongs to a frame
    x = synthetic_noisy(B, L, cfg.sample_rate)
    spec = eng.stft_gemm(torch.from_numpy(x).cuda(), n_frames=T, accurate=accurate).cpu().numpy()
    ref = _ref(cfg, x, T)
    assert spec.shape == (B, N // 2 + 1, T, 2)
    got = spec[..., 0] + 1j * spec[..., 1]
    tol = 1e-5 if accurate else 2e-3
    assert np.abs(got - ref).max() < tol * np.abs(ref).max()
    assert np.sqrt(np.mean(np.abs(got - ref) ** 2)) < 0.3 * tol * np.abs(ref).max()
    assert np.all(spec[:, 0, :, 1] == 0) and np.all(spec[:, -1, :, 1] == 0)           # DC and Nyquist are real


def test_stft_gemm_equals_streaming_stft(canonical):
    """Same frames as the streaming operator fe_stft (cache of n_fft - hop zeros in front)."""
    from fastenhancer_b200.engine import Engine
    cfg = PRESETS["16k_b"]
    eng = Engine(cfg, canonical("16k_b"), "cuda:0")
    N, H, B, nh = cfg.n_fft, cfg.hop_size, 4, 40
    x = torch.from_numpy(synthetic_noisy(B, nh * H, cfg.sample_rate)).cuda()
    a = eng.stft(eng.new_state(B), x)
    b = eng.stft_gemm(torch.cat([torch.zeros(B, N - H, device="cuda"), x], dim=1))
    assert a.shape == b.shape and (a - b).abs().max() < 1e-5 * a.abs().max()


def test_stft_gemm_rejects_short_input(canonical):
    from fastenhancer_b200.engine import Engine
    cfg = PRESETS["16k_b"]
    eng = Engine(cfg, canonical("16k_b"), "cuda:0")
    with pytest.raises(ValueError):
        eng.stft_gemm(torch.zeros(1, cfg.n_fft - 1).cuda())
