"""Device front door (fastenhancer_b200.frontdoor): PCM16 -> float, polyphase resampling, float -> PCM16, against numpy / scipy.
Bit-exact for the integer <-> float conversions; resampling within 2e-6 of scipy.signal.resample_poly (fp32 accumulation order)."""
import os

import numpy as np
import pytest

from fastenhancer_b200.frontdoor import read_wav_pcm16, resample_taps

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("up,down", [(1, 3), (2, 3), (160, 441), (3, 1)])
def test_taps_equal_scipy_firwin(up, down):
    from scipy.signal import firwin
    max_rate = max(up, down)
    want = firwin(2 * 10 * max_rate + 1, 1.0 / max_rate, window=("kaiser", 5.0)) * up
    np.testing.assert_allclose(resample_taps(up, down), want, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("fs,sr,ch", [(48000, 16000, 1), (44100, 16000, 2), (16000, 16000, 1), (16000, 48000, 1)])
def test_load_wav_matches_scipy(tmp_path, fs, sr, ch):
    from math import gcd

    import torch
    from scipy.io import wavfile
    from scipy.signal import resample_poly

    from fastenhancer_b200.frontdoor import load_wav, save_wav
    rs = np.random.RandomState(fs + ch)
    n = fs // 3 + 17
    pcm = (rs.standard_normal((n, ch)) * 4000 + 8000 * np.sin(np.arange(n) * 0.05)[:, None]).clip(-32768, 32767).astype(np.int16)
    path = str(tmp_path / "in.wav")
    wavfile.write(path, fs, pcm if ch > 1 else pcm[:, 0])
    got_pcm, got_fs = read_wav_pcm16(path)
    assert got_fs == fs and np.array_equal(got_pcm, pcm)
    wav, file_fs = load_wav(path, sr, "cuda:0")
    mono = pcm.astype(np.float64).mean(axis=1) / 32768.0
    g = gcd(sr, fs)
    want = resample_poly(mono, sr // g, fs // g) if sr != fs else mono
    got = wav.cpu().numpy()[0]
    assert file_fs == fs and got.shape == want.shape
    if sr == fs:
        assert np.array_equal(got, want.astype(np.float32))                 # conversion alone is exact
    else:
        assert np.abs(got - want).max() < 2e-6
    out = str(tmp_path / "out.wav")
    save_wav(out, wav, sr)
    back_fs, back = wavfile.read(out)
    assert back_fs == sr and back.dtype == np.int16
    assert np.array_equal(back, np.clip(np.rint(got.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16))


@pytest.mark.gpu
def test_enhance_directory_equals_model_forward(tmp_path):
    """the GPU-resident directory loop == Model.forward on the same samples (file at the model's rate: nothing is resampled)"""
    import torch
    from scipy.io import wavfile

    from fastenhancer_b200.config import PRESETS
    from fastenhancer_b200.frontdoor import enhance_directory
    from fastenhancer_b200.model import Model
    from fastenhancer_b200.synth import synthetic_noisy
    cfg = PRESETS["16k_t"]
    x = synthetic_noisy(1, 16000 + 99, cfg.sample_rate)[0]
    pcm = np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16)
    (tmp_path / "in").mkdir()
    wavfile.write(str(tmp_path / "in" / "a.wav"), cfg.sample_rate, pcm)
    m = Model(**cfg.to_model_kwargs()).eval().cuda()
    assert enhance_directory(m, str(tmp_path / "in"), str(tmp_path / "out"), cfg.sample_rate) == ["a.wav"]
    want, _ = m(torch.from_numpy(pcm.astype(np.float32) / 32768.0).cuda()[None])
    _, got = wavfile.read(str(tmp_path / "out" / "a.wav"))
    assert np.array_equal(got, np.clip(np.rint(want.cpu().numpy()[0].astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16))
