"""The audio front door on the device: WAV bytes -> mono float32 at the model's rate -> enhanced -> WAV bytes.

GPU-resident replacement of the per-file body of the reference's directory runner
(/root/reference/scripts/test_pytorch.py:27-37: ``librosa.load(path, sr=wrapper.sr, mono=True)`` -> model -> ``sf.write``):
the host only reads / writes file bytes; sample-format conversion, mono mix-down and polyphase resampling run as CUDA kernels
through the C ABI (fe_pcm16_to_float / fe_resample_poly / fe_float_to_pcm16, fastenhancer_b200/csrc/fe_frontdoor.cu).

Resampling follows ``scipy.signal.resample_poly`` (Kaiser-windowed FIR, beta 5) -- the oracle in tests/test_frontdoor.py.  librosa's
own resampler (soxr) is not in this image; files already at the model's rate, like the reference's fixtures, are unaffected.
"""
from __future__ import annotations

import os
import typing as tp
import wave
from math import gcd

import numpy as np

from .engine import _check, _stream_ptr, load_library


def resample_taps(up: int, down: int, beta: float = 5.0) -> np.ndarray:
    """FIR of scipy.signal.resample_poly(window=('kaiser', beta)): firwin(2*half_len+1, 1/max(up, down)) * up, half_len = 10*max."""
    max_rate = max(up, down)
    half_len = 10 * max_rate
    n = 2 * half_len + 1
    m = np.arange(n, dtype=np.float64) - half_len
    cutoff = 1.0 / max_rate                                   # normalised to Nyquist = 1
    h = cutoff * np.sinc(cutoff * m) * np.kaiser(n, beta)
    h /= h.sum()                                              # firwin scales the pass band to unit gain at DC
    return np.ascontiguousarray(up * h, dtype=np.float32)


def read_wav_pcm16(path: str) -> tp.Tuple[np.ndarray, int]:
    """-> (int16 array [n_frames, n_channels], sample rate); 16-bit PCM WAV only (the format of the reference's fixtures)."""
    with wave.open(str(path), "rb") as w:
        if w.getsampwidth() != 2 or w.getcomptype() != "NONE":
            raise ValueError(f"{path}: only 16-bit PCM WAV is supported by the device front door")
        fs, ch, n = w.getframerate(), w.getnchannels(), w.getnframes()
        pcm = np.frombuffer(w.readframes(n), dtype="<i2").reshape(-1, ch)
    return np.ascontiguousarray(pcm), fs


def load_wav(path: str, sr: int, device) -> "tp.Tuple[tp.Any, int]":
    """WAV file -> (float32 CUDA tensor [1, L] at rate ``sr``, the file's own rate); conversion and resampling on the device."""
    import torch
    lib = load_library()
    device = torch.device(device)
    pcm, fs = read_wav_pcm16(path)
    n = pcm.shape[0]
    with torch.cuda.device(device):
        pcm_d = torch.from_numpy(pcm).to(device, non_blocking=True)
        wav = torch.empty(n, dtype=torch.float32, device=device)
        _check(lib.fe_pcm16_to_float(pcm_d.data_ptr(), n, pcm.shape[1], wav.data_ptr(), _stream_ptr(device)), "fe_pcm16_to_float")
        if sr != fs:
            g = gcd(int(sr), int(fs))
            up, down = int(sr) // g, int(fs) // g
            taps = torch.from_numpy(resample_taps(up, down)).to(device)
            n_out = -(-n * up // down)                         # ceil(n * up / down), as resample_poly
            out = torch.empty(n_out, dtype=torch.float32, device=device)
            _check(lib.fe_resample_poly(wav.data_ptr(), n, up, down, taps.data_ptr(), taps.numel(), out.data_ptr(), n_out,
                                        _stream_ptr(device)), "fe_resample_poly")
            wav = out
    return wav.unsqueeze(0), fs


def save_wav(path: str, wav, sr: int) -> None:
    """float32 CUDA tensor [L] or [1, L] -> 16-bit PCM WAV; the conversion runs on the device."""
    import torch
    lib = load_library()
    x = wav.reshape(-1).contiguous().to(torch.float32)
    with torch.cuda.device(x.device):
        pcm = torch.empty(x.numel(), dtype=torch.int16, device=x.device)
        _check(lib.fe_float_to_pcm16(x.data_ptr(), x.numel(), pcm.data_ptr(), _stream_ptr(x.device)), "fe_float_to_pcm16")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(int(sr))
        w.writeframes(pcm.cpu().numpy().astype("<i2").tobytes())


def enhance_directory(model, input_dir: str, output_dir: str, sr: int, device="cuda:0") -> tp.List[str]:
    """The loop of scripts/test_pytorch.py:26-37 with the audio staying on the GPU between the file read and the file write."""
    import torch
    os.makedirs(output_dir, exist_ok=True)
    done = []
    for name in sorted(f for f in os.listdir(input_dir) if f.endswith(".wav")):
        noisy, _fs = load_wav(os.path.join(input_dir, name), sr, device)
        with torch.no_grad():
            enhanced, _ = model(noisy)
        save_wav(os.path.join(output_dir, name), enhanced, sr)
        done.append(name)
    return done
