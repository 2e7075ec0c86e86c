"""Seeded synthetic noisy speech-like input (SURVEY.md section 8(d) "Concrete inputs").

Per stream ``b``: an amplitude-modulated tone ``0.1*sin(2*pi*f_b*t)*(1+0.5*sin(2*pi*3*t))`` with
``f_b = 200 + 37*(b mod 64)`` Hz plus ``0.05*N(0,1)`` noise from ``RandomState(1234+b)``, clamped to
[-1, 1] like the reference scripts do (/root/reference/scripts/export_onnx.py:85).
numpy's RandomState stream is frozen, so tests on any box regenerate bit-identical inputs.
"""
from __future__ import annotations

import numpy as np


def synthetic_noisy(n_streams: int, n_samples: int, sample_rate: int = 16_000, first_stream: int = 0) -> np.ndarray:
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    out = np.empty((n_streams, n_samples), np.float32)
    for i in range(n_streams):
        b = first_stream + i
        f = 200.0 + 37.0 * (b % 64)
        clean = 0.1 * np.sin(2 * np.pi * f * t) * (1.0 + 0.5 * np.sin(2 * np.pi * 3.0 * t))
        noise = 0.05 * np.random.RandomState(1234 + b).standard_normal(n_samples)
        out[i] = np.clip(clean + noise, -1.0, 1.0).astype(np.float32)
    return out
