#!/usr/bin/env python
"""STFT of a whole batch: the in-kernel packed-real radix-4 FFT operator (fe_stft, one CTA per group of streams walking the hops) vs the
tensor-core DFT-as-GEMM operator (fe_stft_gemm, 3xTF32 and single TF32) vs torch.stft (cuFFT), CUDA-event timed."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.engine import Engine
from fastenhancer_b200.fold import fold_to_canonical
from fastenhancer_b200.schema import synthetic_state_dict
from fastenhancer_b200.synth import synthetic_noisy


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    for case in sys.argv[1:] or ["16k_b:256:626", "16k_b:1:626", "16k_m:512:1003", "48k_l:256:2405"]:
        name, B, T = case.split(":")
        cfg, B, T = PRESETS[name], int(B), int(T)
        N, H = cfg.n_fft, cfg.hop_size
        eng = Engine(cfg, fold_to_canonical(cfg, synthetic_state_dict(cfg, 0)), "cuda:0")
        x = torch.from_numpy(synthetic_noisy(B, T * H, cfg.sample_rate)).cuda()
        xp = torch.cat([torch.zeros(B, N - H, device="cuda"), x], dim=1).contiguous()
        st = eng.new_state(B)
        win = torch.hann_window(N, periodic=True, device="cuda")
        t_fft = timed(lambda: eng.stft(st, x))
        t_x3 = timed(lambda: eng.stft_gemm(xp, n_frames=T, accurate=True))
        t_tf = timed(lambda: eng.stft_gemm(xp, n_frames=T, accurate=False))
        t_cu = timed(lambda: torch.view_as_real(torch.stft(xp, N, H, N, win, center=False, return_complex=True)))
        flop = 2.0 * B * T * N * N
        print(f"STFT {name} B={B} T={T} (N={N}, hop={H}): in-kernel FFT operator {t_fft:.3f} ms | GEMM 3xTF32 {t_x3:.3f} ms ({3 * flop / t_x3 * 1e-9:.0f} TFLOP/s executed) | "
              f"GEMM TF32 {t_tf:.3f} ms ({flop / t_tf * 1e-9:.0f} TFLOP/s) | torch.stft (cuFFT) {t_cu:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
