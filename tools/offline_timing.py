#!/usr/bin/env python
"""Model.forward (fe_offline) on few long utterances: sequential walk vs frame-parallel schedule, CUDA-event timed.
usage: python tools/offline_timing.py [preset:B:seconds ...]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.engine import Engine
from fastenhancer_b200.schema import synthetic_state_dict
from fastenhancer_b200.fold import fold_to_canonical
from fastenhancer_b200.synth import synthetic_noisy


def canonical(cfg):
    return fold_to_canonical(cfg, synthetic_state_dict(cfg, 0))


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
    cases = sys.argv[1:] or ["16k_t:1:10", "16k_b:1:10", "16k_s:1:10", "16k_m:1:10", "16k_l:1:10", "48k_l:1:10", "16k_b:8:10", "16k_b:32:10"]
    for case in cases:
        name, B, sec = case.split(":")
        cfg, B, sec = PRESETS[name], int(B), float(sec)
        L = int(sec * cfg.sample_rate)
        x = torch.from_numpy(synthetic_noisy(B, L, cfg.sample_rate)).cuda()
        eng = Engine(cfg, canonical(cfg), "cuda:0")
        res = {}
        for mode in ("walk", "frame_parallel"):
            eng.set_offline_mode(mode)
            res[mode] = timed(lambda: eng.offline(x, want_spec=False))
        eng.set_offline_mode("walk"); a = eng.offline(x)[0]
        eng.set_offline_mode("frame_parallel"); b = eng.offline(x)[0]
        err = float((a - b).pow(2).mean().sqrt())
        T = 1 + L // cfg.hop_size
        print(f"OFFLINE {name} B={B} {sec:g}s ({T} frames) {eng.precision}: walk {res['walk']:.3f} ms, frame-parallel {res['frame_parallel']:.3f} ms "
              f"({res['walk'] / res['frame_parallel']:.1f}x), RTF {res['frame_parallel'] * 1e-3 / (B * sec):.6f}, "
              f"{B * T / res['frame_parallel'] * 1e3:.0f} frames/s, rms(walk - fp) {err:.2e}", flush=True)


if __name__ == "__main__":
    main()
