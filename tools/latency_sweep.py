#!/usr/bin/env python
"""BASELINE.json configs[4]: per-frame latency at batch 1 -- FastEnhancer T/B/S/M/L (16 kHz), one stream, one
fused-kernel launch per hop (the way a live caller uses fe_stream), p50 / p99 microseconds per hop from CUDA events,
beside the arithmetic floor of the hop.  usage: python tools/latency_sweep.py [n_hops] [precision]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastenhancer_b200.config import PRESETS  # noqa: E402
from fastenhancer_b200.engine import Engine  # noqa: E402
from fastenhancer_b200.fold import fold_to_canonical  # noqa: E402
from fastenhancer_b200.schema import synthetic_state_dict  # noqa: E402
from fastenhancer_b200.synth import synthetic_noisy  # noqa: E402


def main():
    n_hops = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    precision = sys.argv[2] if len(sys.argv) > 2 else "tf32"
    warm = 200
    rows = []
    for name in ("16k_t", "16k_b", "16k_s", "16k_m", "16k_l", "48k_l"):
        cfg = PRESETS[name]
        eng = Engine(cfg, fold_to_canonical(cfg, synthetic_state_dict(cfg, 0)), "cuda:0", precision=precision)
        H = cfg.hop_size
        x = torch.from_numpy(synthetic_noisy(1, (n_hops + warm) * H, cfg.sample_rate)).cuda()
        y = torch.empty_like(x)
        st = eng.new_state(1)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_hops + 1)]
        for i in range(warm):
            eng.stream(st, x[:, i * H:(i + 1) * H], out=y[:, i * H:(i + 1) * H])
        torch.cuda.synchronize()
        evs[0].record()
        for i in range(n_hops):
            j = warm + i
            eng.stream(st, x[:, j * H:(j + 1) * H], out=y[:, j * H:(j + 1) * H])
            evs[i + 1].record()
        torch.cuda.synchronize()
        us = np.array([evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(n_hops)])
        # persistent variant: all hops in one launch (state never leaves the SM)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.stream(st, x[:, :n_hops * H], out=y[:, :n_hops * H]); e1.record(); torch.cuda.synchronize()
        per_hop_persistent = e0.elapsed_time(e1) * 1e3 / n_hops
        hop_us = H / cfg.sample_rate * 1e6
        floor_us = cfg.flops_per_frame() / (128 * 2 * 1.965e9) * 1e6     # one SM's fp32 FMA pipe
        row = {"preset": name, "precision": precision, "hops": n_hops, "p50_us": float(np.percentile(us, 50)),
               "p99_us": float(np.percentile(us, 99)), "mean_us": float(us.mean()), "persistent_us_per_hop": per_hop_persistent,
               "hop_duration_us": hop_us, "rtf_p50": float(np.percentile(us, 50)) / hop_us,
               "one_sm_fp32_floor_us": floor_us}
        rows.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
