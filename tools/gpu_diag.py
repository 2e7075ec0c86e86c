#!/usr/bin/env python
"""GPU bring-up diagnostics: per-stage parity of the fused kernel against the CPU oracle.

usage: python tools/gpu_diag.py PRESET S [n_streams] [n_hops]      (one config per process)
       python tools/gpu_diag.py --time PRESET n_streams n_hops [S]  (kernel timing, CUDA events)
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastenhancer_b200.config import PRESETS  # noqa: E402
from fastenhancer_b200.engine import Engine  # noqa: E402
from fastenhancer_b200.fold import fold_to_canonical  # noqa: E402
from fastenhancer_b200.schema import synthetic_state_dict  # noqa: E402
from fastenhancer_b200.synth import synthetic_noisy  # noqa: E402


def diag(name, S, B=5, nh=6):
    from oracle.oracle import Oracle, tap_schema
    cfg = PRESETS[name]
    canon = fold_to_canonical(cfg, synthetic_state_dict(cfg, 0))
    o = Oracle(cfg, canon)
    H = cfg.hop_size
    x = synthetic_noisy(B, nh * H, cfg.sample_rate)
    st = o.new_state(B)
    y_ref, taps_ref = o.stream(st, x, taps=True)
    eng = Engine(cfg, canon)
    eng.set_streams_per_cta(S)
    state = eng.new_state(B)
    hop = min(3, nh - 1)
    y, taps = eng.stream_taps(state, torch.from_numpy(x).cuda(), hop)
    torch.cuda.synchronize()
    y = y.cpu().numpy(); taps = taps.cpu().numpy()
    off, worst = 0, 0.0
    for nm, shp in tap_schema(cfg):
        n = int(np.prod(shp)); a = taps[off:off + n].reshape(shp); r = taps_ref[nm][hop]; off += n
        d = float(np.abs(a - r).max()) if np.isfinite(a).all() else float("nan")
        rel = d / max(1e-12, float(np.abs(r).max()))
        worst = max(worst, rel) if rel == rel else float("nan")
        print(f"  {nm:14s} max|d|={d:.3e} rel={rel:.2e}")
    e_w = float(np.abs(y - y_ref).max())
    e_s = float(np.abs(state.export().cpu().numpy() - st).max())
    rms = float(np.sqrt(np.mean((y - y_ref) ** 2)))
    print(f"DIAG {name} S={S} {eng.precision}: wav max|d|={e_w:.3e} rms={rms:.3e} state max|d|={e_s:.3e} worst tap rel={worst:.2e} "
          f"{'OK' if rms < (1e-5 if eng.precision == 'fp32' else 5e-5) and e_s < (5e-5 if eng.precision == 'fp32' else 2e-3) else 'FAIL'}")


def timing(name, B, nh, S=0):
    cfg = PRESETS[name]
    canon = fold_to_canonical(cfg, synthetic_state_dict(cfg, 0))
    eng = Engine(cfg, canon)
    if S:
        eng.set_streams_per_cta(S)
    H = cfg.hop_size
    x = torch.randn(B, nh * H, device="cuda") * 0.1
    y = torch.empty_like(x)
    state = eng.new_state(B)
    for _ in range(2):
        eng.stream(state, x, out=y)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    reps = 3
    for _ in range(reps):
        eng.stream(state, x, out=y)
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    fps = B * nh / (ms * 1e-3)
    tf = fps * cfg.flops_per_frame() / 1e12
    print(f"TIME {name} {eng.precision} B={B} hops={nh} S={eng.streams_per_cta(B)}: {ms:.3f} ms/launch, {ms * 1e3 / nh:.2f} us/hop, "
          f"{fps / 1e6:.3f} Mframes/s, {tf:.2f} TFLOP/s alg, RTF/stream={ms * 1e-3 / (nh * H / cfg.sample_rate):.5f}")


def profile(name, B, nh, S=0):
    """per-phase SM cycles of CTA 0 (clock64 stamps after each barrier-separated phase)."""
    cfg = PRESETS[name]
    canon = fold_to_canonical(cfg, synthetic_state_dict(cfg, 0))
    eng = Engine(cfg, canon)
    if S:
        eng.set_streams_per_cta(S)
    x = torch.randn(B, nh * cfg.hop_size, device="cuda") * 0.1
    state = eng.new_state(B)
    eng.stream(state, x)
    prof = eng.enable_profile(True)
    eng.stream(state, x)
    torch.cuda.synchronize()
    p = prof.cpu().numpy().astype(np.float64) / nh
    nph, nsub = len(eng.PHASES), eng.N_SUB
    sub = p[nph:].reshape(nph, nsub)      # per phase: wait_weights, issue, mma_done, tmem_ld, epi_math (thread 0)
    p = p[:nph]
    tot = p[:-nsub].sum()     # the last five are sub-timers inside the tensor-core phases
    print(f"PROF {name} {eng.precision} B={B} S={eng.streams_per_cta(B)}: {tot:.0f} cycles/hop (CTA 0)")
    for nm, v, sb in sorted(zip(eng.PHASES, p, sub), key=lambda t: -t[1]):
        if v > 0:
            extra = ""
            if sb.sum() > 0:
                extra = "   [waitw %5.0f issue %5.0f mma %5.0f ld %5.0f epi %5.0f rest %5.0f]" % (*sb, v - sb.sum())
            print(f"  {nm:10s} {v:9.0f} cyc  {100 * v / tot:5.1f}%{extra}")


if __name__ == "__main__":
    t0 = time.time()
    if sys.argv[1] == "--prof":
        profile(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    elif sys.argv[1] == "--time":
        timing(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 0)
    else:
        diag(sys.argv[1], int(sys.argv[2]), *(int(a) for a in sys.argv[3:]))
    print(f"  ({time.time() - t0:.1f}s)")
