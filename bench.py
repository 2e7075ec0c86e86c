#!/usr/bin/env python
"""Benchmark of the FastEnhancer per-frame hot path on B200 (BASELINE.json metric: frames/sec & RTF).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 1..5]
                    [--preset 16k_b] [--streams 256] [--precision fp32x3|fp32|tf32|f16|bf16] [--scatter]

`--config N` selects a BASELINE.json configuration (1-based; default 2, the one the metric is quoted on):
  1  FastEnhancer_T, 16 kHz, batch 1, one 10 s utterance, offline `Model.forward` (scripts/test_pytorch.py plumbing)
  2  FastEnhancer_B, 16 kHz hop 256 / win 512, 256 streams per GPU, fp32 arithmetic           (headline)
  3  FastEnhancer_M, 16 kHz, 512 streams per GPU (4 096 over 8 GPUs), "bf16 conv / fp32 GRU"
  4  FastEnhancer_L, 48 kHz, 256 streams per GPU (1 024 over 4 GPUs)
  5  batch-1 latency sweep T/B/S/M/L (p50 / p99 microseconds per hop, one launch per hop)

A *step* is one pass of the hot path over the batch: every stream's whole 10 s utterance (626 hops for B) in ONE
persistent fused-kernel launch.  N>1 (torchrun, one rank per GPU): every rank owns its own streams (streams are
independent: weak scaling, no data-path collective); `value` = frames of all ranks / max-over-ranks time.

Precision of the headline: BASELINE config 2 says fp32, so the default arithmetic is the fp32-ACCURATE family
("fp32x3": tensor-core contractions on split-fp16 operands, three MMAs per product, ~7e-8 RMS like the FMA-pipe fp32
kernels); the reduced-precision families (f16, tf32, bf16) are reported beside it under `variants`.

Timing: CUDA events on the launching (torch current) stream, W >= 3 warm-up steps, barrier + synchronize on both
sides.  The per-step input (164 MB) and output (164 MB) exceed the 126 MB L2, so no step sees a warm cache.
`--impl reference` times the reference's OWN PyTorch implementation of the path (staged, unmodified, under
baseline/_ref by tools/stage_reference.py) on the box's host cores; where that tree is absent it falls back to the C
oracle port and says so in `cpu_baseline.kind`.
"""
from __future__ import annotations

import argparse
import json
import os
import platform
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fastenhancer_b200.config import PRESETS  # noqa: E402
from fastenhancer_b200.fold import fold_to_canonical  # noqa: E402
from fastenhancer_b200.schema import synthetic_state_dict  # noqa: E402
from fastenhancer_b200.synth import synthetic_noisy  # noqa: E402

REF_DIR = os.path.join(ROOT, "baseline", "_ref")
#            preset   streams/GPU  precision (None = the engine's identical-to-reference default)   what BASELINE.json says
CONFIGS = {1: ("16k_t", 1, None, "FastEnhancer_T, 16 kHz, batch=1, single 10 s utterance, offline Model.forward"),
           2: ("16k_b", 256, None, "FastEnhancer_B, 16 kHz hop=256 win=512, batch=256 streams, 1xB200, fp32"),
           3: ("16k_m", 512, "bf16", "FastEnhancer_M, 16 kHz, batch=4096 streams sharded over 8xB200, bf16 conv / fp32 GRU"),
           4: ("48k_l", 256, None, "FastEnhancer_L, 48 kHz, batch=1024 streams over 4xB200"),
           5: ("16k_b", 1, None, "per-frame latency sweep T/B/S/M/L at batch 1")}
DTYPE = {"fp32x3": "f32-accurate: split-fp16 tensor-core operands (hi + lo, 3 MMAs per product), f32 accumulate, f32 elsewhere",
         "fp32": "f32 (every multiply-add on the fp32 FMA pipe)",
         "tf32": "tf32 tensor-core operands, f32 accumulate, f32 elsewhere",
         "f16": "f16 tensor-core operands (conv section; RNNFormer f16 for T/B/S, tf32 for M/L), f32 accumulate, f32 state and elsewhere",
         "bf16": "bf16 tensor-core operands in the conv section, tf32 RNNFormer contractions, f32 accumulate, f32 GRU state and elsewhere"}
# single-thread ONNXRuntime RTFs the reference publishes for the spec2spec graph (README.md:166-240, Xeon column)
README_RTF_XEON = {"16k_t": 0.012, "16k_b": 0.022, "16k_s": 0.034, "16k_m": 0.101, "16k_l": 0.313}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def workload(cfg, seconds: float):
    """hops of one utterance under the streaming runner's framing: pad right by n_fft zeros, iterate
    range(0, length + n_fft - hop, hop) (scripts/test_onnx.py:18,44)."""
    length = int(round(seconds * cfg.sample_rate))
    n_hops = len(range(0, length + cfg.n_fft - cfg.hop_size, cfg.hop_size))
    return length, n_hops


def make_input(cfg, n_streams, length, n_hops, first_stream=0):
    """synthetic noisy speech-like audio; 32 distinct streams tiled to n_streams."""
    base = synthetic_noisy(min(n_streams, 32), length, cfg.sample_rate, first_stream=first_stream)
    reps = -(-n_streams // base.shape[0])
    x = np.zeros((n_streams, n_hops * cfg.hop_size), np.float32)
    x[:, :length] = np.tile(base, (reps, 1))[:n_streams]
    return x


def host_info():
    info = {"cores": os.cpu_count() or 1, "machine": platform.machine()}
    try:
        out = subprocess.run(["lscpu"], capture_output=True, text=True, timeout=10).stdout
        for line in out.splitlines():
            for key, name in (("Model name", "cpu_model"), ("Socket(s)", "sockets"), ("Thread(s) per core", "threads_per_core")):
                if line.startswith(key):
                    info[name] = line.split(":", 1)[1].strip()
    except Exception:
        pass
    return info


def workload_config(args, cfg, n_hops, world):
    """the `config` object: identical in both arms (the reference arm runs a bounded SAMPLE of it, described under `sample`)."""
    size = args.preset.split("_")[1].upper()
    if args.config == 1:       # BASELINE config 1: offline Model.forward on whole utterances (n_hops = frames of one utterance, 1 + L // hop)
        L = int(args.seconds * cfg.sample_rate)
        return {"workload": f"FastEnhancer_{size} {cfg.sample_rate // 1000} kHz offline Model.forward(noisy [{args.streams}, {L}]) -> wav, "
                            f"{n_hops} frames per utterance, fp32 audio in/out, random-init weights",
                "baseline_config": args.config, "baseline_config_text": CONFIGS[args.config][3], "preset": args.preset,
                "streams_per_gpu": args.streams, "streams_total": args.streams * world, "hops_per_step": n_hops,
                "frames_per_step": args.streams * world * n_hops, "parallelism": f"utterances sharded x{world}, no data-path collective",
                "l2": "L2 flushed between steps (a 256 MB buffer is written): the utterance itself is 0.6 MB"}
    return {"workload": f"FastEnhancer_{size} {cfg.sample_rate // 1000} kHz streaming wav2wav, {args.streams} streams/GPU x {args.seconds:g} s "
                        f"({n_hops} hops of {cfg.hop_size}), fp32 audio in/out, random-init folded weights",
            "baseline_config": args.config, "baseline_config_text": CONFIGS[args.config][3], "preset": args.preset,
            "streams_per_gpu": args.streams, "streams_total": args.streams * world, "hops_per_step": n_hops,
            "frames_per_step": args.streams * world * n_hops, "parallelism": f"streams sharded x{world}, no data-path collective",
            "l2": "per-step input+output 2x%.0f MB vs the 126 MB L2" % (args.streams * n_hops * cfg.hop_size * 4 / 1e6)}


class ClockSampler:
    """SM clock / throttle reasons of one GPU sampled during the timed region (NVML)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period, self.samples, self.reasons, self._stop = index, period_s, [], set(), threading.Event()
        self.max_mhz, self.th = None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
            return self
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap",
                 "hw_power_brake": "nvmlClocksThrottleReasonHwPowerBrakeSlowdown"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, attr in names.items():
                    if mask & getattr(nv, attr, 0):
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop.set()
        if self.th is not None:
            self.th.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the reference itself (staged under baseline/_ref) or, where that tree is absent, the C oracle port
# ------------------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(cfg, canon, n_streams, n_hops, threads):
    """frames/s of the C oracle port (oracle/fe_oracle.c, OpenMP over streams) on the host cores."""
    from oracle.oracle import Oracle
    o = Oracle(cfg, canon)
    x = make_input(cfg, n_streams, n_hops * cfg.hop_size, n_hops)
    st = o.new_state(n_streams)
    o.stream(st, x[:, :cfg.hop_size].copy(), n_threads=threads)      # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    o.stream(st, x, n_threads=threads)
    dt = time.perf_counter() - t0
    return n_streams * n_hops / dt, dt


class ReferenceTorch:
    """The reference's own classes (models/fastenhancer/default/model.py, functional/audio_modules.py), imported UNMODIFIED from the
    staged tree, holding the same seed-0 synthetic checkpoint as the engine."""

    def __init__(self, cfg):
        import torch
        if not os.path.isdir(os.path.join(REF_DIR, "reference", "models")):
            raise RuntimeError("baseline/_ref/reference is absent (tools/stage_reference.py stages it where /root/reference is mounted)")
        for p in (os.path.join(REF_DIR, "shims"), os.path.join(REF_DIR, "reference")):
            if p not in sys.path:
                sys.path.insert(0, p)
        from models.fastenhancer.default.model import Model, ONNXModel
        self.torch, self.cfg = torch, cfg
        sd = {k: torch.from_numpy(np.array(v)) for k, v in synthetic_state_dict(cfg, seed=0).items()}
        self.offline = Model(**cfg.to_model_kwargs()).eval()
        self.offline.load_state_dict(sd, strict=True)
        self.onnx = ONNXModel(**cfg.to_model_kwargs()).eval()
        self.onnx.load_state_dict(sd, strict=True)
        self.onnx.remove_weight_reparameterizations()          # what scripts/export_onnx.py:78 does before export

    def caches(self, B, device="cpu"):
        c, t = self.cfg, self.torch
        z = lambda *s: t.zeros(*s, device=device)  # noqa: E731
        return z(B, c.cache_len), z(B, c.cache_len), [z(1, B * c.rf_freq, c.rf_channels) for _ in range(c.rf_blocks)]

    def stream(self, x, n_hops, warm=0, device="cpu"):
        """the streaming loop of scripts/export_onnx.py:130-136 / scripts/test_onnx.py:44-49 (stft -> model -> stft.inverse with explicit
        caches), `n_hops` timed hops after `warm` untimed ones; x [B, (warm + n_hops) * hop] on `device`.  -> seconds."""
        t = self.torch
        m = self.onnx.to(device)
        H = self.cfg.hop_size
        c_stft, c_istft, hs = self.caches(x.size(0), device)
        sync = t.cuda.synchronize if str(device).startswith("cuda") else (lambda: None)
        with t.no_grad():
            for i in range(warm + n_hops):
                if i == warm:
                    sync()
                    t0 = time.perf_counter()
                spec_in, c_stft = m.stft(x[:, i * H:(i + 1) * H], c_stft)
                spec_out, *hs = m(spec_in, *hs)
                y, c_istft = m.stft.inverse(spec_out, c_istft)
            sync()
        return time.perf_counter() - t0


def reference_cpu_rate(ref, cfg, n_streams, hops, threads):
    """frames/s of the reference's PyTorch streaming graph, `n_streams` streams batched, `threads` intra-op threads."""
    import torch
    torch.set_num_threads(threads)
    x = torch.from_numpy(make_input(cfg, n_streams, (hops + 2) * cfg.hop_size, hops + 2))
    dt = ref.stream(x, hops, warm=2)
    return n_streams * hops / dt, dt


def run_reference_offline(args, cfg, threads):
    """BASELINE config 1 on the reference: its own Model.forward (models/fastenhancer/default/model.py:711-735, unmodified sources,
    PyTorch CPU, all threads) on the same [B, 10 s] batch of utterances."""
    import torch
    ref = ReferenceTorch(cfg)
    B, L, H = args.streams * max(1, args.gpus), int(args.seconds * cfg.sample_rate), cfg.hop_size
    wav = torch.from_numpy(synthetic_noisy(B, L, cfg.sample_rate))
    T = 1 + L // H
    times, extras = [], {}
    with torch.no_grad():
        torch.set_num_threads(threads)
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            ref.offline(wav)
            if i >= args.warmup:
                times.append(time.perf_counter() - t0)
        torch.set_num_threads(1)
        ref.offline(wav)
        t0 = time.perf_counter()
        ref.offline(wav)
        dt1 = time.perf_counter() - t0
        extras["offline_1thread"] = {"ms": dt1 * 1e3, "rtf": dt1 / (B * args.seconds), "frames_per_s": B * T / dt1}
    value = float(np.mean([B * T / t for t in times]))
    how = "the reference's own Model.forward (models/fastenhancer/default/model.py:711-735), unmodified sources, PyTorch CPU"
    line = {
        "impl": "reference", "metric": "frames_per_second", "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "rtf": float(np.mean(times) / (B * args.seconds)),
        "config": workload_config(args, cfg, T, max(1, args.gpus)),
        "sample": {"hops_per_step": T, "frames_per_step": B * T, "of_hops": T, "note": "the whole workload, every step"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "reference", "sample": f"{B} utterance(s) of {args.seconds:g} s per step; {how}"},
        "host": host_info(), "reference_paths": extras,
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_reference(args, cfg, canon, rank, world):
    """Reference arm: the reference's own implementation of the path on the host cores, all threads, bounded sample of the same workload."""
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    n_streams = args.streams * max(1, args.gpus)          # the whole job of the N-GPU arm (weak scaling)
    _, n_hops_full = workload(cfg, args.seconds)
    extras = {}
    if args.config == 1:
        return run_reference_offline(args, cfg, threads)
    try:
        ref = ReferenceTorch(cfg)
        kind = "reference"
        probe_rate, _ = reference_cpu_rate(ref, cfg, n_streams, 3, threads)
        hops = int(max(3, min(args.ref_seconds * probe_rate / n_streams, n_hops_full)))
        times = []
        for i in range(args.warmup + args.steps):
            _, dt = reference_cpu_rate(ref, cfg, n_streams, hops, threads)
            if i >= args.warmup:
                times.append(dt)
        how = "the reference's own PyTorch streaming graph (ONNXModel + ONNXSTFT, scripts/export_onnx.py:130-136), unmodified sources"
        # the shapes BASELINE.md section 3 asks for, beside the headline: batch 1 at one thread (the published-RTF harness shape),
        # offline Model.forward on one 10 s utterance at 1 / all threads, and the same modules CUDA-eager on the B200
        H = cfg.hop_size
        torch.set_num_threads(1)
        x1 = torch.from_numpy(make_input(cfg, 1, 320 * H, 320))
        dt1 = ref.stream(x1, 300, warm=20)
        extras["streaming_batch1_1thread"] = {"us_per_hop": dt1 / 300 * 1e6, "rtf": dt1 / (300 * H / cfg.sample_rate), "frames_per_s": 300 / dt1,
                                             "readme_ort_rtf_xeon_spec2spec": README_RTF_XEON.get(args.preset)}
        wav = torch.from_numpy(synthetic_noisy(1, int(args.seconds * cfg.sample_rate), cfg.sample_rate))
        for th in (1, threads):
            torch.set_num_threads(th)
            with torch.no_grad():
                ref.offline(wav)
                t0 = time.perf_counter()
                ref.offline(wav)
                dt = time.perf_counter() - t0
            extras[f"offline_10s_{th}thread"] = {"ms": dt * 1e3, "rtf": dt / args.seconds, "frames_per_s": (1 + wav.size(1) // H) / dt}
        if torch.cuda.is_available():
            dev = "cuda:0"
            for B, nh in ((1, 300), (args.streams, 40)):
                xb = torch.from_numpy(make_input(cfg, B, (nh + 10) * H, nh + 10)).to(dev)
                dt = ref.stream(xb, nh, warm=10, device=dev)
                extras[f"cuda_eager_streaming_batch{B}"] = {"us_per_hop": dt / nh * 1e6, "frames_per_s": B * nh / dt}
            ref.onnx.to("cpu")
    except Exception as ex:      # staged tree absent (or broken): the oracle port, labelled as such
        extras["reference_unavailable"] = f"{type(ex).__name__}: {ex}"
        kind = "port"
        probe_rate, _ = cpu_oracle_rate(cfg, canon, min(n_streams, threads), 4, threads)
        hops = int(max(2, min(args.ref_seconds * probe_rate / n_streams, n_hops_full)))
        times = []
        for i in range(args.warmup + args.steps):
            _, dt = cpu_oracle_rate(cfg, canon, n_streams, hops, threads)
            if i >= args.warmup:
                times.append(dt)
        how = "C oracle port of the reference's algorithm (oracle/fe_oracle.c, OpenMP) -- NOT the reference's own runtime"
    value = float(np.mean([n_streams * hops / t for t in times]))
    sample = f"{n_streams} streams x first {hops} of the {n_hops_full} hops of the {args.seconds:g} s utterance per step; {how}"
    line = {
        "impl": "reference", "metric": "frames_per_second", "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "rtf": float(np.mean(times) / (hops * cfg.hop_size / cfg.sample_rate)),
        "config": workload_config(args, cfg, n_hops_full, max(1, args.gpus)),
        "sample": {"hops_per_step": hops, "frames_per_step": n_streams * hops, "of_hops": n_hops_full,
                   "note": "per-frame cost is hop-independent (constant state per stream): frames/s of the sample = frames/s of the workload"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "host": host_info(), "reference_paths": extras,
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def latency_sweep(dev, precision, n_hops=1500, warm=200):
    """BASELINE config 5: one stream, one fused-kernel launch per hop, CUDA-event time of every hop."""
    import torch
    from fastenhancer_b200.engine import Engine
    rows = []
    for name in ("16k_t", "16k_b", "16k_s", "16k_m", "16k_l"):
        cfg = PRESETS[name]
        eng = Engine(cfg, fold_to_canonical(cfg, synthetic_state_dict(cfg, 0)), dev, precision=precision)
        H = cfg.hop_size
        x = torch.from_numpy(synthetic_noisy(1, (n_hops + warm) * H, cfg.sample_rate)).to(dev)
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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.stream(st, x[:, :n_hops * H], out=y[:, :n_hops * H]); e1.record(); torch.cuda.synchronize()
        hop_us = H / cfg.sample_rate * 1e6
        rows.append({"preset": name, "precision": eng.precision, "hops": n_hops, "p50_us": float(np.percentile(us, 50)),
                     "p99_us": float(np.percentile(us, 99)), "mean_us": float(us.mean()),
                     "persistent_us_per_hop": e0.elapsed_time(e1) * 1e3 / n_hops, "hop_duration_us": hop_us,
                     "rtf_p50": float(np.percentile(us, 50)) / hop_us,
                     "one_sm_fp32_floor_us": cfg.flops_per_frame() / (128 * 2 * 1.965e9) * 1e6})
    return rows


def reference_shaped_calls(args, cfg, dev, x_host, n_hops, precision):
    """End-to-end through the reference-facing classes (fastenhancer_b200.model), host buffers in, host buffers out:
      * StreamingModel.forward hop by hop with explicit caches -- the `sess.run` replacement of scripts/test_onnx.py:44-49 --
        at batch 1 and at the bench batch; every hop copies its input chunk host->device and its output chunk device->host;
      * Model.forward on [B, 10 s] -- what scripts/test_pytorch.py:29-37 does per file;
      * ONNXModel.forward frame by frame (spec2spec, scripts/test_onnx_spec.py:55-62), the graph the README RTFs are measured on."""
    import torch
    from fastenhancer_b200.model import Model, ONNXModel, StreamingModel
    H, kw, out = cfg.hop_size, cfg.to_model_kwargs(), {}
    om = ONNXModel(**kw).eval().to(dev)
    om.precision = precision
    sm = StreamingModel(om)
    for B in sorted({1, args.streams}):
        nh = min(n_hops, 300)
        xh = x_host[:B, :nh * H].contiguous().pin_memory()
        yh = torch.empty_like(xh).pin_memory()
        xd = torch.empty(B, H, device=dev)
        caches = [c.to(dev) for c in sm.initialize_cache(xd)]
        for it in range(2):               # first pass = warm-up (engine creation, packing)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(nh):
                xd.copy_(xh[:, i * H:(i + 1) * H], non_blocking=True)
                y, *caches = sm(xd, *caches)
                yh[:, i * H:(i + 1) * H].copy_(y, non_blocking=True)
                torch.cuda.current_stream().synchronize()          # the reference loop is synchronous per hop (sess.run returns numpy)
            dt = time.perf_counter() - t0
        out[f"streaming_model_forward_batch{B}"] = {
            "value": B * nh / dt, "unit": "frames/s", "us_per_hop": dt / nh * 1e6, "rtf": dt / (nh * H / cfg.sample_rate), "hops": nh,
            "h2d_bytes_per_hop": B * H * 4, "d2h_bytes_per_hop": B * H * 4, "kernel_launches_per_hop": 1,
            "api": "StreamingModel.forward(wav_in, cache_stft, cache_istft, *h) with the returned caches fed back (zero-copy views)"}
    # offline Model.forward on one utterance (BASELINE config 1's own call) and on the whole batch of utterances
    m = Model(**kw).eval().to(dev)
    m.precision = precision
    L = int(args.seconds * cfg.sample_rate)
    for B in sorted({1, args.streams}):
        wav_h = x_host[:B, :L].contiguous().pin_memory()
        out_h = torch.empty(B, H * (L // H)).pin_memory()
        for it in range(4):
            torch.cuda.synchronize()
            n0 = m.engine.kernel_launches if it else 0
            t0 = time.perf_counter()
            wav, _spec = m(wav_h.to(dev, non_blocking=True))
            out_h.copy_(wav, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        launches = m.engine.kernel_launches - n0
        frames = B * (1 + L // H)
        out[f"model_forward_offline_batch{B}"] = {
            "value": frames / dt, "unit": "frames/s", "ms": dt * 1e3, "rtf": dt / (B * args.seconds), "h2d_bytes": wav_h.numel() * 4,
            "d2h_bytes": out_h.numel() * 4, "kernel_launches": launches,
            "schedule": "sequential walk (one CTA per group of utterances)" if launches == 1 else
                        "frame-parallel (stage A, GRU scan + stage B per block, overlap-add)",
            "api": "Model.forward(noisy [B, L]) -> wav (spec_hat stays on the device, as in scripts/test_pytorch.py:34-37)"}
    out["model_forward_offline"] = out[f"model_forward_offline_batch{args.streams}"]
    return out


def spec2spec_rtf(dev, precision, n_frames=1000):
    """ONNXModel.forward frame by frame at batch 1 (the loop of scripts/test_onnx_spec.py:55-62), beside the README's ORT RTFs."""
    import torch
    from fastenhancer_b200.model import ONNXModel
    rows = {}
    for name in ("16k_t", "16k_b", "16k_s", "16k_m", "16k_l"):
        cfg = PRESETS[name]
        om = ONNXModel(**cfg.to_model_kwargs()).eval().to(dev)
        om.precision = precision if precision in ("tf32", "fp32", "f16") else None
        spec = torch.randn(1, cfg.n_fft // 2 + 1, n_frames + 20, 2, device=dev) * 0.3
        hs = om.initialize_cache(spec)
        for i in range(n_frames + 20):
            if i == 20:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            o, *hs = om(spec[:, :, i:i + 1].contiguous(), *hs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rows[name] = {"rtf": dt / (n_frames * cfg.hop_size / cfg.sample_rate), "us_per_frame": dt / n_frames * 1e6,
                      "precision": om.engine.precision, "readme_ort_rtf_xeon_1thread": README_RTF_XEON[name]}
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configuration (1-based)")
    ap.add_argument("--preset", default=None, help="override the configuration's model preset")
    ap.add_argument("--streams", type=int, default=None, help="override streams per GPU")
    ap.add_argument("--seconds", type=float, default=10.0, help="utterance length per stream")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline sample")
    ap.add_argument("--ref-seconds", type=float, default=8.0, help="CPU work per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the reference-shaped-call legs, the variants and the sweeps")
    ap.add_argument("--streams-per-cta", type=int, default=0)
    ap.add_argument("--scatter", action="store_true", help="N>1: also time the NCCL scatter / gather of a batch held on rank 0")
    ap.add_argument("--precision", default=None, choices=sorted(DTYPE),
                    help="arithmetic family (default: the configuration's -- fp32-accurate for config 2, bf16 for config 3)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    preset, streams, prec, _ = CONFIGS[args.config]
    args.preset = args.preset or preset
    args.streams = args.streams or streams
    args.precision = args.precision or prec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = PRESETS[args.preset]
    canon = fold_to_canonical(cfg, synthetic_state_dict(cfg, 0))

    if args.impl == "reference":
        run_reference(args, cfg, canon, rank, world)
        return

    import torch
    import torch.distributed as dist
    from fastenhancer_b200.engine import Engine, library_path, measured_fma_tflops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.streams
    length, n_hops = workload(cfg, args.seconds)
    H = cfg.hop_size
    eng = Engine(cfg, canon, dev, precision=args.precision)
    precision = eng.precision                 # what the engine resolved None to
    if args.streams_per_cta:
        eng.set_streams_per_cta(args.streams_per_cta)
    x_host = torch.from_numpy(make_input(cfg, B, length, n_hops, first_stream=rank * B)).pin_memory()
    y_host = torch.empty_like(x_host).pin_memory()
    x = x_host.to(dev)                      # inputs resident in HBM before the timed region
    y = torch.empty_like(x)
    state = eng.new_state(B)
    # BASELINE config 1 is the OFFLINE call: Model.forward on whole utterances (fe_offline; few long utterances run the frame-parallel
    # schedule).  Its working set (0.6 MB per utterance) fits the L2, so the L2 is flushed between steps; the flush is outside the events.
    offline = args.config == 1
    L_off = int(args.seconds * cfg.sample_rate)
    if offline:
        n_hops = 1 + L_off // H                     # frames of one utterance
        x_off = x[:, :L_off].contiguous()
        flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def step(engine, st):
        if offline:
            engine.offline(x_off, want_spec=True)
        else:
            engine.stream(st, x, out=y)

    def timed(engine, st, steps):
        if offline:
            total = 0.0
            for _ in range(steps):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step(engine, st)
                e1.record()
                e1.synchronize()
                total += e0.elapsed_time(e1)
            barrier()
            return max_over_ranks(total) / steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(engine, st)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    # ---------------- device-resident throughput (`value`) ----------------
    for _ in range(args.warmup):
        step(eng, state)
    barrier()
    sampler = ClockSampler(local_rank).start()
    launches0 = eng.kernel_launches
    ms_step = timed(eng, state, args.steps)
    clocks = sampler.stop()
    launches = eng.kernel_launches - launches0
    frames_step = world * B * n_hops
    value = frames_step / (ms_step * 1e-3)

    # ---------------- end to end through the C ABI's host-buffer call (`e2e`) ----------------
    e2e_steps = max(3, args.steps // 2)
    if offline:
        xo_host = x_host[:, :L_off].contiguous().pin_memory()
        yo_host = torch.empty(B, H * (L_off // H)).pin_memory()

        def e2e_step():
            wav, _spec = eng.offline(xo_host.to(dev, non_blocking=True), want_spec=True)
            yo_host.copy_(wav, non_blocking=True)
            torch.cuda.synchronize()
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
        barrier()
        io_bytes, out_bytes = B * L_off * 4, B * H * (L_off // H) * 4
        e2e_api = "Engine.offline (fe_offline) as Model.forward calls it: pinned host wav -> device, enhanced wav -> pinned host, synchronous"
    else:
        e2e_state = eng.new_state(B)
        e2e_state.reserve_host(64)
        for _ in range(2):
            eng.stream_host(e2e_state, x_host, out=y_host)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.stream_host(e2e_state, x_host, out=y_host)     # returns after the last D2H copy completed
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
        barrier()
        io_bytes = out_bytes = B * n_hops * H * 4
        e2e_api = "fe_stream_host via Engine.stream_host (pinned host buffers, copies pipelined with the kernel)"
    e2e_value = frames_step / e2e_s

    # ---------------- NCCL scatter / gather of a batch held on rank 0 (reported separately) ----------------
    scatter = None
    if args.scatter and world > 1:
        from fastenhancer_b200.sharding import gather_streams, scatter_streams
        full = torch.cat([x] * world, dim=0) if rank == 0 else None
        for _ in range(2):
            mine = scatter_streams(full, B * world, x.size(1), 0, dev)
            gather_streams(mine, B * world)
        barrier()
        s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        s0.record()
        mine = scatter_streams(full, B * world, x.size(1), 0, dev)
        s1.record()
        gather_streams(mine, B * world)
        s2.record()
        barrier()
        scatter = {"scatter_ms": max_over_ranks(s0.elapsed_time(s1)), "gather_ms": max_over_ranks(s1.elapsed_time(s2)),
                   "bytes_each_way": int(B * world * x.size(1) * 4), "backend": "nccl"}
        del full, mine

    # ---------------- the other arithmetic families, same workload, a few steps ----------------
    other_ms = {}
    if not args.no_extras:
        for other in ("fp32x3", "f16", "tf32", "bf16", "fp32"):
            if other == precision:
                continue
            try:
                eng_o = Engine(cfg, canon, dev, precision=other)
            except RuntimeError:
                continue              # this model has no kernels of that family
            st_o = eng_o.new_state(B)
            for _ in range(2):
                step(eng_o, st_o)
            barrier()
            other_ms[other] = timed(eng_o, st_o, max(3, args.steps // 3))
            del eng_o, st_o

    calls = s2s = sweep = None
    if rank == 0 and not args.no_extras:
        calls = reference_shaped_calls(args, cfg, dev, x_host, n_hops, args.precision)
        if args.config in (2, 5):
            s2s = spec2spec_rtf(dev, args.precision)
        if args.config == 5:
            sweep = latency_sweep(dev, args.precision)
    barrier()

    if rank == 0:
        peaks = measured_peaks()
        flops_launch = cfg.flops_per_frame() * B * n_hops           # algorithmic FLOP of ONE launch (one rank)
        t_launch = ms_step * 1e-3
        achieved_tf = flops_launch / t_launch / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{args.preset}:{B}:{n_hops}")
        # The contraction-heavy path is arithmetic-bound (~4000 FLOP per HBM byte).  Tensor-core families are judged
        # against the MEASURED dense bf16 tensor peak (sustained: the kernel is timed inside a long step); the fp32 FMA-pipe family
        # against the fp32 FMA throughput measured on this GPU by fe_microbench_fma.
        if precision == "fp32":
            fma_peak = measured_fma_tflops(local_rank)
            roofline = {"bound": "fp32-fma", "achieved": achieved_tf, "peak": fma_peak, "unit": "TFLOP/s", "frac": achieved_tf / fma_peak,
                        "traffic": traffic, "peak_source": "measured on this GPU (fe_microbench_fma: 16 independent FFMA chains per thread)"}
        else:
            peak = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
            roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tf / peak,
                        "traffic": traffic, "peak_source": peaks["source"] + " (MEASURED_PEAKS.json bf16_tflops_sustained)"}
            if precision == "fp32x3":
                roofline["note"] = ("achieved counts ALGORITHMIC flops; the split-fp16 family executes 3 tensor-core MACs per algorithmic MAC, "
                                    "so the tensor pipe does 3x this work")
        roofline["kernel"] = "fe_fused_kernel (one persistent launch per step)" if launches == args.steps else \
            f"fe_fused_kernel stages + fe_gru_scan_kernel + fe_overlap_add_kernel ({launches // max(1, args.steps)} launches per step: frame-parallel offline schedule)"
        roofline["hbm"] = {"achieved_gbs": (io_bytes + out_bytes) / t_launch / 1e9, "peak_gbs": peaks["hbm_gbs"],
                           "frac": (io_bytes + out_bytes) / t_launch / 1e9 / peaks["hbm_gbs"]}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            try:
                ref = ReferenceTorch(cfg)
                probe_rate, _ = reference_cpu_rate(ref, cfg, B, 3, threads)
                hops = int(max(3, min(args.cpu_seconds * probe_rate / B, n_hops)))
                rate, dt = reference_cpu_rate(ref, cfg, B, hops, threads)
                cpu = {"value": rate, "unit": "frames/s", "cores": threads, "kind": "reference",
                       "sample": f"{B} streams (batched) x first {hops} hops of the utterance, {dt:.1f} s of wall time; the reference's own "
                                 f"PyTorch streaming graph (scripts/export_onnx.py:130-136), {threads} intra-op threads", "host": host_info()}
            except Exception as ex:
                probe_rate, _ = cpu_oracle_rate(cfg, canon, min(B, threads), 4, threads)
                hops = int(max(2, min(args.cpu_seconds * probe_rate / B, n_hops)))
                rate, dt = cpu_oracle_rate(cfg, canon, B, hops, threads)
                cpu = {"value": rate, "unit": "frames/s", "cores": threads, "kind": "port",
                       "sample": f"{B} streams x first {hops} hops of the utterance, {dt:.1f} s of wall time (C oracle port, OpenMP; "
                                 f"staged reference unavailable: {type(ex).__name__})", "host": host_info()}
        line = {
            "metric": "frames_per_second", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE[precision], "data": "synthetic",
            "rtf": (ms_step * 1e-3) / (args.seconds if offline else n_hops * H / cfg.sample_rate),
            "config": workload_config(args, cfg, n_hops, world),
            "impl_config": {"precision": precision, "streams_per_cta": eng.streams_per_cta(B)},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": io_bytes, "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": e2e_s * 1e3, "api": e2e_api},
            "e2e_reference_shaped_calls": calls, "spec2spec_rtf": s2s, "latency_sweep": sweep, "scatter_gather": scatter,
            "variants": {precision: {"value": value, "ms_per_step": ms_step},
                         **{o: {"value": frames_step / (ms * 1e-3), "ms_per_step": ms} for o, ms in other_ms.items()},
                         "note": "fp32x3 / fp32 reproduce the fp32 reference (~7e-8 RMS); f16 / tf32 (11-bit significand operands) ~7e-6 RMS and "
                                 "bf16 conv section ~3e-5 RMS on the near-identity-mask checkpoint (tests/test_gpu_parity.py)"},
            "gpu_launches": int(launches), "clocks": clocks, "library": os.path.relpath(library_path(), ROOT),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
