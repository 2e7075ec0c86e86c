"""CPU emulation of the fused kernel body (tests/emu) against the oracle.

The emulation compiles fastenhancer_b200/csrc/fe_kernel.cuh for the host and runs every
barrier-separated phase as a loop over thread ids, reading weights from the very blob the GPU
kernel consumes.  It pins the packer, the tile / layout index math, the buffer aliasing plan and the
chunk schedule without a GPU; the GPU tests (test_gpu_parity.py) pin the real kernel.
Tolerance: 2e-6 absolute on waveforms of RMS ~0.1 (fp32, different summation order)."""
import numpy as np
import pytest

from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.synth import synthetic_noisy
from oracle.oracle import Oracle, tap_schema
from emu import emu

VARIANTS = [("16k_t", 4), ("16k_b", 2), ("16k_b", 1), ("16k_m", 1), ("48k_t", 2), ("48k_s", 1)]
# fp32 variants: FMA-pipe arithmetic; tensor-core variants: TF32 operands (rounded to nearest), fp32 accumulation
# precision 2: as True, with the conv section's operands stored as fp16 (same 11-bit significand as TF32)
# precision 3: bfloat16 conv section (8-bit significand), TF32 RNNFormer -- BASELINE config 3's "bf16 conv / fp32 GRU"
# precision 4: split-fp16 operands (hi + lo, three MMAs per product): held to the fp32 tolerances
TOL = {False: dict(wav=2e-6, state=5e-6, tap=2e-5, spec=1e-5, spec_abs=1e-4), True: dict(wav=5e-5, state=3e-3, tap=3e-3, spec=1e-3, spec_abs=3e-3),
       2: dict(wav=5e-5, state=3e-3, tap=3e-3, spec=1e-3, spec_abs=3e-3),
       3: dict(wav=1e-4, state=5e-3, tap=2e-2, spec=1e-2, spec_abs=2e-2),
       4: dict(wav=2e-6, state=5e-6, tap=2e-5, spec=1e-5, spec_abs=1e-4)}
F16_VARIANTS = [("16k_t", 4), ("16k_b", 2), ("16k_b", 4), ("16k_s", 2), ("16k_m", 1), ("16k_m", 2), ("48k_t", 2), ("48k_s", 2), ("48k_l", 1)]    # a sample of the fp16 conv-section variants
BF16_VARIANTS = [("16k_b", 2), ("16k_s", 2), ("16k_m", 1), ("16k_m", 2), ("48k_l", 1)]
TF32_EXTRA = [("16k_s", 2), ("48k_b", 2)]           # two-stream variants that exist in the tensor-core families only
SPLIT_VARIANTS = [("16k_t", 2), ("16k_b", 2), ("16k_b", 1), ("48k_b", 1), ("48k_b", 2)]


@pytest.mark.parametrize("name,S,tc", [(n, s, t) for n, s in VARIANTS for t in (False, True)] + [(n, s, 2) for n, s in F16_VARIANTS] +
                         [(n, s, 3) for n, s in BF16_VARIANTS] + [(n, s, 4) for n, s in SPLIT_VARIANTS] + [(n, s, True) for n, s in TF32_EXTRA])
def test_streaming_and_state_round_trip(name, S, tc, canonical):
    cfg = PRESETS[name]
    canon = canonical(name)
    o = Oracle(cfg, canon)
    B, nh, H = 3, 4, cfg.hop_size                        # B not a multiple of S: ragged last CTA
    x = synthetic_noisy(B, nh * H, cfg.sample_rate)
    st = o.new_state(B)
    want, taps_ref = o.stream(st, x, taps=True)
    stn = emu.to_native(cfg, o.new_state(B))
    got = np.zeros_like(x)
    dbg = np.zeros(emu.tap_total(cfg), np.float32)
    assert dbg.size == o.tap_floats
    # two launches (2 + 2 hops): the overlap / GRU state must survive the round trip through global memory
    x1, x2 = np.ascontiguousarray(x[:, :2 * H]), np.ascontiguousarray(x[:, 2 * H:])
    y1, y2 = np.zeros_like(x1), np.zeros_like(x2)
    emu.run(cfg, S, canon, emu.MODE_STREAM, stn, x1, y1, n_streams=B, n_hops=2, ld_in=2 * H, ld_out=2 * H, tc=tc)
    emu.run(cfg, S, canon, emu.MODE_STREAM, stn, x2, y2, n_streams=B, n_hops=2, ld_in=2 * H, ld_out=2 * H, dbg=dbg, dbg_hop=1, tc=tc)
    got = np.concatenate([y1, y2], axis=1)
    assert np.sqrt(np.mean((got - want) ** 2)) < TOL[tc]["wav"]
    assert np.abs(emu.to_canonical(cfg, stn) - st).max() < TOL[tc]["state"]
    assert np.abs(emu.to_canonical(cfg, stn) - st).max() < TOL[tc]["state"]
    off = 0
    for nm, shp in tap_schema(cfg):
        n = int(np.prod(shp))
        ref = taps_ref[nm][3]
        assert np.abs(dbg[off:off + n].reshape(shp) - ref).max() < TOL[tc]["tap"] * max(1.0, np.abs(ref).max()), nm
        off += n


@pytest.mark.parametrize("name,S,tc", [("16k_t", 2, False), ("16k_t", 2, True), ("16k_m", 1, False), ("16k_m", 1, True), ("16k_t", 2, 2), ("16k_b", 2, 2),
                                       ("16k_m", 1, 3), ("16k_b", 2, 4)])
def test_spec_and_offline_modes(name, S, tc, canonical):
    cfg = PRESETS[name]
    canon = canonical(name)
    o = Oracle(cfg, canon)
    B, T, N, H = 3, 3, cfg.n_fft, cfg.hop_size
    spec = (np.random.RandomState(1).standard_normal((B, N // 2 + 1, T, 2)) * 0.5).astype(np.float32)
    h = np.zeros((B, cfg.rf_blocks, cfg.rf_freq, cfg.rf_channels), np.float32)
    want = o.spec(h, spec)
    stn = emu.to_native(cfg, o.new_state(B))
    got = np.full_like(spec, np.nan)
    emu.run(cfg, S, canon, emu.MODE_SPEC, stn, spec, got, n_streams=B, n_hops=T, tc=tc)
    assert np.abs(got - want).max() < TOL[tc]["spec"] * np.abs(want).max()
    assert np.all(got[:, -1] == 0)
    L = 5 * H + 37                                       # ragged length: the reference floors to L // hop frames
    w = synthetic_noisy(B, L, cfg.sample_rate)
    w_ref, sp_ref = o.offline(w)
    stn = emu.to_native(cfg, o.new_state(B))
    w_out, sp_out = np.full_like(w_ref, np.nan), np.full_like(sp_ref, np.nan)
    emu.run(cfg, S, canon, emu.MODE_OFFLINE, stn, w, w_out, spec_out=sp_out, n_streams=B, n_hops=1 + L // H, L=L, tc=tc)
    assert np.sqrt(np.mean((w_out - w_ref) ** 2)) < TOL[tc]["wav"]
    assert np.abs(sp_out - sp_ref).max() < TOL[tc]["spec_abs"] * max(1.0, np.abs(sp_ref).max())


@pytest.mark.parametrize("name,S,tc", [("16k_b", 2, True), ("16k_m", 1, False), ("48k_t", 2, True)])
def test_standalone_stft_istft_modes(name, S, tc, canonical):
    """fe_stft / fe_istft kernels (ONNXSTFT.forward / inverse on their own) against numpy's rfft / irfft, including the
    Nyquist bin that the fused path drops, ragged hops (M: hop 160) and the cache hand-over."""
    cfg = PRESETS[name]
    canon = canonical(name)
    o = Oracle(cfg, canon)
    B, nh, H, N = 3, 5, cfg.hop_size, cfg.n_fft
    w, wi = o.windows()
    x = synthetic_noisy(B, nh * H, cfg.sample_rate)
    hist = np.concatenate([np.zeros((B, N - H), np.float32), x], axis=1)
    ref = np.stack([np.fft.rfft(hist[:, t * H:t * H + N].astype(np.float64) * w, axis=1) for t in range(nh)], axis=2)
    stn = emu.to_native(cfg, o.new_state(B))
    spec = np.full((B, N // 2 + 1, nh, 2), np.nan, np.float32)
    emu.run(cfg, S, canon, 3, stn, x, spec, n_streams=B, n_hops=nh, ld_in=nh * H, tc=tc)
    assert np.abs((spec[..., 0] + 1j * spec[..., 1]) - ref).max() < 1e-6 * np.abs(ref).max()
    assert np.array_equal(emu.to_canonical(cfg, stn)[:, :N - H], hist[:, -(N - H):])
    sp = np.random.RandomState(2).standard_normal((B, N // 2 + 1, nh, 2)).astype(np.float32)
    frames = np.fft.irfft(sp[..., 0].astype(np.float64) + 1j * sp[..., 1], n=N, axis=1) * wi[None, :, None]
    acc = np.zeros((B, N + H * nh))
    for t in range(nh):
        acc[:, t * H:t * H + N] += frames[:, :, t]
    stn = emu.to_native(cfg, o.new_state(B))
    wav = np.full((B, nh * H), np.nan, np.float32)
    emu.run(cfg, S, canon, 4, stn, sp, wav, n_streams=B, n_hops=nh, ld_out=nh * H, tc=tc)
    assert np.abs(wav - acc[:, :nh * H]).max() < 1e-6
    assert np.abs(emu.to_canonical(cfg, stn)[:, N - H:2 * (N - H)] - acc[:, nh * H:nh * H + N - H]).max() < 1e-6


@pytest.mark.parametrize("name,S,tc", [("16k_b", 2, True), ("16k_b", 2, 2), ("16k_m", 1, True), ("16k_t", 2, False)])
def test_nonzero_attention_bias(name, S, tc, canonical):
    """attn_bias is False in every shipped config, so the packed qkv bias is all zeros and the tensor-core epilogue skips it;
    a model with a bias must take the other branch and still match the oracle."""
    from fastenhancer_b200.schema import flatten_canonical, split_canonical
    cfg = PRESETS[name]
    parts = split_canonical(cfg, canonical(name))
    rs = np.random.RandomState(5)
    for k in range(cfg.rf_blocks):
        parts[f"blk.{k}.qkv.b"] = (1.0 * rs.standard_normal(parts[f"blk.{k}.qkv.b"].shape)).astype(np.float32)
    canon = flatten_canonical(cfg, parts)
    o = Oracle(cfg, canon)
    B, nh, H = 3, 3, cfg.hop_size
    x = synthetic_noisy(B, nh * H, cfg.sample_rate)
    st = o.new_state(B)
    want = o.stream(st, x)
    base = Oracle(cfg, canonical(name)).stream(Oracle(cfg, canonical(name)).new_state(B), x)
    assert np.abs(want - base).max() > 0           # the bias changes the output (by how much depends on the config)
    stn = emu.to_native(cfg, o.new_state(B))
    got = np.zeros_like(x)
    emu.run(cfg, S, canon, emu.MODE_STREAM, stn, x, got, n_streams=B, n_hops=nh, ld_in=nh * H, ld_out=nh * H, tc=tc)
    assert np.sqrt(np.mean((got - want) ** 2)) < TOL[tc]["wav"]
    assert np.abs(emu.to_canonical(cfg, stn) - st).max() < TOL[tc]["state"]


@pytest.mark.parametrize("name,S,tc", [("16k_b", 2, 2), ("16k_t", 2, 4), ("48k_t", 2, True)])
def test_hop_tiles_without_tma(name, S, tc, canonical, monkeypatch):
    """hop-tiled rings filled / drained with plain loads and stores (the path taken for arrays that are not 16-byte aligned / pitched)
    give the same bits as the emulated TMA tiles"""
    cfg = PRESETS[name]
    canon = canonical(name)
    B, nh, H = 3, 4, cfg.hop_size
    x = synthetic_noisy(B, nh * H, cfg.sample_rate)
    outs = []
    for no_tma in (False, True):
        if no_tma:
            monkeypatch.setenv("FE_EMU_NO_HOP_TMA", "1")
        stn = emu.to_native(cfg, np.zeros((B, cfg.state_floats), np.float32))
        y = np.zeros_like(x)
        emu.run(cfg, S, canon, emu.MODE_STREAM, stn, x, y, n_streams=B, n_hops=nh, ld_in=nh * H, ld_out=nh * H, tc=tc)
        outs.append((y, stn))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert np.abs(outs[0][0]).max() > 0.01


@pytest.mark.parametrize("name,S,grid", [("16k_t", 2, 3), ("16k_t", 4, 2), ("16k_b", 2, 4), ("16k_b", 1, 5), ("16k_m", 1, 3), ("48k_s", 1, 2)])
def test_offline_frame_parallel_schedule(name, S, grid, canonical):
    """Model.forward through the frame-parallel offline schedule (fp32 family: stage A -> GRU scan -> stage B per block -> overlap-add,
    frames instead of streams in the CTA slots) equals the sequential offline walk: ragged length, frame count not a multiple of S,
    more frame groups than CTAs, two utterances whose frames share a group."""
    cfg = PRESETS[name]
    canon = canonical(name)
    o = Oracle(cfg, canon)
    B, H = 2, cfg.hop_size
    L = 6 * H + 37
    w = synthetic_noisy(B, L, cfg.sample_rate)
    w_ref, sp_ref = o.offline(w)
    w_out, sp_out = np.full_like(w_ref, np.nan), np.full_like(sp_ref, np.nan)
    emu.offline_tp(cfg, S, canon, w, w_out, spec_out=sp_out, grid=grid)
    assert np.sqrt(np.mean((w_out - w_ref) ** 2)) < TOL[False]["wav"]
    assert np.abs(sp_out - sp_ref).max() < TOL[False]["spec_abs"] * max(1.0, np.abs(sp_ref).max())


@pytest.mark.parametrize("name,S,tc", [("16k_m", 1, 2), ("16k_m", 2, 3), ("16k_m", 1, True), ("16k_l", 1, False)])
def test_hop_sliced_launch_equals_oracle(name, S, tc, canonical):
    """Hop-sliced streaming launches (KParams::slice_hops): items (hop range, stream group) with the state handed over through global
    memory between ranges -- ring positions, hop-tile prefetch bounds and the overlapped front / back end are range-relative."""
    cfg = PRESETS[name]
    canon = canonical(name)
    o = Oracle(cfg, canon)
    B, nh, H = 3, (7 if name != "16k_l" else 5), cfg.hop_size
    x = synthetic_noisy(B, nh * H, cfg.sample_rate)
    st = o.new_state(B)
    want = o.stream(st, x)
    for slice_hops in ((3, 2) if name != "16k_l" else (2,)):                            # ranges of 3 + 3 + 1 and 2 + 2 + 2 + 1 hops
        stn = emu.to_native(cfg, o.new_state(B))
        got = np.full_like(x, np.nan)
        emu.run(cfg, S, canon, emu.MODE_STREAM, stn, x, got, n_streams=B, n_hops=nh, ld_in=nh * H, ld_out=nh * H, tc=tc, slice_hops=slice_hops)
        assert np.sqrt(np.mean((got - want) ** 2)) < TOL[tc]["wav"], slice_hops
        assert np.abs(emu.to_canonical(cfg, stn) - st).max() < TOL[tc]["state"], slice_hops
