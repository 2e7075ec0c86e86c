"""Parity of the fused CUDA kernel (through the C ABI, fastenhancer_b200.engine) against the CPU oracle and
against the committed golden vectors produced by the reference itself (tools/gen_golden.py).

Tolerance (north-star): <= 1e-4 RMS on the enhanced waveform against the fp32 reference.  All five arithmetic modes of
the engine are tested.  The two fp32-accurate ones -- "fp32x3" (the default where it exists: tensor-core contractions on
split-fp16 operands, three MMAs per product) and "fp32" (every multiply-add on the FMA pipe) -- are held to 1e-5 RMS on the
waveform and 2e-5 on the GRU state; "tf32" (TF32 operands, fp32 accumulation) and "f16" (as tf32, with the conv section's
activations and weights stored as fp16 -- the same 11-bit significand) to 5e-5 RMS; "bf16" (BASELINE config 3's "bf16 conv /
fp32 GRU": bfloat16 conv section, 8-bit significand, TF32 RNNFormer, fp32 state) to 1e-4 RMS.  Frame indexing (hop
alignment, n_fft - hop delay, output lengths, zero Nyquist bin) is checked exactly in all of them.

Those numbers are for the seed-0 synthetic checkpoint, whose mask head has bias [1, 0] (mask ~ identity + a small network
term), which attenuates network error in the waveform.  `test_network_dominated_mask` repeats the comparison on a checkpoint
whose mask is driven by the network alone (zero head bias, final transposed-conv weights x 15) and reports RELATIVE error:
the fp32-accurate modes must stay below 2e-5 relative there, the reduced-precision ones below 5e-3 (tf32 / f16) and
2e-2 (bf16) -- which is why they are opt-in and not the default."""
import numpy as np
import pytest
import torch

from conftest import rms
from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.synth import synthetic_noisy

pytestmark = pytest.mark.gpu
ALL = sorted(PRESETS)
N_HOPS = 24


#                 waveform RMS, state max, tap relative, spectrum relative
TOL = {"fp32": dict(wav=1e-5, state=2e-5, tap=2e-5, spec=1e-5, spec_abs=1e-4),
       "fp32x3": dict(wav=1e-5, state=2e-5, tap=2e-5, spec=1e-5, spec_abs=1e-4),
       "tf32": dict(wav=5e-5, state=3e-3, tap=3e-3, spec=1e-3, spec_abs=3e-3),
       "f16": dict(wav=5e-5, state=3e-3, tap=3e-3, spec=1e-3, spec_abs=3e-3),
       "bf16": dict(wav=1e-4, state=5e-3, tap=2e-2, spec=1e-2, spec_abs=2e-2)}
#: relative waveform RMS error on the network-dominated-mask checkpoint
TOL_NET = {"fp32": 2e-5, "fp32x3": 2e-5, "tf32": 5e-3, "f16": 5e-3, "bf16": 2e-2}
#: presets that have kernels of the optional families (fe_configs.h)
HAS = {"fp32x3": {"16k_t", "16k_b", "48k_t", "48k_b"},
       "bf16": {"16k_t", "16k_b", "16k_s", "16k_m", "16k_l", "48k_m", "48k_l"}}


@pytest.fixture(scope="module", params=["fp32x3", "tf32", "fp32", "f16", "bf16"])
def precision(request):
    return request.param


@pytest.fixture(scope="module")
def engines(canonical, precision):
    from fastenhancer_b200.engine import Engine
    cache = {}

    def get(name):
        if name not in HAS.get(precision, PRESETS):
            pytest.skip(f"{name} has no {precision} kernels")
        if name not in cache:
            cache[name] = Engine(PRESETS[name], canonical(name), "cuda:0", precision=precision)
            assert cache[name].precision == precision
        return cache[name]
    return get


def _oracle(name, canonical):
    from oracle.oracle import Oracle
    return Oracle(PRESETS[name], canonical(name))


@pytest.mark.parametrize("name", ALL)
def test_streaming_matches_reference_golden(name, golden, engines, precision):
    """stream_out / stream_state of the golden files come from the reference's own streaming graph."""
    cfg, g, eng = PRESETS[name], golden(name), engines(name)
    x = synthetic_noisy(2, N_HOPS * cfg.hop_size, cfg.sample_rate)
    st = eng.new_state(2)
    y = eng.stream(st, torch.from_numpy(x).cuda()).cpu().numpy()
    assert y.shape == g["stream_out"].shape
    assert rms(y - g["stream_out"]) < TOL[precision]["wav"]
    assert np.abs(st.export().cpu().numpy() - g["stream_state"]).max() < TOL[precision]["state"]


@pytest.mark.parametrize("name", ALL)
def test_every_variant_matches_oracle(name, canonical, engines, precision):
    """every streams-per-CTA variant, ragged stream counts, state carried across launches."""
    cfg, eng, o = PRESETS[name], engines(name), _oracle(name, canonical)
    H = cfg.hop_size
    try:
        for S in (1, 2, 4):
            try:
                eng.set_streams_per_cta(S)
            except RuntimeError:
                continue
            B = 2 * S + 1
            x = synthetic_noisy(B, 6 * H, cfg.sample_rate, first_stream=7)
            ost = o.new_state(B)
            want = o.stream(ost, x)
            st = eng.new_state(B)
            xd = torch.from_numpy(x).cuda()
            got = torch.cat([eng.stream(st, xd[:, :2 * H]), eng.stream(st, xd[:, 2 * H:])], dim=1).cpu().numpy()
            assert rms(got - want) < TOL[precision]["wav"], (name, S)
            assert np.abs(st.export().cpu().numpy() - ost).max() < TOL[precision]["state"], (name, S)
    finally:
        eng.set_streams_per_cta(0)


@pytest.mark.parametrize("name,B,nh", [("16k_m", 331, 16), ("16k_l", 200, 9), ("48k_m", 170, 10), ("48k_l", 160, 8), ("16k_b", 640, 24)])
def test_hop_sliced_launch_is_bit_identical(name, B, nh, engines, precision):
    """More stream groups than SMs: fe_stream cuts the launch into hop ranges on a persistent grid (items wait for the previous range
    of their streams; state through global memory).  Output and final state must equal the unsliced launch bit for bit."""
    cfg, eng = PRESETS[name], engines(name)
    H = cfg.hop_size
    x = torch.from_numpy(synthetic_noisy(B, nh * H, cfg.sample_rate)).cuda()
    out = []
    try:
        for on in (False, True):
            eng.set_hop_slicing(on)
            st = eng.new_state(B)
            y = torch.cat([eng.stream(st, x[:, :(nh - 3) * H].contiguous()), eng.stream(st, x[:, (nh - 3) * H:].contiguous())], dim=1)
            out.append((y, st.export()))
    finally:
        eng.set_hop_slicing(True)
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])


def test_variant_choice_follows_the_round_model(canonical):
    """streams-per-CTA variant per batch size: rounds of one CTA per SM x the cost of an S-stream CTA (fe_api.cu::pick_variant)."""
    from fastenhancer_b200.engine import Engine
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    eng = Engine(PRESETS["16k_b"], canonical("16k_b"), "cuda:0")                    # fp32x3: S in {1, 2}
    assert [eng.streams_per_cta(n) for n in (1, sms, sms + 1, 2 * sms, 4096)] == [1, 1, 2, 2, 2]
    eng = Engine(PRESETS["16k_b"], canonical("16k_b"), "cuda:0", precision="f16")   # S in {1, 2, 4}
    assert [eng.streams_per_cta(n) for n in (sms, 2 * sms, 4096)] == [1, 2, 4]
    eng = Engine(PRESETS["16k_m"], canonical("16k_m"), "cuda:0", precision="bf16")  # S in {1, 2}, hop-sliced launches
    assert [eng.streams_per_cta(n) for n in (sms, 256, 512)] == [1, 2, 2]


@pytest.mark.parametrize("name", ["16k_t", "16k_b", "16k_m", "48k_l"])
def test_hop_by_hop_equals_one_launch_bit_exact(name, engines):
    """1 launch of n hops == n launches of 1 hop: the state round trip through HBM loses nothing."""
    cfg, eng = PRESETS[name], engines(name)
    H = cfg.hop_size
    x = torch.from_numpy(synthetic_noisy(3, 5 * H, cfg.sample_rate)).cuda()
    s1, s2 = eng.new_state(3), eng.new_state(3)
    y1 = eng.stream(s1, x)
    y2 = torch.cat([eng.stream(s2, x[:, i * H:(i + 1) * H]) for i in range(5)], dim=1)
    assert torch.equal(y1, y2) and torch.equal(s1.export(), s2.export())


@pytest.mark.parametrize("name", ALL)
def test_stage_taps(name, canonical, engines, precision):
    from oracle.oracle import tap_schema
    cfg, eng, o = PRESETS[name], engines(name), _oracle(name, canonical)
    x = synthetic_noisy(2, 5 * cfg.hop_size, cfg.sample_rate)
    _, ref = o.stream(o.new_state(2), x, taps=True)
    _, taps = eng.stream_taps(eng.new_state(2), torch.from_numpy(x).cuda(), 4)
    taps = taps.cpu().numpy()
    off = 0
    for nm, shp in tap_schema(cfg):
        n = int(np.prod(shp))
        r = ref[nm][4]
        assert np.abs(taps[off:off + n].reshape(shp) - r).max() < TOL[precision]["tap"] * max(1.0, np.abs(r).max()), nm
        off += n


@pytest.mark.parametrize("name", ALL)
def test_offline_matches_reference_golden(name, golden, engines, precision):
    cfg, g, eng = PRESETS[name], golden(name), engines(name)
    L = int(g["offline_len"])
    wav, spec = eng.offline(torch.from_numpy(synthetic_noisy(2, L, cfg.sample_rate)).cuda())
    wav, spec = wav.cpu().numpy(), spec.cpu().numpy()
    assert wav.shape == g["offline_wav"].shape == (2, cfg.hop_size * (L // cfg.hop_size))      # exact frame arithmetic
    assert spec.shape == (2, cfg.f_in, 1 + L // cfg.hop_size, 2)
    assert rms(wav - g["offline_wav"]) < TOL[precision]["wav"]
    if "offline_spec_frames" in g.files:
        spec = spec[:, :, g["offline_spec_frames"]]
    assert np.abs(spec - g["offline_spec"]).max() < TOL[precision]["spec_abs"] * max(1.0, np.abs(g["offline_spec"]).max())


@pytest.mark.parametrize("name", ALL)
def test_offline_schedules_agree_and_match_reference_golden(name, golden, canonical):
    """Model.forward through both schedules of fe_offline: the sequential walk (one CTA per group of utterances) and the
    frame-parallel schedule (stage A -> GRU scan -> stage B per block -> overlap-add).  Both must reproduce the reference's golden
    output; between themselves they differ by fp32 summation order only."""
    from fastenhancer_b200.engine import Engine
    cfg, g = PRESETS[name], golden(name)
    eng = Engine(cfg, canonical(name), "cuda:0", precision="fp32")
    L = int(g["offline_len"])
    x = torch.from_numpy(synthetic_noisy(2, L, cfg.sample_rate)).cuda()
    out = {}
    for mode in ("walk", "frame_parallel"):
        eng.set_offline_mode(mode)
        n0 = eng.kernel_launches
        wav, spec = eng.offline(x)
        out[mode] = (wav.cpu().numpy(), spec.cpu().numpy(), eng.kernel_launches - n0)
        assert rms(out[mode][0] - g["offline_wav"]) < TOL["fp32"]["wav"], mode
        sp = out[mode][1][:, :, g["offline_spec_frames"]] if "offline_spec_frames" in g.files else out[mode][1]
        assert np.abs(sp - g["offline_spec"]).max() < TOL["fp32"]["spec_abs"] * max(1.0, np.abs(g["offline_spec"]).max()), mode
    assert out["walk"][2] == 1 and out["frame_parallel"][2] == 2 * cfg.rf_blocks + 2   # launches: 1 vs stage A + (scan + stage B) per block + overlap-add
    assert rms(out["walk"][0] - out["frame_parallel"][0]) < 2e-6
    assert np.abs(out["walk"][1] - out["frame_parallel"][1]).max() < 2e-5 * max(1.0, np.abs(out["walk"][1]).max())


@pytest.mark.parametrize("name,B,seconds", [("16k_t", 1, 10.0), ("16k_b", 1, 10.0), ("16k_b", 3, 2.5), ("48k_b", 2, 1.0), ("16k_l", 1, 2.0)])
def test_offline_frame_parallel_long_utterance(name, B, seconds, canonical):
    """BASELINE config 1 (T, batch 1, one 10 s utterance through Model.forward) and friends on the frame-parallel schedule against the
    oracle: hundreds of sequential scan steps, more frame groups than SMs, ragged lengths, several utterances."""
    from fastenhancer_b200.engine import Engine
    cfg = PRESETS[name]
    eng = Engine(cfg, canonical(name), "cuda:0")                  # default precision: fp32x3 where it exists, else fp32
    eng.set_offline_mode("frame_parallel")
    L = int(seconds * cfg.sample_rate) + 77
    x = synthetic_noisy(B, L, cfg.sample_rate)
    w_ref, sp_ref = _oracle(name, canonical).offline(x)
    wav, spec = eng.offline(torch.from_numpy(x).cuda())
    assert rms(wav.cpu().numpy() - w_ref) < TOL["fp32"]["wav"]
    assert np.abs(spec.cpu().numpy() - sp_ref).max() < TOL["fp32"]["spec_abs"] * max(1.0, np.abs(sp_ref).max())
    eng.set_offline_mode("auto")                                  # few utterances: automatic mode picks the same schedule
    n0 = eng.kernel_launches
    wav2, _ = eng.offline(torch.from_numpy(x).cuda())
    assert eng.kernel_launches - n0 == 2 * cfg.rf_blocks + 2 and torch.equal(wav, wav2)


def test_offline_frame_parallel_needs_accurate_mode(canonical):
    from fastenhancer_b200.engine import Engine
    cfg = PRESETS["16k_b"]
    eng = Engine(cfg, canonical("16k_b"), "cuda:0", precision="f16")
    eng.set_offline_mode("frame_parallel")
    with pytest.raises(RuntimeError, match="fp32-accurate"):
        eng.offline(torch.zeros(1, 4000).cuda())
    eng.set_offline_mode("auto")                                  # reduced-precision modes keep the walk
    n0 = eng.kernel_launches
    eng.offline(torch.zeros(1, 4000).cuda())
    assert eng.kernel_launches - n0 == 1


@pytest.mark.parametrize("name", ALL)
def test_spec2spec_matches_reference_golden(name, golden, engines, precision):
    cfg, g, eng = PRESETS[name], golden(name), engines(name)
    st = eng.new_state(2)
    sp = torch.from_numpy(g["spec_in"]).cuda()
    out = torch.cat([eng.spec(st, sp[:, :, :3].contiguous()), eng.spec(st, sp[:, :, 3:6].contiguous())], dim=2).cpu().numpy()
    assert np.abs(out - g["spec_out"]).max() < TOL[precision]["spec"] * np.abs(g["spec_out"]).max()
    assert np.all(out[:, -1] == 0)                                   # Nyquist bin padded with zeros
    h = st.export().cpu().numpy()[:, 2 * cfg.cache_len:].reshape(g["spec_h"].shape)
    assert np.abs(h - g["spec_h"]).max() < TOL[precision]["state"]


def test_state_import_export_round_trip(engines):
    eng = engines("16k_b")
    st = eng.new_state(5)
    ref = torch.randn(5, eng.state_floats, device="cuda")
    st.load(ref)
    assert torch.equal(st.export(), ref)
    st.reset()
    assert float(st.export().abs().max()) == 0.0


def test_streaming_delay_is_nfft_minus_hop(canonical, engines):
    """bit-exact frame indexing: with an identity-like mask the streaming output is the input delayed by n_fft - hop
    (docs/docs/onnx.md:37-72); an impulse at sample p must peak at p + n_fft - hop, and nothing may come out before
    the hop that first sees it (the frame containing the impulse legitimately spreads inside its own n_fft window)."""
    cfg, eng = PRESETS["16k_b"], engines("16k_b")
    H, N = cfg.hop_size, cfg.n_fft
    x = np.zeros((1, 8 * H), np.float32)
    p = 3 * H + 17
    x[0, p] = 1.0
    y = eng.stream(eng.new_state(1), torch.from_numpy(x).cuda()).cpu().numpy()[0]
    want = _oracle("16k_b", canonical).stream(np.zeros((1, cfg.state_floats), np.float32), x)[0]
    assert int(np.argmax(np.abs(want))) == p + N - H
    assert int(np.argmax(np.abs(y))) == p + N - H
    assert np.all(y[:(p // H) * H] == 0.0) and np.all(want[:(p // H) * H] == 0.0)


def test_host_buffer_path_equals_device_path(engines):
    """fe_stream_host (pipelined H2D / kernel / D2H in pieces) == fe_stream on resident buffers, bit for bit."""
    cfg, eng = PRESETS["16k_b"], engines("16k_b")
    H = cfg.hop_size
    x = torch.from_numpy(synthetic_noisy(9, 37 * H, cfg.sample_rate)).pin_memory()
    y_dev = eng.stream(eng.new_state(9), x.cuda()).cpu()
    y_host = eng.stream_host(eng.new_state(9), x, hops_per_chunk=8)
    assert torch.equal(y_dev, y_host)


def test_full_size_properties(canonical, engines, precision):
    """BASELINE config 2 at full size (FastEnhancer_B, 256 streams, 10 s = 626 hops): sampled streams against
    the oracle, stream independence, and launch-split invariance."""
    cfg, eng, o = PRESETS["16k_b"], engines("16k_b"), _oracle("16k_b", canonical)
    H, B, nh = cfg.hop_size, 256, 626
    x = synthetic_noisy(B, nh * H, cfg.sample_rate)
    x[200] = x[3]                                                     # two streams with identical input
    xd = torch.from_numpy(x).cuda()
    st = eng.new_state(B)
    y = eng.stream(st, xd)
    pick = [0, 3, 131, 255]
    want = o.stream(o.new_state(len(pick)), x[pick])
    assert rms(y[pick].cpu().numpy() - want) < TOL[precision]["wav"]
    assert torch.equal(y[200], y[3])                                  # streams never mix
    st2 = eng.new_state(B)
    y2 = torch.cat([eng.stream(st2, xd[:, :300 * H]), eng.stream(st2, xd[:, 300 * H:])], dim=1)
    assert torch.equal(y, y2)
    assert rms(y.cpu().numpy()) > 0.05                                # mask ~ identity: output carries the signal


def test_model_classes_drop_in(golden):
    """Model / ONNXModel / StreamingModel (the reference-facing host classes) on the engine."""
    from fastenhancer_b200.model import Model, ONNXModel, StreamingModel
    cfg, g = PRESETS["16k_t"], golden("16k_t")
    m = Model(**cfg.to_model_kwargs()).eval().cuda()
    L = int(g["offline_len"])
    wav, spec = m(torch.from_numpy(synthetic_noisy(2, L, cfg.sample_rate)).cuda())
    assert rms(wav.cpu().numpy() - g["offline_wav"]) < 5e-5 and spec.shape[1:] == (cfg.f_in, 1 + L // cfg.hop_size, 2)
    om = ONNXModel(**cfg.to_model_kwargs()).eval().cuda()
    sp = torch.from_numpy(g["spec_in"]).cuda()
    o1, *h = om(sp[:, :, :3].contiguous())
    o2, *h = om(sp[:, :, 3:6].contiguous(), *h)
    out = torch.cat([o1, o2], dim=2).cpu().numpy()
    assert np.abs(out - g["spec_out"]).max() < 1e-3 * np.abs(g["spec_out"]).max()
    assert tuple(h[0].shape) == (1, 2 * cfg.rf_freq, cfg.rf_channels)
    # streaming graph with explicit caches, hop by hop, like scripts/test_onnx.py:44-49
    sm = StreamingModel(om)
    H = cfg.hop_size
    x = torch.from_numpy(synthetic_noisy(2, N_HOPS * H, cfg.sample_rate)).cuda()
    caches = sm.initialize_cache(x[:, :H])
    hops = []
    for i in range(N_HOPS):
        y, *caches = sm(x[:, i * H:(i + 1) * H], *caches)
        hops.append(y)
    assert rms(torch.cat(hops, dim=1).cpu().numpy() - g["stream_out"]) < 5e-5
    assert rms(sm.run(x).cpu().numpy() - g["stream_out"]) < 5e-5


def test_reference_style_composition_with_stft_shims(golden):
    """The streaming graph composed the reference's way (scripts/export_onnx.py:53-57): stft -> model -> stft.inverse,
    explicit caches, hop by hop -- three kernel launches per hop through ONNXModel.stft / ONNXModel."""
    from fastenhancer_b200.model import ONNXModel
    for name in ("16k_t", "16k_m"):
        cfg, g = PRESETS[name], golden(name)
        m = ONNXModel(**cfg.to_model_kwargs()).eval().cuda()
        H = cfg.hop_size
        x = torch.from_numpy(synthetic_noisy(2, N_HOPS * H, cfg.sample_rate)).cuda()
        c_stft, c_istft = m.stft.initialize_cache(x[:, :H])
        hs = m.initialize_cache(x[:, :H])
        hops = []
        for i in range(N_HOPS):
            spec_in, c_stft = m.stft(x[:, i * H:(i + 1) * H], c_stft)
            assert spec_in.shape == (2, cfg.n_fft // 2 + 1, 1, 2)
            spec_out, *hs = m(spec_in, *hs)
            y, c_istft = m.stft.inverse(spec_out, c_istft)
            hops.append(y)
        assert rms(torch.cat(hops, dim=1).cpu().numpy() - g["stream_out"]) < 5e-5, name
        # the front end alone against torch.fft on the same frames
        ref = torch.fft.rfft(torch.cat([torch.zeros(2, cfg.n_fft - H, device="cuda"), x[:, :H]], dim=1) * m.stft.window.cuda(), dim=1)
        got, _ = m.stft(x[:, :H], None)
        assert (torch.view_as_complex(got[:, :, 0].contiguous()) - ref).abs().max() < 1e-5 * ref.abs().max()


def _net_mask_canonical(name, canonical):
    """canonical weights of the seed-0 checkpoint with a NETWORK-DOMINATED mask: zero bias of the final transposed conv
    (instead of [1, 0]) and its weights x 15, so the mask is the network output alone and precision loss in the network
    is not attenuated in the waveform."""
    from fastenhancer_b200.schema import canonical_schema
    cfg = PRESETS[name]
    w = canonical(name).copy()
    off = 0
    for nm, shp in canonical_schema(cfg):
        n = int(np.prod(shp))
        if nm == "dec_post.wt":
            w[off:off + n] *= 15.0
        if nm == "dec_post.bt":
            w[off:off + n] = 0.0
        off += n
    return w


@pytest.mark.parametrize("name", ["16k_b", "16k_m"])
def test_network_dominated_mask(name, canonical, precision):
    from fastenhancer_b200.engine import Engine
    from oracle.oracle import Oracle
    if name not in HAS.get(precision, PRESETS):
        pytest.skip(f"{name} has no {precision} kernels")
    cfg = PRESETS[name]
    w = _net_mask_canonical(name, canonical)
    x = synthetic_noisy(4, 40 * cfg.hop_size, cfg.sample_rate)
    want = Oracle(cfg, w).stream(np.zeros((4, cfg.state_floats), np.float32), x)
    eng = Engine(cfg, w, "cuda:0", precision=precision)
    got = eng.stream(eng.new_state(4), torch.from_numpy(x).cuda()).cpu().numpy()
    rel = rms(got - want) / rms(want)
    print(f"network-dominated mask, {name} {precision}: output rms {rms(want):.3f}, relative rms error {rel:.2e}")
    assert rms(want) > 0.02                      # the mask really is the network
    assert rel < TOL_NET[precision]


#: BASELINE configs at their real horizon: B 626 hops, M 1 003, L 1 605, 48 kHz L 2 405
LONG = [("16k_b", 626), ("16k_m", 1003), ("16k_l", 1605), ("48k_l", 2405)]
_LONG_CACHE = {}


@pytest.mark.parametrize("name,n_hops", LONG)
def test_long_horizon_against_oracle(name, n_hops, canonical, engines, precision):
    """the GRU recurrence over the whole utterance: waveform and final state against the oracle, no drift in any mode."""
    cfg, eng = PRESETS[name], engines(name)
    B = 2 if cfg.channels < 128 else 1                          # the L configs cost the CPU oracle ~1 min per stream
    x = synthetic_noisy(B, n_hops * cfg.hop_size, cfg.sample_rate, first_stream=11)
    if name not in _LONG_CACHE:                                 # one oracle run serves every precision
        o = _oracle(name, canonical)
        ost = o.new_state(B)
        _LONG_CACHE[name] = (o.stream(ost, x, n_threads=B), ost)
    want, ost = _LONG_CACHE[name]
    st = eng.new_state(B)
    got = eng.stream(st, torch.from_numpy(x).cuda()).cpu().numpy()
    assert rms(got - want) < TOL[precision]["wav"]
    tail = slice((n_hops - 50) * cfg.hop_size, None)          # the error is not growing: the last 50 hops on their own
    assert rms(got[:, tail] - want[:, tail]) < TOL[precision]["wav"]
    assert np.abs(st.export().cpu().numpy() - ost).max() < TOL[precision]["state"]


@pytest.mark.parametrize("name", ["16k_b", "16k_m"])
def test_long_horizon_against_reference_golden(name, golden, engines, precision):
    """>= 600 hops of the REFERENCE's own streaming graph (tools/gen_golden.py --long): the tail of the waveform and the
    final state, so that the oracle itself is pinned over a long recurrence."""
    g = golden(name + "_long")
    cfg, eng = PRESETS[name], engines(name)
    n_hops, keep = int(g["n_hops"]), int(g["keep_hops"])
    x = synthetic_noisy(1, n_hops * cfg.hop_size, cfg.sample_rate, first_stream=5)
    st = eng.new_state(1)
    y = eng.stream(st, torch.from_numpy(x).cuda()).cpu().numpy()
    assert rms(y[:, -keep * cfg.hop_size:] - g["stream_tail"]) < TOL[precision]["wav"]
    assert np.abs(st.export().cpu().numpy() - g["stream_state"]).max() < TOL[precision]["state"]


def test_two_states_stream_host_concurrently(engines):
    """fe_stream_host keeps its staging buffers / streams in the state: two states of one engine driven from two host
    threads give the same bits as one after the other."""
    import threading
    cfg, eng = PRESETS["16k_b"], engines("16k_b")
    H = cfg.hop_size
    xs = [torch.from_numpy(synthetic_noisy(8, 64 * H, cfg.sample_rate, first_stream=8 * i)).pin_memory() for i in range(2)]
    want = [eng.stream(eng.new_state(8), x.cuda()).cpu() for x in xs]
    states = [eng.new_state(8) for _ in range(2)]
    outs = [None, None]

    def work(i):
        outs[i] = eng.stream_host(states[i], xs[i], hops_per_chunk=4)
    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert torch.equal(outs[0], want[0]) and torch.equal(outs[1], want[1])


@pytest.mark.parametrize("name", ["16k_t", "16k_b", "48k_b"])
def test_hop_tiles_by_tma_equal_plain_loads(name, engines):
    """Streaming launches of the hop-tiled variants move the input / output hop as 2-D TMA tiles (cp.async.bulk.tensor, box
    [S streams][hop tile]) when the caller's arrays are 16-byte aligned and pitched; otherwise they use plain loads / stores.
    Same bits either way, ragged last CTA included (out-of-bounds rows of the box are zero-filled / clipped by the TMA unit)."""
    cfg, eng = PRESETS[name], engines(name)
    H, B, nh = cfg.hop_size, 5, 7
    x = torch.from_numpy(synthetic_noisy(B, nh * H, cfg.sample_rate)).cuda()
    y_tma = eng.stream(eng.new_state(B), x)
    pad = torch.zeros(B, nh * H + 1, device="cuda")
    pad[:, 1:] = x                                            # rows start 4 bytes off a 16-byte boundary, pitch not a multiple of 4 floats
    out = torch.empty(B, nh * H + 1, device="cuda")
    y_plain = eng.stream(eng.new_state(B), pad[:, 1:], out=out[:, 1:])
    assert torch.equal(y_tma, y_plain)
