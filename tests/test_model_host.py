"""Host-side mirror of the reference's Model / ONNXModel interface (no GPU needed): parameter names and shapes
are the reference's, reference-style model_kwargs are accepted, and using the model without a GPU fails loudly."""
import numpy as np
import pytest
import torch

from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.model import Model, ONNXModel, StreamingModel
from fastenhancer_b200.schema import state_dict_schema, synthetic_state_dict


@pytest.mark.parametrize("name", ["16k_t", "16k_b", "48k_l"])
def test_state_dict_uses_reference_names(name, golden):
    cfg = PRESETS[name]
    m = Model(**cfg.to_model_kwargs())
    sd = m.state_dict()
    want = [n for n, _, _ in state_dict_schema(cfg)]
    assert list(sd.keys()) == want
    for n, shape, kind in state_dict_schema(cfg):
        assert tuple(sd[n].shape) == tuple(shape), n
    # strict load of a (synthetic) reference-format checkpoint
    ck = {k: torch.from_numpy(np.array(v)) for k, v in synthetic_state_dict(cfg, seed=3).items()}
    m.load_state_dict(ck, strict=True)
    np.testing.assert_array_equal(m.state_dict()["dec_post.3.bias"].numpy(), [1.0, 0.0])
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in ck.items() if k != "enc_pre.0.weight"}, strict=True)


def test_reference_surface():
    cfg = PRESETS["16k_b"]
    m = ONNXModel(**cfg.to_model_kwargs()).eval()
    assert (m.stft.n_fft, m.stft.hop_size) == (512, 256) and m.stft.window.shape == (512,)
    x = torch.zeros(2, 256)
    c = m.stft.initialize_cache(x)
    assert [tuple(t.shape) for t in c] == [(2, 256), (2, 256)]
    h = m.initialize_cache(x)
    assert len(h) == cfg.rf_blocks and tuple(h[0].shape) == (1, 2 * cfg.rf_freq, cfg.rf_channels)
    m.flatten_parameters()
    s = StreamingModel(m)
    assert len(s.initialize_cache(x)) == 2 + cfg.rf_blocks
    assert m.canonical_weights().dtype == np.float32


def test_no_gpu_is_a_loud_error():
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is visible")
    m = Model(**PRESETS["16k_t"].to_model_kwargs())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4000))


def test_unsupported_kwargs_rejected():
    kw = PRESETS["16k_t"].to_model_kwargs()
    kw["mask"] = "sigmoid"
    with pytest.raises(ValueError):
        Model(**kw)


def _makespan(n_groups, n_hops, num_sms, slice_hops):
    """rounds of a full chain that a launch takes: items (range, group) dealt round-robin to min(SMs, items) persistent CTAs, an item
    waits for the previous range of its group (the schedule of fe_kernel.cuh::Frame::run)."""
    if slice_hops == 0:
        return float(-(-n_groups // num_sms))
    nr = -(-n_hops // slice_hops)
    total, G = n_groups * nr, min(num_sms, n_groups * nr)
    fin, cta = [0.0] * total, [0.0] * G
    for i in range(total):
        r = i // n_groups
        start = max(cta[i % G], fin[i - n_groups] if r else 0.0)
        fin[i] = start + min(slice_hops, n_hops - r * slice_hops) / n_hops
        cta[i % G] = fin[i]
    return max(fin)


def test_hop_slice_plan_shortens_multi_round_launches():
    """fe_plan_hop_slices (host arithmetic of the C ABI, no device): nothing to gain with at most one stream group per SM (the chains are
    the critical path) or on short launches; multi-round launches get ranges of at least 4 hops that shorten the simulated schedule."""
    from fastenhancer_b200.engine import load_library
    lib = load_library()
    assert lib.fe_plan_hop_slices(128, 626, 148) == 0 and lib.fe_plan_hop_slices(148, 626, 148) == 0      # one round either way
    assert lib.fe_plan_hop_slices(256, 4, 148) == 0                                                        # too short to slice
    for groups, hops in ((256, 626), (256, 64), (512, 1003), (200, 2405), (1024, 100)):
        h = lib.fe_plan_hop_slices(groups, hops, 148)
        assert h == 0 or h >= 4
        if h:
            assert _makespan(groups, hops, 148, h) < 0.97 * _makespan(groups, hops, 148, 0), (groups, hops, h)
    assert lib.fe_plan_hop_slices(256, 626, 148) > 0 and lib.fe_plan_hop_slices(512, 1003, 148) > 0
