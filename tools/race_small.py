import sys, os
sys.path.insert(0, "/root/repo")
import torch
from fastenhancer_b200.config import PRESETS
from fastenhancer_b200.engine import Engine
from fastenhancer_b200.fold import fold_to_canonical
from fastenhancer_b200.schema import synthetic_state_dict
from fastenhancer_b200.synth import synthetic_noisy
cfg = PRESETS["16k_t"]
eng = Engine(cfg, fold_to_canonical(cfg, synthetic_state_dict(cfg, 0)), "cuda:0")
x = torch.from_numpy(synthetic_noisy(1, 16 * cfg.hop_size + 5, cfg.sample_rate)).cuda()
eng.set_offline_mode("frame_parallel"); a = eng.offline(x)[0]
eng.set_offline_mode("walk"); b = eng.offline(x)[0]
print("offline tp vs walk", float((a - b).abs().max()))
s = eng.stft_gemm(torch.from_numpy(synthetic_noisy(2, 3 * cfg.hop_size + cfg.n_fft, cfg.sample_rate)).cuda())
print("stft_gemm", tuple(s.shape), float(s.abs().max()))
