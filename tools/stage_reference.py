#!/usr/bin/env python
"""Stage the UNMODIFIED reference under git-ignored baseline/_ref/ so that it travels to the GPU box with the repo snapshot.

    baseline/_ref/reference/   copy of /root/reference (sources, configs, onnx/*.wav fixtures; no .git, no docs), plus the overlay
                               package models/fastenhancer_b200/ a maintainer would drop in (integration/models/fastenhancer_b200)
    baseline/_ref/shims/       tools/ref_shims (librosa / soundfile / ... stand-ins for packages missing from the image)
    baseline/_ref/logs/<name>/ config.yaml + 00001.pth: seeded synthetic checkpoints in the reference's own format
                               (wrappers/ns.py:288-321), <name> = <preset>_ref (model: fastenhancer.default) and
                               <preset>_b200 (model: fastenhancer_b200 -- the only changed key)
    baseline/_ref/wavs/        two short 16 kHz WAVs (synthetic noisy input) for the directory-level script

Used by: bench.py --impl reference (times the reference's own PyTorch path), tests/test_reference_scripts.py (runs
scripts/test_pytorch.py unmodified).  Nothing under baseline/_ref is tracked or imported by the product.
Runs only where /root/reference is mounted (the build container); called from __graft_entry__.build().
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("FE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PRESET_YAML = {"16k_t": "configs/fastenhancer/t.yaml", "16k_b": "configs/fastenhancer/b.yaml", "16k_m": "configs/fastenhancer/m.yaml"}


def stage(force: bool = False) -> str:
    if not os.path.isdir(os.path.join(REF_SRC, "models")):
        raise RuntimeError(f"{REF_SRC} is not mounted: nothing to stage")
    ref = os.path.join(DST, "reference")
    marker = os.path.join(DST, ".staged")
    if os.path.exists(marker) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(REF_SRC, ref, ignore=shutil.ignore_patterns(".git", "docs", "__pycache__", "*.pyc", "assets"))
    shutil.copytree(os.path.join(ROOT, "tools", "ref_shims"), os.path.join(DST, "shims"), ignore=shutil.ignore_patterns("__pycache__"))
    # the overlay a maintainer drops into <reference>/models/ 
    shutil.copytree(os.path.join(ROOT, "integration", "models", "fastenhancer_b200"), os.path.join(ref, "models", "fastenhancer_b200"),
                    ignore=shutil.ignore_patterns("__pycache__"))
    # seeded checkpoints + configs in the reference's own layout
    import numpy as np
    import torch
    import yaml
    sys.path.insert(0, ROOT)
    from fastenhancer_b200.config import PRESETS
    from fastenhancer_b200.schema import synthetic_state_dict
    from fastenhancer_b200.synth import synthetic_noisy
    for preset, ypath in PRESET_YAML.items():
        cfg = PRESETS[preset]
        hps = yaml.safe_load(open(os.path.join(REF_SRC, ypath)))
        sd = {k: torch.from_numpy(np.array(v)) for k, v in synthetic_state_dict(cfg, seed=0).items()}
        for suffix, model in (("ref", hps["model"]), ("b200", "fastenhancer_b200")):
            d = os.path.join(DST, "logs", f"{preset}_{suffix}")
            os.makedirs(d)
            h2 = dict(hps)
            h2["model"] = model
            yaml.safe_dump(h2, open(os.path.join(d, "config.yaml"), "w"))
            torch.save({"model": sd, "epoch": 1}, os.path.join(d, "00001.pth"))
    from scipy.io import wavfile
    os.makedirs(os.path.join(DST, "wavs"))
    x = synthetic_noisy(2, 3 * 16000 + 123, 16000)
    for i in range(2):
        wavfile.write(os.path.join(DST, "wavs", f"noisy_{i}.wav"), 16000, (np.clip(x[i], -1, 1) * 32767).astype(np.int16))
    open(marker, "w").write("staged from " + REF_SRC + "\n")
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
