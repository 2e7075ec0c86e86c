"""The reference's OWN scripts and plumbing, unmodified, against the engine.

tools/stage_reference.py copies /root/reference to git-ignored baseline/_ref/reference (plus the overlay package
models/fastenhancer_b200 a maintainer would drop in, stand-ins for the third-party packages missing from the image, and seeded
checkpoints in the reference's `logs/<name>/{config.yaml,NNNNN.pth}` layout); the copy travels to the GPU box with the repo
snapshot.  Everything here is skipped where that tree is absent."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, ".staged")), reason="baseline/_ref not staged (tools/stage_reference.py)")


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(REF, "reference"), os.path.join(REF, "shims"), ROOT, env.get("PYTHONPATH", "")])
    return env


def test_reference_plumbing_builds_and_loads_the_overlay_model():
    """get_hparams -> get_wrapper -> wrapper.load() (scripts/test_pytorch.py:20-23, wrappers/ns.py:29-32, 308-321) with
    `model: fastenhancer_b200` as the only changed key: the reference's own loader fills our Model, strict."""
    code = ("from utils import get_hparams; from wrappers import get_wrapper\n"
            "hps = get_hparams(base_dir='logs/16k_b_b200')\n"
            "w = get_wrapper(hps.wrapper)(hps, device='cpu'); w.load()\n"
            "import fastenhancer_b200.model as m\n"
            "assert type(w.model) is m.Model and w.epoch == 1 and w.model.stft.hop_size == 256\n"
            "print('OK', type(w.model).__module__)\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=REF, env=_env(), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK fastenhancer_b200.model" in r.stdout, r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["16k_t", "16k_b"])
def test_test_pytorch_script_unmodified(preset, tmp_path):
    """scripts/test_pytorch.py, byte for byte, once with the reference model and once with `model: fastenhancer_b200`:
    same WAV files in, enhanced WAV files out, <= 1e-5 RMS apart (the drop-in default is the fp32-accurate arithmetic)."""
    from scipy.io import wavfile
    script = os.path.join(REF, "reference", "scripts", "test_pytorch.py")
    outs = {}
    for which in ("ref", "b200"):
        out = tmp_path / which
        r = subprocess.run([sys.executable, script, "-n", f"{preset}_{which}", "-i", os.path.join(REF, "wavs"), "-o", str(out)],
                           cwd=REF, env=_env(), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        outs[which] = {f: wavfile.read(str(out / f))[1] for f in sorted(os.listdir(out))}
    assert sorted(outs["ref"]) == sorted(outs["b200"]) == ["noisy_0.wav", "noisy_1.wav"]
    for f in outs["ref"]:
        a, b = outs["ref"][f].astype(np.float64), outs["b200"][f].astype(np.float64)
        assert a.shape == b.shape and a.std() > 0.02
        assert np.sqrt(np.mean((a - b) ** 2)) < 1e-5, f
