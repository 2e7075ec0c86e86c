"""librosa stand-in: `load` (WAV via scipy, mono, polyphase resampling) is all the reference's inference scripts use."""
from math import gcd

import numpy as np

from . import filters  # noqa: F401


def load(path, sr=22050, mono=True, dtype=np.float32, **_):
    from scipy.io import wavfile
    from scipy.signal import resample_poly
    fs, x = wavfile.read(str(path))
    if x.dtype.kind == "i":
        x = x.astype(np.float64) / float(np.iinfo(x.dtype).max + 1)
    elif x.dtype.kind == "u":
        x = (x.astype(np.float64) - 128.0) / 128.0
    x = x.astype(np.float64)
    if x.ndim == 2:
        x = x.mean(axis=1) if mono else x.T
    if sr is not None and sr != fs:
        g = gcd(int(sr), int(fs))
        x = resample_poly(x, int(sr) // g, int(fs) // g, axis=-1)
        fs = sr
    return x.astype(dtype), fs
