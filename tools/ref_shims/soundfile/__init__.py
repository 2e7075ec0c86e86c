"""soundfile stand-in: `write` (float WAV via scipy) is all the reference's inference scripts use."""
import numpy as np


def write(file, data, samplerate, subtype=None, **_):
    from scipy.io import wavfile
    wavfile.write(str(file), int(samplerate), np.asarray(data, dtype=np.float32))


def read(file, dtype="float32", **_):
    from scipy.io import wavfile
    fs, x = wavfile.read(str(file))
    if x.dtype.kind == "i":
        x = x.astype(np.float64) / float(np.iinfo(x.dtype).max + 1)
    return x.astype(dtype), fs
