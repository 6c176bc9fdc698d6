"""l3embedding/audio.py:4-31 -- host-side pcm2float kept for callers that scale on the CPU.  The hot path feeds
int16 straight to the device, where the front-end kernel applies the same (x - 0) / 32768 scaling."""
import numpy as np


def pcm2float(sig, dtype='float64'):
    """Convert PCM signal to floating point with a range from -1 to 1 (same contract as the reference)."""
    sig = np.asarray(sig)
    if sig.dtype.kind not in 'iu':
        raise TypeError("'sig' must be an array of integers")
    dtype = np.dtype(dtype)
    if dtype.kind != 'f':
        raise TypeError("'dtype' must be a floating point type")
    i = np.iinfo(sig.dtype)
    abs_max = 2 ** (i.bits - 1)
    offset = i.min + abs_max
    return (sig.astype(dtype) - offset) / abs_max
