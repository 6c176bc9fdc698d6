"""Weight inventory and checkpoint I/O in the reference's saved-weight layout.

The reference saves with keras `save_weights` (HDF5; l3embedding/train.py:316-355) and loads with `load_weights`
(l3embedding/model.py:119).  h5py does not exist in the build environment, so the canonical interchange format
here is an ordered .npz whose entries are the arrays of keras `Model.get_weights()` in that order, stored under
the keras-style names below.  Paths ending in .h5 / .hdf5 are written in the keras 2.0.9 HDF5 layout (one group per
top-level layer, `weight_names` attributes) -- through h5py when it is importable, otherwise through the package's own
minimal HDF5 writer (minihdf5.py; its reader is checked against a libhdf5-written file, the writer by round trip).
Files are recognised by magic bytes, not by extension.
"""
from __future__ import annotations

import math
import zipfile
from functools import lru_cache
from typing import Dict

import numpy as np

from . import _lib

SR = 48000


@lru_cache(maxsize=None)
def weight_shapes(model_type: str) -> Dict[str, tuple]:
    return {name: shape for name, _a, _o, shape in _lib.tensor_table(model_type)}


def he_normal_weights(model_type: str, seed: int = 20180123) -> Dict[str, np.ndarray]:
    """Fresh weights as the reference builders create them: he_normal kernels (truncated normal, std sqrt(2/fan_in);
    l3embedding/audio_model.py:376-432, vision_model.py:130-186, model.py:26-31), zero biases, BN gamma 1 / beta 0 /
    moving mean 0 / moving variance 1."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in weight_shapes(model_type).items():
        leaf = name.rsplit("/", 1)[1]
        if leaf == "kernel":
            std = math.sqrt(2.0 / float(np.prod(shape[:-1])))
            x = rng.standard_normal(shape)
            bad = np.abs(x) > 2.0
            while bad.any():
                x[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(x) > 2.0
            out[name] = (x * std).astype(np.float32)
        elif leaf in ("gamma", "moving_variance"):
            out[name] = np.ones(shape, np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


def mel_filterbank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """librosa 0.5.1 filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=True, norm=1): (n_mels, 1+n_fft/2)."""
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mels = np.linspace(0.0, 2595.0 * np.log10(1.0 + (sr / 2.0) / 700.0), n_mels + 2)
    mel_f = 700.0 * (10.0 ** (mels / 2595.0) - 1.0)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        w[i] = np.maximum(0.0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    return w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]


@lru_cache(maxsize=4)
def kapre_constants(model_type: str) -> Dict[str, np.ndarray]:
    """The non-trainable arrays the kapre (Mel)Spectrogram layer holds in a keras checkpoint (real/imag DFT
    kernels (n_dft,1,1,n_freq) and freq2mel (n_freq,n_mels)).  The device front-end computes the same transform
    with an FFT and does not read them; they exist so get_weights()/save_weights keep the reference's layout."""
    from .model import AUDIO_FRONTEND
    fe = AUDIO_FRONTEND[model_type]
    n = fe["n_dft"]
    nf = n // 2 + 1
    t = np.arange(n, dtype=np.float64)[:, None]
    k = np.arange(nf, dtype=np.float64)[None, :]
    win = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n))[:, None]
    ang = 2.0 * np.pi * k * t / n
    out = {"kapre/real_kernels": (np.cos(ang) * win).astype(np.float32).reshape(n, 1, 1, nf),
           "kapre/imag_kernels": (-np.sin(ang) * win).astype(np.float32).reshape(n, 1, 1, nf)}
    if fe["n_mels"]:
        out["kapre/freq2mel"] = mel_filterbank(SR, n, fe["n_mels"]).T.astype(np.float32)
    return out


# ---- file formats ----------------------------------------------------------------------------------------------

def _is_hdf5(path) -> bool:
    with open(path, "rb") as f:
        return f.read(8) == b"\x89HDF\r\n\x1a\n"


def save_weights(path, model):
    """keras `Model.save_weights`: a path ending in .h5 / .hdf5 gets the keras 2.0.9 HDF5 layout (through h5py when it is
    installed, else through the built-in writer in minihdf5.py); anything else the ordered .npz."""
    names = model.weight_names()
    arrays = model.get_weights()
    if str(path).endswith((".h5", ".hdf5")):
        return _save_h5(path, model, names, arrays)
    payload = {"__order__": np.array(names), "__model_type__": np.array(model.model_type)}
    for n, a in zip(names, arrays):
        payload[n] = a
    with open(path, "wb") as f:   # a file object keeps numpy from appending '.npz'
        np.savez(f, **payload)


def load_weights(path, model):
    if _is_hdf5(path):
        return _load_h5(path, model)
    if not zipfile.is_zipfile(path):
        raise ValueError("%s is neither a keras HDF5 file nor an l3embedding_b200 .npz checkpoint" % path)
    with np.load(path, allow_pickle=False) as z:
        names = [str(n) for n in z["__order__"]]
        if names != model.weight_names():
            raise ValueError("checkpoint %s holds the weights of a different model layout" % path)
        model.set_weights([z[n] for n in names])


def _keras_name(cn: str) -> str:
    """canonical 'vision/conv1a/kernel' -> keras variable-style 'vision_conv1a/kernel:0' (the slash nests a sub-group)"""
    parts = cn.split("/")
    return "%s/%s:0" % ("_".join(parts[:-1]), parts[-1])


def _keras_groups(model):
    """[(top-level layer name, [(keras weight name, canonical name)])] as keras 2.0.9 save_weights groups them: EVERY
    layer of `model.layers` gets a group (weight-less ones with an empty `weight_names`), datasets are named after the
    backend variables ('<layer>/<weight>:0').

    A model wrapped by multi_gpu_model (num_gpus > 1) is saved the way keras saves the reference's wrapper
    (training_utils.py:121-170): inputs, one Lambda slice per input and replica, the WHOLE template model as one nested
    layer holding every array in Container order (all trainable, then all non-trainable), and the output
    concatenation -- the layout `load_model(..., src_num_gpus=N)` expects (model.py:117-119)."""
    if getattr(model, "num_gpus", 0) > 1:
        groups = [("input_1", []), ("input_2", [])]
        groups += [("lambda_%d" % (i + 1), []) for i in range(2 * model.num_gpus)]
        groups.append((model.name, [(_keras_name(cn), cn) for cn in model.container_weight_names()]))
        groups.append(("dense_2", []))     # training_utils.py:168-170: the merged output carries the output's name
        return groups
    return [(layer.name, [(_keras_name(cn), cn) for cn in layer._weight_names]) for layer in model.layers]


def _save_h5(path, model, names, arrays):
    by_name = dict(zip(names, arrays))
    groups = _keras_groups(model)
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:   # not exercised in the build environment (no h5py there)
        with h5py.File(path, "w") as f:
            f.attrs["layer_names"] = [g.encode("utf8") for g, _ in groups]
            f.attrs["backend"] = b"tensorflow"
            f.attrs["keras_version"] = b"2.0.9"
            for gname, entries in groups:
                g = f.create_group(gname)
                g.attrs["weight_names"] = [kn.encode("utf8") for kn, _ in entries]
                for kn, cn in entries:
                    g.create_dataset(kn, data=by_name[cn])
        return
    from . import minihdf5
    tree = {}
    for gname, entries in groups:
        wn = np.array([kn.encode("utf8") for kn, _ in entries]) if entries else np.zeros((0,), "S1")
        node = {"__attrs__": {"weight_names": wn}}
        for kn, cn in entries:
            sub = node
            parts = kn.split("/")
            for part in parts[:-1]:
                sub = sub.setdefault(part, {})
            sub[parts[-1]] = np.asarray(by_name[cn])
        tree[gname] = node
    minihdf5.write_tree(path, tree, attrs={"layer_names": np.array([g.encode("utf8") for g, _ in groups]),
                                           "backend": np.bytes_(b"tensorflow"), "keras_version": np.bytes_(b"2.0.9")})


def _load_h5(path, model):
    """keras `load_weights` (topological): arrays are taken group by group in `layer_names` order and, inside a group,
    in `weight_names` order -- names themselves are not compared, so real keras checkpoints (variables called
    conv2d_7/kernel:0 ...) and the files written here load alike."""
    try:
        import h5py
        f = h5py.File(path, "r")
    except ImportError:
        from . import minihdf5
        f = minihdf5.File(path)
    dec = lambda n: n.decode("utf8") if isinstance(n, bytes) else str(n)
    root = f["model_weights"] if ("layer_names" not in f.attrs and "model_weights" in f) else f   # Model.save() files
    groups = []
    for ln in [dec(n) for n in root.attrs["layer_names"]]:
        g = root[ln]
        arrs = [np.asarray(g[dec(wn)]) for wn in g.attrs["weight_names"]]
        if arrs:
            groups.append((ln, arrs))
    if hasattr(f, "close"):
        f.close()
    arrays = [a for _, arrs in groups for a in arrs]
    expected = model.weight_names()
    if len(arrays) != len(expected):
        raise ValueError("checkpoint %s holds the weights of a different model layout (%d arrays, the %s model has %d)"
                         % (path, len(arrays), model.model_type, len(expected)))
    # Which layout?  single-GPU files have one group per weighted top-level layer (vision_model, audio_model, dense_1,
    # dense_2); a file saved from a multi_gpu_model has ONE weight-bearing group -- the nested template model -- with
    # every array in Container order (all trainable, then all non-trainable).  Decided by structure and confirmed by
    # shapes, so a wrong src_num_gpus cannot load arrays into the wrong tensors silently.
    shapes = dict(weight_shapes(model.model_type))
    shapes.update({k: v.shape for k, v in kapre_constants(model.model_type).items()})

    def fits(order):
        return all(tuple(a.shape) == tuple(shapes[n]) for a, n in zip(arrays, order))
    nested = len(groups) == 1
    order = model.container_weight_names() if nested else expected
    if not fits(order):
        other = expected if nested else model.container_weight_names()
        if not fits(other):
            raise ValueError("checkpoint %s: array shapes match neither the single-model layout nor the multi-GPU "
                             "(nested template model) layout of %s" % (path, model.model_type))
        order = other
    by_name = dict(zip(order, arrays))
    model.set_weights([by_name[n] for n in expected])
