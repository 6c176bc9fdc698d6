"""data/usc/features.py:256-323 -- L3 embedding extraction for one audio file, on the B200 path.

`get_l3_frames_uniform(audio, l3embedding_model, hop_size=0.1, sr=48000)` keeps the reference's signature and its
framing rule, including the operator-precedence quirk at features.py:288-289 that leaves clips longer than one
second unpadded (the trailing partial hop is dropped).  With an l3embedding_b200 EmbeddingModel the overlapping 1 s
windows are read in place on the device (no 10x framed copy); any other object with `.predict` gets the framed
array, exactly like the reference.
"""
import numpy as np


def load_audio(path, sr):
    """features.py:18-28 (soundfile + resampy); those packages are optional here."""
    import soundfile as sf
    import resampy
    data, sr_orig = sf.read(path, dtype='float32', always_2d=True)
    data = data.mean(axis=-1)
    if sr_orig != sr:
        data = resampy.resample(data, sr_orig, sr)
    return data


def frame_signal(audio, hop_size=0.1, sr=48000):
    """The padding + framing arithmetic of get_l3_frames_uniform: returns (padded audio, hop_length, n_frames)."""
    hop_length = int(hop_size * sr)
    frame_length = sr * 1
    audio_length = len(audio)
    if audio_length < frame_length:
        pad_length = frame_length - audio_length          # make sure there is at least one frame
    else:
        # reference: int(np.ceil(audio_length - frame_length)/hop_length) * hop_length - (audio_length - frame_length)
        # -- np.ceil binds to the integer difference, so this is <= 0 and long clips are never padded
        pad_length = int(np.ceil(audio_length - frame_length) / hop_length) * hop_length - (audio_length - frame_length)
    if pad_length > 0:
        left_pad = pad_length // 2
        audio = np.pad(audio, (left_pad, pad_length - left_pad), mode='constant')
    n_frames = 1 + (len(audio) - frame_length) // hop_length
    return audio, hop_length, n_frames


def get_l3_frames_uniform(audio, l3embedding_model, hop_size=0.1, sr=48000):
    """Get L3 embedding for each frame in the given audio file (np.ndarray or path)."""
    if type(audio) == str:
        audio = load_audio(audio, sr)
    audio, hop_length, n_frames = frame_signal(np.asarray(audio), hop_size, sr)
    if hasattr(l3embedding_model, "predict_frames") and sr == 48000:
        return l3embedding_model.predict_frames(audio, hop_length)
    idx = np.arange(sr)[None, :] + hop_length * np.arange(n_frames)[:, None]
    x = audio[idx].reshape((n_frames, 1, sr))
    return l3embedding_model.predict(x)


def compute_file_features(path, feature_type, l3embedding_model=None, **feature_args):
    if feature_type == 'l3':
        if not l3embedding_model:
            err_msg = 'Must provide L3 embedding model to use {} features'
            raise ValueError(err_msg.format(feature_type))
        hop_size = feature_args.get('hop_size', 0.1)
        return get_l3_frames_uniform(path, l3embedding_model, hop_size=hop_size)
    raise ValueError('Invalid feature type: {}'.format(feature_type))
