"""GPU tests of the reference-facing Python surface: keras-style model object (compile / fit_generator / save_weights /
load_embedding / predict), device-side framing for get_l3_frames_uniform, and train() with its output files and resume
rule (l3embedding/train.py:218-421, l3embedding/model.py:85-181, data/usc/features.py:256-306).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import l3_oracle as O
from _gpu_common import MODEL_TYPES, GOLDEN, F64, engine as _engine, rel_l2, pad as _pad, oracle_inputs as _oracle_inputs

pytestmark = pytest.mark.gpu

def test_keras_style_api_end_to_end(tmp_path):
    from l3embedding_b200 import model as M
    m, _, _ = M.MODELS["cnn_L3_melspec2"]()
    m.configure(dtype="f32")
    m.compile(M.Adam(lr=1e-4), loss="categorical_crossentropy", metrics=["accuracy"])
    video, audio, label = O.synthetic_batch(2, seed=17)

    def gen():
        while True:
            yield [O.scale_video(video), O.pcm2float(audio, "float32")], label
    h = m.fit_generator(gen(), steps_per_epoch=2, epochs=2, validation_data=gen(), validation_steps=1, verbose=0)
    assert set(h.history) == {"loss", "acc", "val_loss", "val_acc"} and len(h.history["loss"]) == 2
    p = str(tmp_path / "model_latest.h5")
    m.save_weights(p)
    e = M.load_embedding(p, "cnn_L3_melspec2", "audio", "original")
    e.parent.configure(dtype="f32")
    x = O.pcm2float(audio, "float32")
    emb = e.predict(x)
    assert emb.shape == (2, 6144)
    w = O.to_torch(m.named_weights(), dtype=torch.float64)
    ref = O.audio_embedding(torch.from_numpy(x).double(), w, "cnn_L3_melspec2", "original", F64).numpy()
    assert np.abs(emb - ref).max() <= 1e-3


def test_device_side_framing_equals_host_framing():
    """get_l3_frames_uniform: embeddings of overlapping 1 s windows read in place on the device == the reference's
    framed-copy route, for int16 and float32 signals, across a max_batch boundary."""
    from l3embedding_b200 import model as M
    from l3embedding_b200.features import get_l3_frames_uniform, frame_signal
    m, _, _ = M.MODELS["cnn_L3_melspec2"]()
    m.configure(dtype="f32")
    m.set_named_weights(O.init_weights("cnn_L3_melspec2", seed=5, randomize_bn=True))
    e, _, _ = M.convert_audio_model_to_embedding(m.get_layer("audio_model"), m.inputs[1], "cnn_L3_melspec2", "short")
    rng = np.random.default_rng(0)
    sig = (0.1 * rng.standard_normal(48000 + 6 * 4800 + 100)).astype(np.float32)
    got = get_l3_frames_uniform(sig, e, hop_size=0.1)
    a, hop, n = frame_signal(sig)
    assert got.shape == (7, 512) and n == 7
    idx = np.arange(48000)[None, :] + hop * np.arange(n)[:, None]
    ref = e.predict(a[idx].reshape(n, 1, 48000), batch_size=4)
    assert np.abs(got - ref).max() <= 1e-5
    w = O.to_torch(m.named_weights(), dtype=torch.float64)
    orc = O.audio_embedding(torch.from_numpy(a[idx].reshape(n, 1, 48000)).double(), w, "cnn_L3_melspec2", "short", F64).numpy()
    assert np.abs(got - orc).max() <= 1e-3
    eng = e._get_engine(4)       # smaller than n: exercises the chunk loop
    i16 = np.clip(np.round(sig * 32767), -32768, 32767).astype(np.int16)
    g16 = eng.embed_audio_frames(i16, hop, "short").cpu().numpy()
    r16 = eng.embed_audio(i16[idx].reshape(n, 1, 48000)[:4], "short").cpu().numpy()
    assert np.abs(g16[:4] - r16).max() <= 1e-5


def test_train_function_writes_reference_files_and_resumes(tmp_path):
    """train() (train.py:218-421): same output files, CSV history, and resume from model_latest.h5 + the CSV."""
    from l3embedding_b200 import train as T
    from l3embedding_b200.synthetic import synthetic_batch
    for name, seed in (("demo_train", 0), ("demo_valid", 50)):
        d = tmp_path / name
        d.mkdir()
        for i in range(2):
            v, a, l = synthetic_batch(4, seed=seed + i)
            np.savez(d / ("b%d.npz" % i), video=v, audio=a, label=l)
    kw = dict(train_epoch_size=2, validation_epoch_size=1, train_batch_size=2, validation_batch_size=2,
              model_type="cnn_L3_melspec2", learning_rate=1e-4, checkpoint_interval=1, disable_logging=True, gpus=1,
              dtype="bf16")
    model_dir, hist = T.train(str(tmp_path / "demo_train"), str(tmp_path / "demo_valid"), str(tmp_path / "out"),
                              num_epochs=2, **kw)
    files = set(os.listdir(model_dir))
    assert {"config.json", "model.json", "model_spec.pkl", "model_latest.h5", "model_best_valid_accuracy.h5",
            "model_best_valid_loss.h5", "model_checkpoint.01.h5", "model_checkpoint.02.h5", "history_checkpoint.pkl",
            "history_csvlog.csv", "history.pkl"} <= files
    assert "embedding/demo/cnn_L3_melspec2" in model_dir.replace(os.sep, "/")
    assert len(hist.history["loss"]) == 2 and set(hist.history) == {"loss", "acc", "val_loss", "val_acc"}
    assert T.get_restart_info(os.path.join(model_dir, "history_csvlog.csv"))[0] == 1
    _, hist2 = T.train(str(tmp_path / "demo_train"), str(tmp_path / "demo_valid"), str(tmp_path / "out"), num_epochs=3,
                       continue_model_dir=model_dir, **kw)
    assert len(hist2.history["loss"]) == 1                      # only epoch index 2 ran
    assert T.get_restart_info(os.path.join(model_dir, "history_csvlog.csv"))[0] == 2
    # the Adam state travels with the checkpoints (the reference restarts the moments from zero on resume): 2 epochs x 2
    # steps before the resume, 2 more after it -- the step count continues instead of restarting at 0
    with np.load(T.optimizer_state_path(os.path.join(model_dir, "model_checkpoint.02.h5"))) as z:
        assert int(z["t"]) == 4 and z["m"].shape == z["v"].shape and np.abs(z["m"]).max() > 0
    with np.load(T.optimizer_state_path(os.path.join(model_dir, "model_latest.h5"))) as z:
        assert int(z["t"]) == 6


def test_reference_embedding_extraction_call_sequence(tmp_path):
    """05_generate_embedding_samples.py:142-157 + data/usc/us8k.py:146-162 + data/usc/features.py:256-306, statement for
    statement, against the shim package: parse the model type out of the model path, `from l3embedding.model import
    load_embedding`, `load_embedding(model_path, model_type, 'audio', pooling_type, tgt_num_gpus=num_gpus)`,
    `get_l3_frames_uniform(audio, model, hop_size)`, `np.savez_compressed(path, X=..., y=...)` -- and the saved X against
    the fp64 oracle on the same frames.  The checkpoint is a multi-GPU-layout file (src: a 4-GPU training run) written
    by train()'s own ModelCheckpoint path."""
    from l3embedding.model import load_embedding, load_model          # the reference's import line
    from l3embedding_b200 import model as M
    from l3embedding_b200.features import get_l3_frames_uniform, frame_signal
    model_type, pooling_type, num_gpus = "cnn_L3_melspec2", "short", 0
    model_dir = tmp_path / "embedding" / "music" / model_type / "20180920112233"
    model_dir.mkdir(parents=True)
    m, _, _ = M.MODELS[model_type]()
    w_np = O.init_weights(model_type, seed=21, randomize_bn=True)
    m.set_named_weights(w_np)
    model_path = str(model_dir / "model_best_valid_accuracy.h5")
    m.save_weights(model_path)
    # 05_generate_embedding_samples.py:144-153
    model_desc_start_idx = model_path.rindex("embedding") + 10
    model_desc_end_idx = os.path.dirname(model_path).rindex("/")
    embedding_desc_str = model_path[model_desc_start_idx:model_desc_end_idx]
    assert embedding_desc_str.split("/")[-1] == model_type
    l3embedding_model = load_embedding(model_path, embedding_desc_str.split("/")[-1], "audio", pooling_type,
                                       tgt_num_gpus=num_gpus)
    l3embedding_model.parent.configure(dtype="f32")
    rng = np.random.default_rng(3)
    audio = (0.1 * rng.standard_normal(48000 + 3 * 4800 + 17)).astype(np.float32)
    X = get_l3_frames_uniform(audio, l3embedding_model, hop_size=0.1)
    out = str(tmp_path / "clip.npz")
    np.savez_compressed(out, X=X, y=3)                                   # us8k.py:162
    with np.load(out) as z:
        assert z["X"].shape == (4, 512) and z["X"].dtype == np.float32 and int(z["y"]) == 3
    a, hop, n = frame_signal(audio)
    idx = np.arange(48000)[None, :] + hop * np.arange(n)[:, None]
    ref = O.audio_embedding(torch.from_numpy(a[idx].reshape(n, 1, 48000)).double(), O.to_torch(w_np, dtype=torch.float64),
                            model_type, pooling_type, F64).numpy()
    assert np.abs(X - ref).max() <= 1e-3
    # a checkpoint in the multi-GPU layout (what a 4-GPU reference run saves) loads through the same calls
    multi = str(model_dir / "model_latest.h5")
    M.multi_gpu_model(m, gpus=4).save_weights(multi)
    m.num_gpus = 0
    e4 = load_embedding(multi, model_type, "audio", "original", src_num_gpus=4, tgt_num_gpus=1)
    e4.parent.configure(dtype="f32")
    X4 = e4.predict(a[idx].reshape(n, 1, 48000))
    ref4 = O.audio_embedding(torch.from_numpy(a[idx].reshape(n, 1, 48000)).double(), O.to_torch(w_np, dtype=torch.float64),
                             model_type, "original", F64).numpy()
    assert X4.shape == (n, 6144) and np.abs(X4 - ref4).max() <= 1e-3


def test_embedding_predict_pipeline_equals_direct_calls():
    """EmbeddingModel.predict streams host arrays through a three-stage pinned pipeline in device batches of 512
    (keras' batch_size of 32 is a request, not a constraint: inference results do not depend on the batching).  Across
    chunk boundaries and a ragged last chunk it returns exactly what direct engine calls on the same clips return, for
    int16 and for float32 input."""
    from l3embedding_b200 import model as M
    m, _, _ = M.MODELS["cnn_L3_melspec2"]()
    m.configure(dtype="bf16")
    m.set_named_weights(O.init_weights("cnn_L3_melspec2", seed=5, randomize_bn=True))
    e, _, _ = M.convert_audio_model_to_embedding(m.get_layer("audio_model"), m.inputs[1], "cnn_L3_melspec2", "short")
    _, a64, _ = O.synthetic_batch(64, seed=9)
    n = 1100                                                     # 512 + 512 + 76
    x = np.concatenate([a64] * 18)[:n]
    x[::7] //= 2                                                 # not all chunks alike
    got = e.predict(x)
    assert got.shape == (n, 512) and got.dtype == np.float32
    eng = e._get_engine(512)
    for s0 in (0, 512, 1024):
        ref = eng.embed_audio(x[s0:s0 + 512], "short").cpu().numpy()
        assert np.array_equal(got[s0:s0 + len(ref)], ref), s0
    got_f = e.predict(O.pcm2float(x[:600], "float32"), batch_size=32)
    assert np.array_equal(got_f, got[:600])                      # pcm2float is exact in fp32: same embeddings
