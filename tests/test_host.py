"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/l3b200.h declares, the
weight inventory matches the oracle's, the keras-compatible model objects behave like the reference's at the
boundary (names, argument meaning, error behaviour), and the N>1 data-parallel glue works over gloo."""
import os
import re
import socket
import sys

import numpy as np
import pytest

from l3embedding_b200 import _lib, dp, weights_io
from l3embedding_b200 import model as M
from oracle import l3_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL_TYPES = ["cnn_L3_orig", "cnn_L3_kapredbinputbn", "cnn_L3_melspec1", "cnn_L3_melspec2"]


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "l3b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(l3_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libl3b200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.l3_version() == 1


@pytest.mark.parametrize("model_type", MODEL_TYPES)
def test_tensor_table_matches_oracle_layout(model_type):
    table = _lib.tensor_table(model_type)
    lay = O.model_layout(model_type)
    assert [(n, s) for n, _a, _o, s in table] == [(n, tuple(s)) for n, s, _t in lay]
    assert [a for _n, a, _o, _s in table] == [0 if t else 1 for _n, _s, t in lay]
    lib = _lib.load()
    mid = _lib.model_id(model_type)
    c = O.count_params(model_type)
    assert lib.l3_param_count(mid) == c["trainable"] and lib.l3_state_count(mid) == c["bn_moving"]
    # arenas are dense and the l2-regularised kernels come first
    for arena in (0, 1):
        spans = sorted((o, o + int(np.prod(s))) for _n, a, o, s in table if a == arena)
        assert spans[0][0] == 0 and all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
    kernels = [(o, int(np.prod(s))) for n, a, o, s in table if n.endswith("/kernel")]
    assert max(o + k for o, k in kernels) == lib.l3_l2_count(mid) == sum(k for _o, k in kernels)


def test_geometry_queries():
    lib = _lib.load()
    import ctypes as C
    a, b = C.c_int(), C.c_int()
    for mt, fe, emb in (("cnn_L3_melspec2", (256, 199), (32, 24)), ("cnn_L3_melspec1", (128, 199), (16, 24)),
                        ("cnn_L3_orig", (257, 197), (32, 24)), ("cnn_L3_kapredbinputbn", (257, 197), (32, 24))):
        assert lib.l3_frontend_shape(_lib.model_id(mt), C.byref(a), C.byref(b)) == 0 and (a.value, b.value) == fe
        assert lib.l3_embedding_map_shape(_lib.model_id(mt), C.byref(a), C.byref(b)) == 0 and (a.value, b.value) == emb
    assert lib.l3_frontend_shape(9, C.byref(a), C.byref(b)) < 0 and b"invalid model type" in lib.l3_last_error()
    assert lib.l3_workspace_bytes(3, 0, 0, 7) < 0
    assert lib.l3_workspace_bytes(3, 64, 1, 7) > lib.l3_workspace_bytes(3, 64, 1, 6) > 0   # training needs more


def test_models_registry_and_errors():
    # model.py:307-313 keys; model.py:113-114 / :175-176 error behaviour
    assert set(M.MODELS) == {"cnn_L3_orig", "tiny_L3", "cnn_L3_kapredbinputbn", "cnn_L3_melspec1", "cnn_L3_melspec2"}
    with pytest.raises(ValueError, match='Invalid model type: "nope"'):
        M.load_model("/nonexistent", "nope")
    m, inputs, y = M.MODELS["cnn_L3_melspec2"]()
    assert m.name == "cnn_L3_melspec2" and len(inputs) == 2 and inputs[0].shape == (None, 224, 224, 3)
    assert inputs[1].shape == (None, 1, 48000) and y.shape == (None, 2)
    assert [l.name for l in m.layers[2:]] == ["vision_model", "audio_model", "concatenate_1", "dense_1", "dense_2"]
    assert m.get_layer("audio_model").get_layer("audio_embedding_layer").name == "audio_embedding_layer"
    assert m.get_layer("vision_model").get_layer("vision_embedding_layer") is not None
    with pytest.raises(ValueError):
        m.get_layer("missing")
    m4, _, _ = M.MODELS["cnn_L3_melspec2"](num_gpus=4)
    assert m4.num_gpus == 4
    with pytest.raises(RuntimeError, match="compile"):
        m.train_on_batch([np.zeros((1, 224, 224, 3), np.float32), np.zeros((1, 1, 48000), np.float32)], np.zeros((1, 2)))


@pytest.mark.parametrize("model_type", ["cnn_L3_melspec2", "cnn_L3_orig"])
def test_keras_weight_order_and_counts(model_type):
    m, _, _ = M.MODELS[model_type]()
    names = m.weight_names()
    c = O.count_params(model_type)
    assert m.count_params() == c["total"]
    # nested towers: all trainable arrays before any non-trainable one (keras Container.weights)
    vis = m.get_layer("vision_model")._weight_names
    first_nt = min(i for i, n in enumerate(vis) if n.endswith("moving_mean"))
    assert all(n.endswith(("moving_mean", "moving_variance")) for n in vis[first_nt:])
    aud = m.get_layer("audio_model")._weight_names
    k = aud.index("kapre/real_kernels")
    assert aud[k + 1] == "kapre/imag_kernels" and not any(n.endswith(("kernel", "bias", "gamma", "beta")) for n in aud[k:])
    # layer-by-layer order on the tower itself starts with the kapre arrays
    # (notebooks/extract_spectrogram_models_from_avc_models.ipynb:446: get_weights()[3:] strips them for mel models)
    lw = m.get_layer("audio_model").layerwise_weight_names()
    assert lw[0] == "kapre/real_kernels" and lw[1] == "kapre/imag_kernels"
    assert names[-4:] == ["dense_1/kernel", "dense_1/bias", "dense_2/kernel", "dense_2/bias"]
    kc = weights_io.kapre_constants(model_type)
    n_dft = 2048 if "melspec" in model_type else 512
    assert kc["kapre/real_kernels"].shape == (n_dft, 1, 1, n_dft // 2 + 1)


def test_save_load_round_trip_and_convert(tmp_path):
    m, _, _ = M.MODELS["cnn_L3_melspec1"]()
    w = m.get_weights()
    w2 = [a + 1.0 if a.ndim == 1 and a.shape[0] == 64 else a for a in w]
    m.set_weights(w2)
    p = str(tmp_path / "model_latest.h5")     # the reference's file name (train.py:316); npz container without h5py
    m.save_weights(p)
    m2 = M.load_model(p, "cnn_L3_melspec1")
    assert all(np.array_equal(a, b) for a, b in zip(m.get_weights(), m2.get_weights()))
    m3 = M.load_model(p, "cnn_L3_melspec1", src_num_gpus=4, tgt_num_gpus=1)
    assert m3.num_gpus == 0 and all(np.array_equal(a, b) for a, b in zip(m.get_weights(), m3.get_weights()))
    with pytest.raises(ValueError, match="different model layout"):
        M.load_model(p, "cnn_L3_orig")
    with pytest.raises(ValueError, match="expecting"):
        m.set_weights(w[:-1])
    bad = list(w)
    bad[-1] = np.zeros(3, np.float32)
    with pytest.raises(ValueError, match="not compatible"):
        m.set_weights(bad)


def test_load_embedding_boundary(tmp_path):
    m, _, _ = M.MODELS["cnn_L3_melspec2"]()
    p = str(tmp_path / "w.npz")
    m.save_weights(p)
    e = M.load_embedding(p, "cnn_L3_melspec2", "audio", "original")
    assert e.output_dim == 6144
    e2, x, y = M.load_embedding(p, "cnn_L3_melspec2", "audio", "short", return_io=True)
    assert e2.output_dim == 512 and x.shape == (None, 1, 48000) and y.shape == (None, 512)
    assert M.load_embedding(p, "cnn_L3_melspec2", "vision", "original").output_dim == 8192
    with pytest.raises(ValueError, match='Invalid embedding type: "smell"'):
        M.load_embedding(p, "cnn_L3_melspec2", "smell", "original")
    with pytest.raises(KeyError):
        M.load_embedding(p, "cnn_L3_melspec2", "audio", "tiny")


def test_framing_rule_matches_reference_quirk():
    """data/usc/features.py:279-298: short clips are padded symmetrically to one frame; longer clips are NEVER padded
    (operator-precedence quirk) and the trailing partial hop is dropped."""
    from l3embedding_b200.features import frame_signal

    class Rec:
        def predict(self, x):
            self.x = x
            return np.zeros((len(x), 4))
    a, hop, n = frame_signal(np.arange(30000, dtype=np.float32))
    assert len(a) == 48000 and hop == 4800 and n == 1 and a[9000] == 0.0 and a[9001] == 1.0
    a, hop, n = frame_signal(np.zeros(48000 * 4 + 1234, np.float32))
    assert len(a) == 48000 * 4 + 1234 and n == 1 + (3 * 48000 + 1234) // 4800
    from l3embedding_b200.features import get_l3_frames_uniform, compute_file_features
    r = Rec()
    sig = np.arange(48000 + 2 * 4800 + 7, dtype=np.float32)
    out = get_l3_frames_uniform(sig, r)
    assert out.shape == (3, 4) and r.x.shape == (3, 1, 48000) and r.x[2, 0, 0] == 9600.0
    with pytest.raises(ValueError, match="Must provide L3 embedding model"):
        compute_file_features("x.wav", "l3")
    with pytest.raises(ValueError, match="Invalid feature type"):
        compute_file_features("x.wav", "mfcc", l3embedding_model=r)


def _write_batches(d, n_files=2, per_file=5, seed=0):
    from l3embedding_b200.synthetic import synthetic_batch
    os.makedirs(d, exist_ok=True)
    for i in range(n_files):
        v, a, l = synthetic_batch(per_file, seed=seed + i)
        np.savez(os.path.join(d, "batch_%02d.npz" % i), video=v, audio=a, label=l)


def test_data_generator_and_restart_info(tmp_path):
    """train.py:142-215: batches are cut across file boundaries; start_batch_idx skips without yielding; raw u8/i16 by
    default, the reference's float scaling on request; get_restart_info reads the last CSV row."""
    from l3embedding_b200 import train as T
    d = str(tmp_path / "subset_train")
    _write_batches(d)
    g = T.data_generator(d, batch_size=4, random_state=1)
    cp = lambda b: {k: v.copy() for k, v in b.items()}     # ring slots are recycled after two further batches
    b0, b1, b2 = cp(next(g)), cp(next(g)), cp(next(g))
    assert b0["video"].shape == (4, 224, 224, 3) and b0["video"].dtype == np.uint8 and b0["audio"].dtype == np.int16
    assert b1["label"].shape == (4, 2) and b1["label"].dtype == np.float32
    g2 = T.data_generator(d, batch_size=4, random_state=1, start_batch_idx=1)
    assert np.array_equal(next(g2)["audio"], b1["audio"])          # resume lands on the same batch
    gs = T.data_generator(d, batch_size=4, random_state=1, scale_on_host=True)
    s0 = next(gs)
    assert s0["video"].dtype == np.float32 and s0["video"].min() >= -1 and s0["video"].max() <= 1
    assert np.array_equal(s0["audio"], b0["audio"].astype(np.float32) / 32768)
    x, y = next(T.keras_tuples(T.data_generator(d, batch_size=2), ["video", "audio"], "label"))
    assert len(x) == 2 and x[0].shape == (2, 224, 224, 3) and y.shape == (2, 2)
    ev = T.single_epoch_data_generator(d, 2, batch_size=4, random_state=1)
    e = [cp(next(ev)) for _ in range(4)]
    assert np.array_equal(e[0]["audio"], e[2]["audio"]) and np.array_equal(e[1]["audio"], e[3]["audio"])
    p = tmp_path / "history_csvlog.csv"
    p.write_text("epoch,acc,loss,val_acc,val_loss\n0,0.5,0.9,0.55,0.8\n1,0.6,0.7,0.65,0.6\n")
    assert T.get_restart_info(str(p)) == (1, 0.65, 0.6)


def test_training_callbacks(tmp_path):
    from l3embedding_b200 import train as T

    class Dummy:
        saved = []

        def save_weights(self, path):
            self.saved.append(os.path.basename(path))
    m = Dummy()
    cbs = [T.ModelCheckpoint(str(tmp_path / "model_latest.h5")),
           T.ModelCheckpoint(str(tmp_path / "best_acc.h5"), save_best_only=True, monitor="val_acc"),
           T.ModelCheckpoint(str(tmp_path / "best_loss.h5"), save_best_only=True, monitor="val_loss"),
           T.ModelCheckpoint(str(tmp_path / "model_checkpoint.{epoch:02d}.h5"), period=2),
           T.CSVLogger(str(tmp_path / "history_csvlog.csv"), append=True), T.LossHistory(str(tmp_path / "h.pkl")),
           T.TimeHistory()]
    for c in cbs:
        c.set_model(m)
        if hasattr(c, "on_train_begin"):
            c.on_train_begin()
    logs = [dict(loss=1.0, acc=0.5, val_loss=0.9, val_acc=0.5), dict(loss=0.8, acc=0.6, val_loss=1.1, val_acc=0.7),
            dict(loss=0.7, acc=0.7, val_loss=0.5, val_acc=0.6)]
    for e, lg in enumerate(logs):
        for c in cbs:
            if hasattr(c, "on_epoch_begin"):
                c.on_epoch_begin(e)
        for c in cbs:
            c.on_epoch_end(e, lg)
    assert m.saved.count("model_latest.h5") == 3
    assert m.saved.count("best_acc.h5") == 2 and m.saved.count("best_loss.h5") == 2
    assert "model_checkpoint.02.h5" in m.saved and "model_checkpoint.01.h5" not in m.saved
    assert T.get_restart_info(str(tmp_path / "history_csvlog.csv")) == (2, 0.6, 0.5)
    import pickle
    assert pickle.load(open(tmp_path / "h.pkl", "rb")) == {"loss": [1.0, 0.8, 0.7], "val_loss": [0.9, 1.1, 0.5]}


def test_compute_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m, _, _ = M.MODELS["cnn_L3_orig"]()
    with pytest.raises(_lib.L3Error, match="no CPU fallback"):
        m.predict([np.zeros((1, 224, 224, 3), np.float32), np.zeros((1, 1, 48000), np.float32)])


def test_replica_slice_follows_get_slice():
    # training_utils.py:121-133
    assert [dp.replica_slice(64, r, 4) for r in range(4)] == [slice(0, 16), slice(16, 32), slice(32, 48), slice(48, 64)]
    assert [dp.replica_slice(10, r, 4) for r in range(4)] == [slice(0, 2), slice(2, 4), slice(4, 6), slice(6, 10)]
    assert dp.current(0).slice(7) == slice(0, 7)
    with pytest.raises(RuntimeError, match="one process per GPU"):
        dp.current(2)


def _dp_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from l3embedding_b200 import train as T
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        par = dp.current(world)

        class FakeEngine:
            """stands in for the CUDA engine: records what the bootstrap hands to l3_dp_init"""
            dp_world = None
            joined = None

            @staticmethod
            def dp_unique_id():
                return bytes([65 + rank]) * 128          # only rank 0's id may reach l3_dp_init

            def dp_init(self, uid, r, n):
                self.joined, self.dp_world = (uid, r, n), (r, n)
        eng = FakeEngine()
        par.attach(eng)
        par.attach(eng)                                   # idempotent: joins once
        s = par.sum_scalars(float(rank), 1.0)
        # train()'s rank awareness: rank 0 is the chief and names the directory for everybody
        r, w, bcast = T._replicas(world)
        model_dir = bcast("dir-from-rank-%d" % rank if r == 0 else None)
        q.put((rank, par.slice(10), eng.joined, s, (r, w, model_dir)))
    finally:
        dist.destroy_process_group()


def test_data_parallel_bootstrap_over_gloo():
    """world_size 2 on CPU (gloo): slices, the NCCL-id hand-over of the bootstrap, the validation-scalar sum and the
    rank-0-names-the-directory rule of train().  The gradient exchange itself is CUDA/NCCL (tests/test_gpu_g_dp.py)."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == slice(0, 5) and res[1][1] == slice(5, 10)
    assert res[0][2] == (b"A" * 128, 0, 2) and res[1][2] == (b"A" * 128, 1, 2)
    assert all(r[3] == (1.0, 2.0) for r in res)
    assert res[0][4] == (0, 2, "dir-from-rank-0") and res[1][4] == (1, 2, "dir-from-rank-0")


def test_batch_plan_reproduces_the_reference_sequence(tmp_path):
    """pipeline.BatchPlan against a literal simulation of train.py:134-195 (cycle_shuffle over the listing, one global
    `random` seeded once, samples cut into consecutive runs of batch_size across file and pass boundaries), including
    the resume seek; and the prefetching reader against the thread-less path."""
    import random
    from l3embedding_b200 import pipeline as P
    d = str(tmp_path / "subset_train")
    os.makedirs(d)
    sizes = {"b0.npz": 5, "b1.npz": 3, "b2.npz": 7, "b3.npz": 2}
    rng = np.random.default_rng(0)
    for name, n in sizes.items():
        np.savez(os.path.join(d, name), video=rng.integers(0, 256, (n, 224, 224, 3), dtype=np.uint8),
                 audio=rng.integers(-9, 9, (n, 1, 48000)).astype(np.int16),
                 label=np.eye(2, dtype=np.float32)[rng.integers(0, 2, n)])

    def reference_stream(batch_size, seed, n_batches):
        random.seed(seed)                                   # train.py:144
        lst = sorted(sizes)
        out, cur = [], []
        need = batch_size
        while len(out) < n_batches:
            for f in list(lst):                             # cycle_shuffle: one pass, then shuffle in place
                pos = 0
                while pos < sizes[f]:
                    take = min(need, sizes[f] - pos)
                    cur.append((f, pos, pos + take))
                    pos += take
                    need -= take
                    if need == 0:
                        out.append(cur)
                        cur, need = [], batch_size
            random.shuffle(lst)
        return out[:n_batches]
    for bs, seed in ((4, 1), (6, 20180123), (17, 3)):
        plan = P.BatchPlan(d, bs, random_state=seed)
        want = reference_stream(bs, seed, 12)
        got = [[(os.path.basename(p), s, e) for p, s, e in segs] for segs, _ in zip(plan.segments(0), range(12))]
        assert got == want
        seek = [[(os.path.basename(p), s, e) for p, s, e in segs] for segs, _ in zip(plan.segments(7), range(3))]
        assert seek == want[7:10]
    plan = P.BatchPlan(d, 4, random_state=1)
    reader = P.PinnedBatchReader(plan, start_batch=2, max_batches=7, pinned=False)
    ref = P.read_batches(plan, start_batch=2)
    n, held = 0, []
    for b in reader:
        r = next(ref)
        assert all(np.array_equal(b[k], r[k]) and b[k].dtype == r[k].dtype for k in P.KEYS)
        held = (held + [(b, r)])[-2:]          # the batch just taken and the one before it are still intact
        assert all(np.array_equal(hb["video"], hr["video"]) for hb, hr in held)
        n += 1
    assert n == 7
    reader.close()


def test_engine_is_created_once_across_train_validate_train(monkeypatch):
    """validation with a bigger batch (or any predict) between steps must not rebuild the engine and thereby
    reset Adam's moments / step count; a genuinely larger training batch rebuilds ONCE and adopts the state."""
    from l3embedding_b200 import model as M, engine as E

    created = []

    class StubEngine:
        def __init__(self, model_type, max_batch, dtype="f32", training=True, weights=None, **_):
            self.model_type, self.max_batch, self.dtype, self.training = model_type, max_batch, dtype, training
            self.adam_t, self.adopted, self.closed, self.staged = 0, None, False, []
            created.append(self)

        def adopt_state(self, other):
            self.adopted, self.adam_t = other, other.adam_t

        def close(self):
            self.closed = True

        def upload_host(self, v, a, l):
            self.staged.append(len(v))
            return len(v)

        def train_step_staged(self, n, lr):
            assert self.staged.pop(0) == n
            self.adam_t += 1
            return dict(loss=1.0, acc=0.5)

        def predict(self, v, a, y=None):
            assert len(v) <= self.max_batch
            return np.zeros((len(v), 2), np.float32), np.zeros((len(v), 2), np.float32)

        def metrics(self):
            return dict(ce_sum=1.0, correct=1.0, l2=0.0)

        def get_weights(self):
            return {}
    monkeypatch.setattr(E, "Engine", StubEngine)
    m, _, _ = M.MODELS["cnn_L3_orig"]()
    m.compile(M.Adam(lr=1e-4))
    x = lambda n: [np.zeros((n, 224, 224, 3), np.uint8), np.zeros((n, 1, 48000), np.int16)]
    y = lambda n: np.tile(np.array([[1, 0]], np.float32), (n, 1))
    for _ in range(2):
        m.train_on_batch(x(32), y(32))
        m.test_on_batch(x(64), y(64))           # validation batch twice the training batch: chunked, not rebuilt
        m.predict(x(8))
    assert len(created) == 1 and created[0].training and created[0].adam_t == 2
    m.train_on_batch(x(48), y(48))              # a larger TRAINING batch: one rebuild that adopts the optimizer state
    assert len(created) == 2 and created[1].adopted is created[0] and created[0].closed
    assert created[1].training and created[1].max_batch == 48 and created[1].adam_t == 3


def test_multi_gpu_checkpoint_layout_round_trip(tmp_path):
    """model.py:76-77,117-119: a checkpoint saved from a multi_gpu_model keeps the whole template model as ONE nested
    layer with its arrays in Container order (all trainable, then all non-trainable) -- a different array order than
    the single-model file.  Both layouts load into the right tensors whatever src_num_gpus says; a file of another
    model type is refused."""
    from l3embedding_b200 import model as M, minihdf5
    m, _, _ = M.MODELS["cnn_L3_kapredbinputbn"]()
    rng = np.random.default_rng(1)
    w = {k: (v + 0.01 * rng.standard_normal(v.shape)).astype(np.float32) for k, v in m.named_weights().items()}
    m.set_named_weights(w)
    single, multi = str(tmp_path / "single.h5"), str(tmp_path / "multi.h5")
    m.save_weights(single)
    M.multi_gpu_model(m, gpus=4).save_weights(multi)
    f = minihdf5.File(multi)
    names = [n.decode() for n in f.attrs["layer_names"]]
    assert names == ["input_1", "input_2"] + ["lambda_%d" % i for i in range(1, 9)] + ["cnn_L3_kapredbinputbn", "dense_2"]
    wn = [n.decode() for n in f["cnn_L3_kapredbinputbn"].attrs["weight_names"]]
    assert len(wn) == len(m.weight_names())
    first_nt = next(i for i, n in enumerate(wn) if "moving" in n or "kapre" in n)
    assert all(("moving" in n or "kapre" in n) for n in wn[first_nt:])          # trainable block, then non-trainable
    assert m.container_weight_names() != m.weight_names()
    for path in (single, multi):
        for src in (0, 4):
            got = M.load_model(path, "cnn_L3_kapredbinputbn", src_num_gpus=src, tgt_num_gpus=1).named_weights()
            assert all(np.array_equal(got[k], w[k]) for k in w), (path, src)
    with pytest.raises(ValueError):
        M.load_model(multi, "cnn_L3_orig", src_num_gpus=4)


REFERENCE = "/root/reference"
needs_reference = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree exists in the build container only")


def test_shim_package_resolves_the_reference_imports():
    """05_generate_embedding_samples.py:5, 03_train_embedding.py:4, classifier/train.py:28, l3embedding/train.py:17-18,
    l3embedding/model.py:2-4: the reference's import statements, against the shim package at the repo root."""
    ns = {}
    exec("from l3embedding.model import load_embedding\n"
         "from l3embedding.train import *\n"
         "from l3embedding.train import LossHistory\n"
         "from l3embedding.model import MODELS, load_model\n"
         "from l3embedding.audio import pcm2float\n"
         "from l3embedding.training_utils import multi_gpu_model\n"
         "from l3embedding.vision_model import *\n"
         "from l3embedding.audio_model import *\n", ns)
    from l3embedding_b200 import model as M, train as T
    assert ns["load_embedding"] is M.load_embedding and ns["train"] is T.train and ns["LossHistory"] is T.LossHistory
    assert set(ns["MODELS"]) == {"cnn_L3_orig", "tiny_L3", "cnn_L3_kapredbinputbn", "cnn_L3_melspec1", "cnn_L3_melspec2"}
    # tower sub-builders + merge reproduce the registry's pairings (model.py:198-284)
    for vis, aud, mt in (("construct_cnn_L3_orig_vision_model", "construct_cnn_L3_orig_audio_model", "cnn_L3_orig"),
                         ("construct_cnn_L3_orig_inputbn_vision_model", "construct_cnn_L3_kapredbinputbn_audio_model", "cnn_L3_kapredbinputbn"),
                         ("construct_cnn_L3_orig_inputbn_vision_model", "construct_cnn_L3_melspec1_audio_model", "cnn_L3_melspec1"),
                         ("construct_cnn_L3_orig_inputbn_vision_model", "construct_cnn_L3_melspec2_audio_model", "cnn_L3_melspec2")):
        vm, x_i, y_i = ns[vis]()
        am, x_a, y_a = ns[aud]()
        m, inputs, y = M.L3_merge_audio_vision_models(vm, x_i, am, x_a, mt)
        ref, _, _ = M.MODELS[mt]()
        assert m.name == mt and m.weight_names() == ref.weight_names() and inputs == [x_i, x_a]
        assert am.get_layer("audio_embedding_layer") and vm.get_layer("vision_embedding_layer")
    vm, x_i, _ = ns["construct_cnn_L3_orig_vision_model"]()
    am, x_a, _ = ns["construct_cnn_L3_melspec2_audio_model"]()
    with pytest.raises(ValueError, match="unsupported tower pairing"):
        M.L3_merge_audio_vision_models(vm, x_i, am, x_a, "x")


@needs_reference
def test_reference_train_cli_runs_against_the_shim_and_flag_surfaces_match():
    """The reference's OWN 03_train_embedding.py, unmodified, imports `l3embedding.train` from the shim and parses its
    flags; l3embedding_b200.cli yields the same argument dictionary (03_train_embedding.py:7-153)."""
    import runpy
    import subprocess
    # run as a script, but with the repo root (the shim) on the path instead of the script's own directory
    script = os.path.join(REFERENCE, "03_train_embedding.py")
    code = ("import sys, runpy; sys.path.insert(0, %r); sys.argv = ['03_train_embedding.py', '-h']; "
            "runpy.run_path(%r, run_name='__main__')" % (ROOT, script))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "--continue-model-dir" in r.stdout, r.stderr
    from l3embedding_b200 import cli
    argv = ["-e", "3", "-tes", "5", "--gpus", "4", "-mt", "cnn_L3_melspec2", "-cmd", "/x", "-v", "tr", "va", "out"]
    old = sys.argv
    try:
        sys.argv = ["03_train_embedding.py"] + argv
        ref_ns = runpy.run_path(os.path.join(REFERENCE, "03_train_embedding.py"), run_name="not_main")
        ref_args = ref_ns["parse_arguments"]()
    finally:
        sys.argv = old
    ours = cli.parse_arguments(argv)
    assert {k: ours[k] for k in ref_args} == ref_args
    assert set(ours) - set(ref_args) == {"dtype", "scale_on_host"}
    import inspect
    from l3embedding_b200 import train as T
    assert set(ref_args) <= set(inspect.signature(T.train).parameters)


@needs_reference
def test_framing_equals_the_reference_function_executed():
    """get_l3_frames_uniform: the reference's OWN function body (data/usc/features.py:256-306, read from the reference
    tree at test time and executed with a numpy stand-in for librosa.util.utils.frame) hands its model exactly the
    frames ours does -- including the operator-precedence quirk that leaves long clips unpadded."""
    import re
    import types
    from l3embedding_b200 import features as F
    src = open(os.path.join(REFERENCE, "data/usc/features.py")).read()
    body = re.search(r"^def get_l3_frames_uniform\(.*?(?=^def )", src, re.S | re.M).group(0)

    def frame(y, frame_length, hop_length):
        n = 1 + (len(y) - frame_length) // hop_length
        return np.stack([y[i * hop_length:i * hop_length + frame_length] for i in range(n)], axis=1)
    librosa = types.SimpleNamespace(util=types.SimpleNamespace(utils=types.SimpleNamespace(frame=frame)))
    ns = {"np": np, "librosa": librosa}
    exec(body, ns)

    class Recorder:
        def predict(self, x):
            self.x = np.array(x)
            return np.zeros((len(x), 512), np.float32)
    rng = np.random.default_rng(0)
    for n in (1000, 48000, 48001, 52799, 52800, 52801, 48000 * 3 + 123, 48000 + 4800 * 7):
        audio = rng.standard_normal(n).astype(np.float32)
        a, b = Recorder(), Recorder()
        ns["get_l3_frames_uniform"](audio.copy(), a, hop_size=0.1)
        F.get_l3_frames_uniform(audio.copy(), b, hop_size=0.1)
        assert a.x.shape == b.x.shape and np.array_equal(a.x, b.x), n
        padded, hop, n_frames = F.frame_signal(audio)
        assert n_frames == a.x.shape[0] and hop == 4800


def test_frontend_register_fft_on_the_host(tmp_path):
    """The front-end's FFT (csrc/frontend_fft.cuh: 16 x 16 x R3 in registers, three shared-memory exchanges) is
    __host__ __device__: the harness runs the kernel's own per-thread phases on the CPU against a float64 DFT."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "fe_fft_host.cu")
    exe = str(tmp_path / "fe_fft_host")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", exe, src], check=True,
                   capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    errs = dict(line.split() for line in r.stdout.strip().splitlines())
    assert r.returncode == 0, r.stdout
    assert all(float(v) < 5e-7 for v in errs.values()), errs
