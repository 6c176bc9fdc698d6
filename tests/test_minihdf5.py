"""CPU tests of the built-in HDF5 reader / writer (l3embedding_b200/minihdf5.py) that stands in for h5py: the reader
against a file written by libhdf5 itself, the writer by round trip, and the keras-layout weight files and gzip batch
files the reference path uses (l3embedding/model.py:119, train.py:142-195, data/avc/sample.py:373-377)."""
import os

import numpy as np
import pytest

from l3embedding_b200 import minihdf5 as H

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reads_a_file_written_by_libhdf5():
    """tests/golden/libhdf5_written_matlab73.mat is scipy's test fixture testhdf5_7.4_GLNX86.mat (BSD licence): a MATLAB
    v7.3 file, i.e. HDF5 1.8 output with a 512-byte user block, a v0 superblock, an old-style root group and a
    version-2 layout message -- the same generation of the format keras 2.0.9 / h5py 2.7 wrote."""
    f = H.File(os.path.join(HERE, "golden", "libhdf5_written_matlab73.mat"))
    assert f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.attrs["MATLAB_class"] == b"double"
    np.testing.assert_allclose(np.asarray(d).ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


def test_round_trip_groups_attributes_dtypes(tmp_path):
    rng = np.random.default_rng(0)
    tree = {"g1": {"__attrs__": {"weight_names": np.array([b"a/kernel:0", b"a/bias:0"]), "n": np.int32(7)},
                   "a": {"kernel:0": rng.standard_normal((3, 3, 3, 64)).astype(np.float32),
                         "bias:0": np.zeros(64, np.float32)}},
            "empty_group": {"__attrs__": {"weight_names": np.zeros((0,), "S1")}},
            "i16": np.arange(-5, 5, dtype=np.int16).reshape(2, 5), "u8": np.arange(6, dtype=np.uint8),
            "f64": np.linspace(0, 1, 7), "scalar": np.float32(2.5), "none": np.zeros((0, 4), np.float32)}
    p = str(tmp_path / "t.h5")
    H.write_tree(p, tree, attrs={"layer_names": np.array([b"g1", b"empty_group"]), "backend": np.bytes_(b"tensorflow")})
    with open(p, "rb") as fh:
        assert fh.read(8) == H.SIGNATURE
    f = H.File(p)
    assert sorted(f.keys()) == sorted(k for k in tree)
    assert list(f.attrs["layer_names"]) == [b"g1", b"empty_group"] and f.attrs["backend"] == b"tensorflow"
    assert list(f["g1"].attrs["weight_names"]) == [b"a/kernel:0", b"a/bias:0"] and int(f["g1"].attrs["n"]) == 7
    assert len(f["empty_group"].attrs["weight_names"]) == 0 and f["empty_group"].keys() == []
    for path, ref in (("g1/a/kernel:0", tree["g1"]["a"]["kernel:0"]), ("i16", tree["i16"]), ("u8", tree["u8"]),
                      ("f64", tree["f64"]), ("none", tree["none"])):
        got = np.asarray(f[path])
        assert got.dtype == ref.dtype and got.shape == ref.shape and np.array_equal(got, ref), path
    assert float(np.asarray(f["scalar"]).reshape(-1)[0]) == 2.5
    with pytest.raises(KeyError):
        f["g1/missing"]
    assert "g1/a" in f and "nope" not in f


def test_many_members_in_one_group(tmp_path):
    tree = {"layer_%03d" % i: np.full((2,), i, np.float32) for i in range(100)}
    p = str(tmp_path / "many.h5")
    H.write_tree(p, tree)
    f = H.File(p)
    assert len(f.keys()) == 100 and all(float(np.asarray(f["layer_%03d" % i])[0]) == i for i in range(100))
    with pytest.raises(H.HDF5Error):
        H.write_tree(p, {"x%d" % i: np.zeros(1, np.float32) for i in range(200)})


@pytest.mark.parametrize("shuffle", [False, True])
def test_chunked_gzip_batch_file(tmp_path, shuffle):
    """The AVC batch layout of data/avc/sample.py:373-377 (gzip datasets `audio`, `video`, `label`) at a small size,
    with ragged edge chunks."""
    rng = np.random.default_rng(1)
    audio = (rng.standard_normal((5, 1, 4800)) * 3000).astype(np.int16)
    video = rng.integers(0, 255, (5, 20, 20, 3)).astype(np.uint8)
    label = np.eye(2, dtype=np.float32)[rng.integers(0, 2, 5)]
    p = str(tmp_path / "batch.h5")
    H.write_tree(p, {"audio": H.Chunked(audio, (2, 1, 1700), shuffle=shuffle), "video": H.Chunked(video, (3, 7, 20, 3)),
                     "label": label})
    f = H.File(p)
    assert np.array_equal(np.asarray(f["audio"]), audio) and np.array_equal(np.asarray(f["video"]), video)
    assert np.array_equal(f["label"][1:3], label[1:3])


def test_two_level_chunk_btree(tmp_path):
    """More than 64 chunks: the chunk index becomes a two-level B-tree (what h5py's auto-chunking produces for the
    reference's 1024-sample batch files); the reader walks internal nodes."""
    rng = np.random.default_rng(3)
    x = rng.integers(0, 255, (40, 30, 7), dtype=np.uint8)
    p = str(tmp_path / "deep.h5")
    H.write_tree(p, {"x": H.Chunked(x, (3, 4, 7))})   # 14 * 8 = 112 chunks
    assert np.array_equal(np.asarray(H.File(p)["x"]), x)


def test_data_generator_reads_hdf5_batches(tmp_path):
    from l3embedding_b200 import train as T
    rng = np.random.default_rng(2)
    d = tmp_path / "train"
    d.mkdir()
    blobs = []
    for i in range(2):
        audio = (rng.standard_normal((3, 1, 48000)) * 3000).astype(np.int16)
        video = rng.integers(0, 255, (3, 224, 224, 3)).astype(np.uint8)
        label = np.eye(2, dtype=np.float32)[rng.integers(0, 2, 3)]
        blobs.append((audio, video, label))
        H.write_tree(str(d / ("b%d.h5" % i)), {"audio": H.Chunked(audio, (1, 1, 48000)),
                                               "video": H.Chunked(video, (1, 56, 224, 3)), "label": label})
    gen = T.data_generator(str(d), batch_size=2, random_state=1)
    b0 = next(gen)
    assert b0["video"].dtype == np.uint8 and b0["audio"].dtype == np.int16 and b0["label"].shape == (2, 2)
    assert np.array_equal(b0["audio"], blobs[0][0][:2]) and np.array_equal(b0["video"], blobs[0][1][:2])
    b1 = next(gen)   # straddles the two files
    assert np.array_equal(b1["audio"], np.concatenate([blobs[0][0][2:], blobs[1][0][:1]]))


def test_keras_layout_weight_file(tmp_path):
    """save_weights(.h5) writes the keras 2.0.9 layout: root attrs layer_names / backend / keras_version, one group per
    top-level layer incl. weight-less ones, weight_names per group, datasets nested under '<variable scope>/'."""
    from l3embedding_b200 import model as M
    m, _, _ = M.MODELS["cnn_L3_melspec2"]()
    p = str(tmp_path / "model_latest.h5")
    m.save_weights(p)
    f = H.File(p)
    names = [n.decode() for n in f.attrs["layer_names"]]
    assert names == [l.name for l in m.layers] and f.attrs["keras_version"] == b"2.0.9"
    assert {"vision_model", "audio_model", "dense_1", "dense_2"} <= set(names)
    total = 0
    for ln in names:
        g = f[ln]
        for wn in g.attrs["weight_names"]:
            a = np.asarray(g[wn.decode()])
            assert a.dtype == np.float32
            total += 1
    assert total == len(m.get_weights())
    m2 = M.load_model(p, "cnn_L3_melspec2")
    assert all(np.array_equal(a, b) for a, b in zip(m.get_weights(), m2.get_weights()))


def test_writer_output_opens_with_h5py(tmp_path):
    """Interoperability of the built-in HDF5 WRITER with libhdf5: a keras-layout checkpoint written by minihdf5 is
    re-opened with h5py (skipped where h5py does not exist -- the build container; tools/convert_weights_h5py.py verify
    does the same from a shell)."""
    h5py = pytest.importorskip("h5py")
    from l3embedding_b200 import model as M
    m, _, _ = M.MODELS["cnn_L3_orig"]()
    p = str(tmp_path / "w.h5")
    import l3embedding_b200.weights_io as W
    names, arrays = m.weight_names(), m.get_weights()
    by = dict(zip(names, arrays))
    # force the built-in writer even though h5py is importable
    import builtins
    real_import = builtins.__import__

    def no_h5py(name, *a, **k):
        if name == "h5py":
            raise ImportError
        return real_import(name, *a, **k)
    builtins.__import__ = no_h5py
    try:
        W.save_weights(p, m)
    finally:
        builtins.__import__ = real_import
    with h5py.File(p, "r") as f:
        layers = [n.decode() for n in f.attrs["layer_names"]]
        assert layers == [l.name for l in m.layers]
        seen = 0
        for ln, layer in zip(layers, m.layers):
            wn = [n.decode() for n in f[ln].attrs["weight_names"]]
            assert len(wn) == len(layer._weight_names)
            for n, cn in zip(wn, layer._weight_names):
                assert np.array_equal(np.asarray(f[ln][n]), by[cn])
                seen += 1
        assert seen == len(names)
