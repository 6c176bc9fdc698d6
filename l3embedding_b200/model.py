"""Host-side mirror of the reference's model-builder interface (l3embedding/model.py), backed by libl3b200.so.

Same names, argument meaning and error behaviour as the reference:
    MODELS, construct_cnn_L3_{orig,kapredbinputbn,melspec1,melspec2}, gpu_wrapper      model.py:184-313
    load_model, load_embedding, convert_num_gpus                                       model.py:38-181
    convert_audio_model_to_embedding, construct_cnn_l3_orig_vision_embedding_model     audio_model.py:445, vision_model.py:198
The returned objects implement the subset of the keras Model API the reference's callers use:
compile / fit_generator / train_on_batch / predict / get_weights / set_weights / load_weights / save_weights /
layers / get_layer / to_json / name.  Building a model and moving weights around works without a GPU; anything
that computes requires the CUDA library and a device and raises L3Error otherwise (no CPU fallback).
"""
from __future__ import annotations

import json
import os
from collections import namedtuple
from typing import Dict, List, Optional

import numpy as np

from . import _lib, weights_io
from ._lib import L3Error

TensorSpec = namedtuple("TensorSpec", ["name", "shape"])

CONV_NAMES = ["1a", "1b", "2a", "2b", "3a", "3b", "4a", "4b"]
CONV_CH = [64, 64, 128, 128, 256, 256, 512, 512]
# audio_model.py:461-478
AUDIO_EMBED_POOL = {
    "cnn_L3_orig": {"original": (8, 8), "short": (32, 24)},
    "cnn_L3_kapredbinputbn": {"original": (8, 8), "short": (32, 24)},
    "cnn_L3_melspec1": {"original": (4, 8), "short": (16, 24)},
    "cnn_L3_melspec2": {"original": (8, 8), "short": (32, 24)},
}
AUDIO_FRONTEND = {  # audio_model.py:26-43,138-151,245-260,355-370
    "cnn_L3_orig": dict(n_dft=512, n_mels=None, input_bn=False),
    "cnn_L3_kapredbinputbn": dict(n_dft=512, n_mels=None, input_bn=True),
    "cnn_L3_melspec1": dict(n_dft=2048, n_mels=128, input_bn=True),
    "cnn_L3_melspec2": dict(n_dft=2048, n_mels=256, input_bn=True),
}
VISION_INPUT_BN = {"cnn_L3_orig": False, "cnn_L3_kapredbinputbn": True, "cnn_L3_melspec1": True, "cnn_L3_melspec2": True}


class Adam:
    """Stand-in for keras.optimizers.Adam(lr) (train.py:282); only the learning rate is configurable, the
    update rule is keras 2.0.9's (beta 0.9/0.999, eps 1e-8), fused on the device."""

    def __init__(self, lr=0.001, **kwargs):
        if kwargs:
            raise TypeError("unsupported Adam arguments: %s" % sorted(kwargs))
        self.lr = float(lr)


class History:
    def __init__(self):
        self.history: Dict[str, list] = {}
        self.epoch: List[int] = []


class _Layer:
    def __init__(self, name, weight_names=(), owner=None, prefix=None):
        self.name = name
        self._weight_names = list(weight_names)   # canonical names, trainable first then non-trainable
        self._owner = owner
        self.prefix = prefix

    def get_weights(self):
        w = self._owner._weights_dict()
        return [w[n] if not n.startswith("kapre/") else self._owner._kapre(n) for n in self._weight_names]

    def set_weights(self, arrays):
        self._owner._assign(self._weight_names, arrays)


def _bn_names(prefix):
    return ([prefix + "/gamma", prefix + "/beta"], [prefix + "/moving_mean", prefix + "/moving_variance"])


class TowerModel(_Layer):
    """'vision_model' / 'audio_model' nested model (a keras Container inside the AVC model)."""

    def __init__(self, owner, tower, model_type):
        super().__init__(tower + "_model", owner=owner)
        self.tower = tower
        self.model_type = model_type
        self.layers = []
        inp = _Layer("input_1" if tower == "vision" else "input_2", owner=owner)
        self.layers.append(inp)
        if tower == "audio":
            fe = AUDIO_FRONTEND[model_type]
            kn = ["kapre/real_kernels", "kapre/imag_kernels"] + (["kapre/freq2mel"] if fe["n_mels"] else [])
            self.layers.append(_Layer("melspectrogram_1" if fe["n_mels"] else "spectrogram_1", kn, owner))
            has_bn0 = fe["input_bn"]
        else:
            has_bn0 = VISION_INPUT_BN[model_type]
        if has_bn0:
            tr, nt = _bn_names(tower + "/bn0")
            self.layers.append(_Layer("batch_normalization_0", tr + nt, owner, tower + "/bn0"))
        for i, nm in enumerate(CONV_NAMES):
            lname = "conv2d_" + nm
            if nm == "4b":
                lname = tower + "_embedding_layer"     # audio_model.py:428, vision_model.py:182
            self.layers.append(_Layer(lname, [f"{tower}/conv{nm}/kernel", f"{tower}/conv{nm}/bias"], owner, f"{tower}/conv{nm}"))
            tr, nt = _bn_names(f"{tower}/bn{nm}")
            self.layers.append(_Layer("batch_normalization_" + nm, tr + nt, owner, f"{tower}/bn{nm}"))
        # Container.weights: all trainable (layer order) then all non-trainable (layer order)
        tr_all, nt_all = [], []
        for l in self.layers:
            for n in l._weight_names:
                (nt_all if (n.startswith("kapre/") or n.endswith(("moving_mean", "moving_variance"))) else tr_all).append(n)
        self._weight_names = tr_all + nt_all

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError("No such layer: " + name)

    def layerwise_weight_names(self):
        """keras Model.get_weights() order when called on the tower itself (layer by layer)."""
        return [n for l in self.layers for n in l._weight_names]


class L3Model:
    """The AVC model object `MODELS[model_type]()` returns (keras Model in the reference, model.py:32-35)."""

    def __init__(self, model_type: str, name: str, num_gpus: int = 0, seed: int = 20180123):
        if model_type not in _lib.MODEL_IDS:
            raise ValueError('Invalid model type: "{}"'.format(model_type))
        self.model_type = model_type
        self.name = name
        self.num_gpus = int(num_gpus)
        self._seed = seed
        self._host_weights: Optional[Dict[str, np.ndarray]] = None
        self._engine = None
        self._adam_host = None      # optimizer state waiting for the first training engine (resume)
        self._pending = 0           # batches staged on the device that no step has consumed yet
        self.dtype = os.environ.get("L3B200_DTYPE", "bf16")
        self.optimizer = None
        self.loss = None
        self.metrics_names = ["loss", "acc"]
        self.stop_training = False
        self.vision_model = TowerModel(self, "vision", model_type)
        self.audio_model = TowerModel(self, "audio", model_type)
        self.layers = [self.vision_model.layers[0], self.audio_model.layers[0], self.vision_model, self.audio_model,
                       _Layer("concatenate_1", owner=self),
                       _Layer("dense_1", ["dense_1/kernel", "dense_1/bias"], self, "dense_1"),
                       _Layer("dense_2", ["dense_2/kernel", "dense_2/bias"], self, "dense_2")]
        self.inputs = [TensorSpec("input_1", (None, 224, 224, 3)), TensorSpec("input_2", (None, 1, 48000))]
        self.outputs = [TensorSpec("dense_2/Softmax", (None, 2))]

    # ---- configuration ---------------------------------------------------------------------------------
    def configure(self, dtype: Optional[str] = None):
        """dtype 'f32' (parity mode), 'f32tc' (parity mode on tensor cores) or 'bf16' (tcgen05 throughput mode).
        Re-creates the device state."""
        if dtype is not None and dtype != self.dtype:
            if dtype not in _lib.DTYPES:
                raise ValueError("dtype must be one of %s" % sorted(_lib.DTYPES))
            self._pull()
            self._drop_engine()
            self.dtype = dtype
        return self

    # ---- weights ---------------------------------------------------------------------------------------
    def _weights_dict(self) -> Dict[str, np.ndarray]:
        if self._engine is not None:
            return self._engine.get_weights()
        if self._host_weights is None:
            self._host_weights = weights_io.he_normal_weights(self.model_type, self._seed)
        return self._host_weights

    def _pull(self):
        if self._engine is not None:
            self._host_weights = self._engine.get_weights()

    def _drop_engine(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None
    
    def _kapre(self, name):
        return weights_io.kapre_constants(self.model_type)[name]

    def _assign(self, names, arrays):
        arrays = list(arrays)
        if len(arrays) != len(names):
            raise ValueError("You called `set_weights(weights)` with a weight list of length %d, but the layer was "
                             "expecting %d weights." % (len(arrays), len(names)))
        w = dict(self._weights_dict())
        shapes = weights_io.weight_shapes(self.model_type)
        for n, a in zip(names, arrays):
            a = np.asarray(a, dtype=np.float32)
            if n.startswith("kapre/"):
                if tuple(a.shape) != tuple(self._kapre(n).shape):
                    raise ValueError("Layer weight shape %s not compatible with provided weight shape %s"
                                     % (self._kapre(n).shape, a.shape))
                continue  # constants of the on-device front-end; regenerated, not stored
            if tuple(a.shape) != tuple(shapes[n]):
                raise ValueError("Layer weight shape %s not compatible with provided weight shape %s" % (shapes[n], a.shape))
            w[n] = a
        self._host_weights = w
        if self._engine is not None:
            self._engine.set_weights(w)

    def weight_names(self) -> List[str]:
        """Order of keras Model.get_weights() on the AVC model (nested towers: trainable then non-trainable)."""
        return [n for l in self.layers for n in l._weight_names]

    def container_weight_names(self) -> List[str]:
        """Order of keras `Container.weights` of the whole AVC model: ALL trainable weights (layer order, nested towers
        flattened) and then all non-trainable ones -- the order in which the arrays sit in the single nested-model group
        of a checkpoint written from a `multi_gpu_model` (model.py:76-77,117-119)."""
        nt = lambda n: n.startswith("kapre/") or n.endswith(("moving_mean", "moving_variance"))
        names = self.weight_names()
        return [n for n in names if not nt(n)] + [n for n in names if nt(n)]

    def get_weights(self):
        w = self._weights_dict()
        return [w[n] if not n.startswith("kapre/") else self._kapre(n) for n in self.weight_names()]

    def set_weights(self, arrays):
        self._assign(self.weight_names(), arrays)

    def named_weights(self) -> Dict[str, np.ndarray]:
        return dict(self._weights_dict())

    def set_named_weights(self, w: Dict[str, np.ndarray]):
        names = [n for n in self.weight_names() if not n.startswith("kapre/")]
        self._assign(names, [w[n] for n in names])

    def save_weights(self, path):
        weights_io.save_weights(path, self)

    def load_weights(self, path):
        weights_io.load_weights(path, self)

    def count_params(self):
        return int(sum(int(np.prod(a.shape)) for a in self.get_weights()))

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError("No such layer: " + name)

    def to_json(self):
        return json.dumps(dict(class_name="Model", backend="l3embedding_b200",
                               config=dict(name=self.name, model_type=self.model_type, num_gpus=self.num_gpus,
                                           layers=[l.name for l in self.layers])))

    def get_config(self):
        return json.loads(self.to_json())["config"]

    # ---- device state ------------------------------------------------------------------------------------
    def _get_engine(self, batch: int, training: bool):
        """The ONE device engine of this model.  It is only ever replaced by a larger or more capable one (never
        downgraded from training to inference or to a smaller batch), and a replacement adopts weights, BN statistics,
        Adam moments and the step count device-to-device: keras keeps a single optimizer state for the whole fit
        (train.py:282,408), so validation with a bigger batch or a predict() between steps must not reset it."""
        from .engine import Engine  # imports torch
        e = self._engine
        if e is not None and e.dtype != self.dtype:
            self._pull()
            self._adam_host = e.get_adam_state() if e.training else self._adam_host
            self._drop_engine()
            e = None
        if e is not None and (e.training or not training) and e.max_batch >= batch:
            return e
        new = Engine(self.model_type, max_batch=max(batch, e.max_batch if e else 1, 1), dtype=self.dtype,
                     training=bool(training or (e is not None and e.training)), weights=None if e else self._weights_dict())
        if e is not None:
            new.adopt_state(e)
            e.close()
        elif new.training and self._adam_host is not None:
            new.set_adam_state(self._adam_host)      # restored from a checkpoint before the engine existed
            self._adam_host = None
        self._engine = new
        return new

    def get_optimizer_state(self):
        """Adam step count and moments ({'t','m','v'}, flat arena order) or None before the first training step."""
        if self._engine is not None and self._engine.training:
            return self._engine.get_adam_state()
        return self._adam_host

    def set_optimizer_state(self, state):
        if self._engine is not None and self._engine.training:
            self._engine.set_adam_state(state)
        else:
            self._adam_host = dict(t=np.int64(state["t"]), m=np.asarray(state["m"], np.float32),
                                   v=np.asarray(state["v"], np.float32))

    # ---- keras training API ------------------------------------------------------------------------------
    def compile(self, optimizer, loss="categorical_crossentropy", metrics=("accuracy",)):
        if loss != "categorical_crossentropy":
            raise ValueError("only loss='categorical_crossentropy' is implemented (train.py:270)")
        if not hasattr(optimizer, "lr"):
            raise TypeError("optimizer must expose .lr (use l3embedding_b200.model.Adam)")
        self.optimizer = optimizer
        self.loss = loss

    def _dp(self):
        from . import dp
        return dp.current(self.num_gpus)

    # The step is split in two so that fit_generator can overlap the upload of batch k+1 with step k:
    #   _stage(x, y)  -> asynchronous H2D of this rank's slice into a staging slot (library copy stream)
    #   _step(ticket) -> forward/backward on the staged slot, gradient all-reduce (N > 1), metrics, Adam
    def _stage(self, x, y):
        if self.optimizer is None:
            raise RuntimeError("You must compile a model before training/testing. Use `model.compile(optimizer, loss)`.")
        video, audio = x
        par = self._dp()
        B = len(video)
        sl = par.slice(B)
        n = sl.stop - sl.start
        if self._engine is not None and self._pending and (n > self._engine.max_batch or not self._engine.training):
            raise RuntimeError("batch of %d samples while a batch staged for a smaller engine is still pending: "
                               "batch sizes must not grow between consecutive training batches" % n)
        eng = self._get_engine(n, True)
        par.attach(eng)

        def host(a, kinds):
            a = np.asarray(a[sl])
            if a.dtype not in kinds:
                raise TypeError("unsupported input dtype %s (want one of %s)" % (a.dtype, [np.dtype(k).name for k in kinds]))
            return np.ascontiguousarray(a)
        v, a = host(video, (np.uint8, np.float32)), host(audio, (np.int16, np.float32))
        lab = np.ascontiguousarray(np.asarray(y[sl], dtype=np.float32))
        if v.shape[1:] != (224, 224, 3) or a.size != n * 48000 or lab.shape != (n, 2):
            raise ValueError("expected video (B,224,224,3), audio (B,1,48000), labels (B,2); got %s %s %s"
                             % (v.shape, a.shape, lab.shape))
        eng.upload_host(v, a, lab)
        self._pending += 1
        return dict(B=B, n=n, keep=(v, a, lab))     # the host arrays must outlive the asynchronous copy

    def _step(self, ticket):
        par, eng = self._dp(), self._engine
        B, n = ticket["B"], ticket["n"]
        self._pending -= 1
        if par.world_size == 1:
            m = eng.train_step_staged(n, self.optimizer.lr)
            return [m["loss"], m["acc"]]
        # data parallel: the library sums gradients and the two loss scalars over the ranks while backward runs
        m = eng.dp_train_step_staged(n, B, self.optimizer.lr)
        return [m["ce_sum"] / B + m["l2"], m["correct"] / B]

    def train_on_batch(self, x, y):
        """x = [video (B,224,224,3), audio (B,1,48000)]; float inputs in [-1,1] as the reference generator yields
        (train.py:186,189), or raw uint8 / int16 (scaled on the device).  Returns [loss, acc] (global batch)."""
        return self._step(self._stage(x, y))

    def test_on_batch(self, x, y):
        """keras test_on_batch (inference-mode BN).  Under data parallelism every rank evaluates its slice of the batch
        (training_utils.py:121-133) and the sums are all-reduced; a batch larger than the engine is run in chunks --
        the engine is never rebuilt for validation."""
        video, audio = x
        par = self._dp()
        B = len(video)
        sl = par.slice(B)
        n = sl.stop - sl.start
        eng = self._engine if self._engine is not None else self._get_engine(n, False)
        ce_sum = correct = 0.0
        l2 = 0.0
        for s0 in range(sl.start, sl.stop, eng.max_batch):
            s1 = min(s0 + eng.max_batch, sl.stop)
            eng.predict(video[s0:s1], audio[s0:s1], y[s0:s1])
            m = eng.metrics()
            ce_sum += m["ce_sum"]; correct += m["correct"]; l2 = m["l2"]
        ce_sum, correct = par.sum_scalars(ce_sum, correct)
        return [ce_sum / B + l2, correct / B]

    def sync_replicas(self):
        """Data parallel only: BN moving statistics are per replica during training (reference semantics: one momentum
        update per replica, training_utils.py:141-162); before validation / checkpointing they are replaced by their
        mean over the ranks so that every rank evaluates -- and rank 0 saves -- the same model."""
        par = self._dp()
        if par.world_size > 1 and self._engine is not None:
            par.average_bn_state(self._engine)

    def predict(self, x, batch_size=32, verbose=0):
        video, audio = x
        n = len(video)
        eng = self._engine if self._engine is not None else self._get_engine(min(batch_size, max(n, 1)), False)
        step = min(batch_size, eng.max_batch)
        out = np.empty((n, 2), np.float32)
        for s in range(0, n, step):
            out[s:s + step] = eng.predict(video[s:s + step], audio[s:s + step])[0]
        return out

    def evaluate_generator(self, generator, steps, **_):
        tot, loss, acc = 0, 0.0, 0.0
        for _i in range(steps):
            x, y = next(generator)[:2]
            l, a = self.test_on_batch(x, y)
            b = len(y)
            tot += b; loss += l * b; acc += a * b
        return [loss / max(tot, 1), acc / max(tot, 1)]

    def fit_generator(self, generator, steps_per_epoch, epochs=1, verbose=1, callbacks=None, validation_data=None,
                      validation_steps=None, initial_epoch=0, **_):
        """keras fit_generator as used at train.py:408-414: one train_on_batch per generator item, epoch logs
        {loss, acc, val_loss, val_acc} as batch-size-weighted means."""
        callbacks = list(callbacks or [])
        hist = History()
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            elif hasattr(cb, "model"):
                cb.model = self
        _call(callbacks, "on_train_begin")
        self.stop_training = False
        for epoch in range(initial_epoch, epochs):
            _call(callbacks, "on_epoch_begin", epoch)
            tot, loss, acc = 0, 0.0, 0.0
            # software pipeline: the next item is pulled from the generator and its upload enqueued BEFORE the current
            # step is run, so H2D (and the generator's own work) overlap the device step.  Within an epoch only: the
            # generator is not advanced past the epoch's last batch before validation, as in keras.
            item = next(generator)
            ticket = self._stage(item[0], item[1])
            for step in range(steps_per_epoch):
                b = ticket["B"]
                nxt = None
                if step + 1 < steps_per_epoch:
                    item = next(generator)
                    nxt = self._stage(item[0], item[1])
                _call(callbacks, "on_batch_begin", step, {"batch": step, "size": b})
                l, a = self._step(ticket)
                tot += b; loss += l * b; acc += a * b
                _call(callbacks, "on_batch_end", step, {"batch": step, "size": b, "loss": l, "acc": a})
                ticket = nxt
            logs = {"loss": loss / max(tot, 1), "acc": acc / max(tot, 1)}
            self.sync_replicas()
            if validation_data is not None:
                vl, va = self.evaluate_generator(validation_data, validation_steps)
                logs["val_loss"], logs["val_acc"] = vl, va
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            if verbose:
                print("Epoch %d/%d - %s" % (epoch + 1, epochs, " - ".join("%s: %.4f" % kv for kv in logs.items())))
            _call(callbacks, "on_epoch_end", epoch, logs)
            if self.stop_training:
                break
        _call(callbacks, "on_train_end")
        return hist


def _call(callbacks, hook, *args):
    for cb in callbacks:
        fn = getattr(cb, hook, None)
        if fn is not None:
            fn(*args) if args else fn()


class EmbeddingModel:
    """What load_embedding returns: `.predict(x)` maps (n,1,48000) audio to (n,6144|512) embeddings
    (or (n,224,224,3) frames to (n,8192)) with inference-mode BN, cut at the raw conv4b output."""

    def __init__(self, parent: L3Model, embedding_type: str, pooling_type: Optional[str]):
        self.parent = parent
        self.embedding_type = embedding_type
        self.pooling_type = pooling_type
        self.name = embedding_type + "_embedding_model"
        self._engine = None
        if embedding_type == "audio":
            eh, ew = {"cnn_L3_melspec1": (16, 24)}.get(parent.model_type, (32, 24))
            ph, pw = AUDIO_EMBED_POOL[parent.model_type][pooling_type]
            self.output_dim = (eh // ph) * (ew // pw) * 512
        else:
            self.output_dim = 4 * 4 * 512

    def _get_engine(self, batch):
        from .engine import Engine
        if self._engine is None or self._engine.max_batch < batch:
            if self._engine is not None:
                self._engine.close()
            self._engine = Engine(self.parent.model_type, max_batch=batch, dtype=self.parent.dtype, training=False,
                                  towers=(self.embedding_type,), weights=self.parent._weights_dict(), host_staging=False)
        return self._engine

    def predict_frames(self, signal, hop_length, batch_size=256):
        """Embeddings of all 1 s windows of a 1-D signal at `hop_length` samples, framed on the device."""
        if self.embedding_type != "audio":
            raise ValueError("predict_frames is an audio-embedding call")
        sig = np.ascontiguousarray(signal).reshape(-1)
        es = 2 if sig.dtype == np.int16 else 4
        if sig.dtype not in (np.int16, np.float32):
            sig = sig.astype(np.float32)
        n_frames = 1 + (len(sig) - 48000) // hop_length
        if (hop_length * es) % 16 != 0:      # unaligned frame starts: frame on the host as the reference does
            idx = np.arange(48000)[None, :] + hop_length * np.arange(n_frames)[:, None]
            return self.predict(sig[idx].reshape(n_frames, 1, 48000), batch_size=batch_size)
        eng = self._get_engine(min(batch_size, max(n_frames, 1)))
        return eng.embed_audio_frames(sig, hop_length, self.pooling_type).cpu().numpy()

    # device batch of predict(): keras' default batch_size of 32 (data/usc/features.py:304 passes none) would leave the
    # GPU idle between tiny launch sequences; inference results do not depend on how samples are batched (BN uses the
    # moving statistics), so larger chunks are used internally
    DEVICE_BATCH = 512

    def predict(self, x, batch_size=32, verbose=0):
        """keras `Model.predict`: (n,1,48000) audio [float32 in [-1,1) or int16 PCM] -> (n, 6144|512) float32, or
        (n,224,224,3) frames [float32 in [-1,1] or uint8] -> (n, 8192).  Host arrays are streamed through a three-stage
        pipeline: a worker thread copies chunk k+1 into page-locked memory while chunk k is on the device (H2D, towers,
        D2H all asynchronous on the engine's stream) and chunk k-1 is copied out of its pinned result buffer."""
        import threading

        import torch
        x = np.asarray(x)
        n = len(x)
        audio = self.embedding_type == "audio"
        want = (np.int16, np.float32) if audio else (np.uint8, np.float32)
        if x.dtype not in want:
            x = x.astype(np.float32)
        out = np.empty((n, self.output_dim), np.float32)
        if n == 0:
            return out
        step = min(max(int(batch_size), self.DEVICE_BATCH if audio else 64), n)
        eng = self._get_engine(step)
        step = min(step, eng.max_batch)
        dev = eng.device
        tdt = {np.dtype(np.int16): torch.int16, np.dtype(np.float32): torch.float32, np.dtype(np.uint8): torch.uint8}[x.dtype]
        sample = tuple(x.shape[1:])
        pin_in = [torch.empty((step,) + sample, dtype=tdt).pin_memory() for _ in range(2)]
        pin_out = [torch.empty((step, self.output_dim), dtype=torch.float32).pin_memory() for _ in range(2)]
        dev_in = [torch.empty((step,) + sample, dtype=tdt, device=dev) for _ in range(2)]
        dev_out = [torch.empty((step, self.output_dim), dtype=torch.float32, device=dev) for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]
        chunks = [(s, min(s + step, n)) for s in range(0, n, step)]
        staged = [threading.Event() for _ in chunks]
        free = [threading.Semaphore(1), threading.Semaphore(1)]      # pinned input slot may be overwritten
        err = []

        def stage():
            try:
                for k, (a, b) in enumerate(chunks):
                    free[k % 2].acquire()
                    np.copyto(pin_in[k % 2].numpy()[:b - a], x[a:b])
                    staged[k].set()
            except BaseException as e:       # surfaced in the caller
                err.append(e)
                for ev in staged:
                    ev.set()
        worker = threading.Thread(target=stage, name="l3-predict-stage", daemon=True)
        worker.start()

        def drain(k):
            a, b = chunks[k]
            done[k % 2].synchronize()
            out[a:b] = pin_out[k % 2].numpy()[:b - a]
            free[k % 2].release()            # its H2D copy completed long before the D2H did
        with torch.cuda.device(dev):
            for k, (a, b) in enumerate(chunks):
                staged[k].wait()
                if err:
                    raise err[0]
                m = b - a
                dev_in[k % 2][:m].copy_(pin_in[k % 2][:m], non_blocking=True)
                if audio:
                    eng.embed_audio(dev_in[k % 2][:m], self.pooling_type, out=dev_out[k % 2])
                else:
                    eng.embed_vision(dev_in[k % 2][:m], out=dev_out[k % 2])
                pin_out[k % 2][:m].copy_(dev_out[k % 2][:m], non_blocking=True)
                done[k % 2].record()
                if k >= 1:
                    drain(k - 1)
            drain(len(chunks) - 1)
        worker.join()
        return out


# ---- builders (model.py:184-313) -----------------------------------------------------------------------------

def gpu_wrapper(model_f):
    """Decorator for creating multi-gpu models (model.py:184-195).  With num_gpus > 1 the model trains
    data-parallel: one process per GPU (torchrun), each taking its contiguous slice of every batch
    (training_utils.py:121-133) and all-reducing gradients over NCCL."""
    def wrapped(num_gpus=0, *args, **kwargs):
        m, inp, out = model_f(*args, **kwargs)
        if num_gpus > 1:
            m = multi_gpu_model(m, gpus=num_gpus)
        return m, inp, out
    wrapped.__name__ = model_f.__name__
    wrapped.__doc__ = model_f.__doc__
    return wrapped


def multi_gpu_model(model, gpus):
    """training_utils.py:21 -- here a flag on the model; the replicas are the torchrun ranks."""
    if gpus <= 1:
        raise ValueError("For multi-gpu usage to be effective, call `multi_gpu_model` with `gpus >= 2`. "
                         "Received: `gpus=%d`" % gpus)
    model.num_gpus = int(gpus)
    return model


def _construct(model_type, name):
    m = L3Model(model_type, name)
    return m, list(m.inputs), m.outputs[0]


# ---- tower sub-builders (audio_model.py:8,118,225,335,490; vision_model.py:7,102,221) and the merge (model.py:7-35) --

class TowerHandle(TowerModel):
    """What the reference's tower builders return as `m`: a stand-alone description of one tower (layer names, weight
    inventory) that only L3_merge_audio_vision_models can turn into something that computes -- the library implements
    the four tower pairings the reference's MODELS registry builds, as whole models."""

    def __init__(self, tower, variant, model_type):
        super().__init__(None, tower, model_type)
        self.variant = variant


# (vision variant, audio variant) -> model type, exactly the pairings of model.py:198-284
_PAIRINGS = {("orig", "orig"): "cnn_L3_orig", ("inputbn", "kapredbinputbn"): "cnn_L3_kapredbinputbn",
             ("inputbn", "melspec1"): "cnn_L3_melspec1", ("inputbn", "melspec2"): "cnn_L3_melspec2"}
_X_I = TensorSpec("input_1", (None, 224, 224, 3))
_X_A = TensorSpec("input_2", (None, 1, 48000))


def _tower(tower, variant, model_type):
    m = TowerHandle(tower, variant, model_type)
    return m, (_X_A if tower == "audio" else _X_I), TensorSpec(tower + "_model/flatten", (None, 512))


def construct_cnn_L3_orig_audio_model():
    """audio_model.py:8-115: Spectrogram(n_dft 512, 'valid'), log(max(x, 1e-12))/5, no input BN."""
    return _tower("audio", "orig", "cnn_L3_orig")


def construct_cnn_L3_kapredbinputbn_audio_model():
    """audio_model.py:118-223: Spectrogram dB + input batch normalisation."""
    return _tower("audio", "kapredbinputbn", "cnn_L3_kapredbinputbn")


def construct_cnn_L3_melspec1_audio_model():
    """audio_model.py:225-332: Melspectrogram(n_dft 2048, 128 mels, 'same') dB + input BN."""
    return _tower("audio", "melspec1", "cnn_L3_melspec1")


def construct_cnn_L3_melspec2_audio_model():
    """audio_model.py:335-442: Melspectrogram(n_dft 2048, 256 mels, 'same') dB + input BN."""
    return _tower("audio", "melspec2", "cnn_L3_melspec2")


def construct_cnn_L3_orig_vision_model():
    """vision_model.py:7-99."""
    return _tower("vision", "orig", "cnn_L3_orig")


def construct_cnn_L3_orig_inputbn_vision_model():
    """vision_model.py:102-195 (input batch normalisation)."""
    return _tower("vision", "inputbn", "cnn_L3_melspec2")


def construct_tiny_L3_audio_model():
    raise NotImplementedError("tiny_L3 passes n_win to kapre.Spectrogram, which stock kapre 0.1.3.1/0.1.4 rejects "
                              "(audio_model.py:516); it is not part of the B200 path")


def construct_tiny_L3_vision_model():
    raise NotImplementedError("tiny_L3 is not part of the B200 path (see construct_tiny_L3_audio_model)")


def L3_merge_audio_vision_models(vision_model, x_i, audio_model, x_a, model_name, layer_size=128):
    """model.py:7-35: concatenate([vision, audio]) -> Dense(layer_size, relu) -> Dense(2, softmax), l2(1e-5) on the
    kernels.  Returns (model, [x_i, x_a], y).  The device library implements the reference's own pairings and head
    width; anything else raises."""
    if not isinstance(vision_model, TowerHandle) or not isinstance(audio_model, TowerHandle) \
            or vision_model.tower != "vision" or audio_model.tower != "audio":
        raise TypeError("L3_merge_audio_vision_models expects the towers of construct_*_vision_model / _audio_model")
    key = (vision_model.variant, audio_model.variant)
    if key not in _PAIRINGS:
        raise ValueError("unsupported tower pairing %s: the reference builds %s" % (key, sorted(_PAIRINGS)))
    if layer_size != 128:
        raise ValueError("layer_size=%d: the reference's models (and libl3b200) use 128" % layer_size)
    m = L3Model(_PAIRINGS[key], model_name)
    return m, [x_i, x_a], m.outputs[0]


@gpu_wrapper
def construct_cnn_L3_orig():
    """Original L3 model (model.py:198-218)."""
    return _construct("cnn_L3_orig", "cnn_L3_orig")


@gpu_wrapper
def construct_cnn_L3_kapredbinputbn():
    """L3 with kapre dB spectrogram and input batch normalisation (model.py:220-240)."""
    return _construct("cnn_L3_kapredbinputbn", "cnn_L3_kapredbinputbn")


@gpu_wrapper
def construct_cnn_L3_melspec1():
    """L3 with a 128-band mel front-end (model.py:242-262)."""
    return _construct("cnn_L3_melspec1", "cnn_L3_melspec1")


@gpu_wrapper
def construct_cnn_L3_melspec2():
    """L3 with a 256-band mel front-end (model.py:264-284)."""
    return _construct("cnn_L3_melspec2", "cnn_L3_melspec2")


def construct_tiny_L3(*_a, **_k):
    raise NotImplementedError("tiny_L3 passes n_win to kapre.Spectrogram, which stock kapre 0.1.3.1/0.1.4 rejects "
                              "(audio_model.py:516); it is not part of the B200 path")


MODELS = {
    "cnn_L3_orig": construct_cnn_L3_orig,
    "tiny_L3": construct_tiny_L3,
    "cnn_L3_kapredbinputbn": construct_cnn_L3_kapredbinputbn,
    "cnn_L3_melspec1": construct_cnn_L3_melspec1,
    "cnn_L3_melspec2": construct_cnn_L3_melspec2,
}


def convert_num_gpus(model, inputs, outputs, model_type, src_num_gpus, tgt_num_gpus):
    """model.py:38-82.  Checkpoints written by this package are always in the single-model layout."""
    if src_num_gpus <= 1 and tgt_num_gpus <= 1:
        return model, inputs, outputs
    m_new, inputs_new, output_new = MODELS[model_type]()
    m_new.set_weights(model.get_weights())
    m_new.dtype = model.dtype
    if tgt_num_gpus > 1:
        m_new = multi_gpu_model(m_new, gpus=tgt_num_gpus)
    return m_new, inputs_new, output_new


def load_model(weights_path, model_type, src_num_gpus=0, tgt_num_gpus=None, return_io=False):
    """Loads an audio-visual correspondence model (model.py:85-128)."""
    if model_type not in MODELS:
        raise ValueError('Invalid model type: "{}"'.format(model_type))
    m, inputs, output = MODELS[model_type]()
    if src_num_gpus > 1:
        m = multi_gpu_model(m, gpus=src_num_gpus)
    m.load_weights(weights_path)
    if tgt_num_gpus is not None and src_num_gpus != tgt_num_gpus:
        m, inputs, output = convert_num_gpus(m, inputs, output, model_type, src_num_gpus, tgt_num_gpus)
    if return_io:
        return m, inputs, output
    return m


def convert_audio_model_to_embedding(audio_model, x_a, model_type, pooling_type="original"):
    """audio_model.py:445-487: MaxPooling2D(pool,'same') over the raw 'audio_embedding_layer' output + Flatten."""
    pool_size = AUDIO_EMBED_POOL[model_type][pooling_type]   # KeyError on unknown keys, as in the reference
    del pool_size
    audio_model.get_layer("audio_embedding_layer")
    m = EmbeddingModel(audio_model._owner, "audio", pooling_type)
    return m, x_a, TensorSpec("audio_embedding", (None, m.output_dim))


def construct_cnn_l3_orig_vision_embedding_model(vision_model, x_i):
    """vision_model.py:198-218: MaxPooling2D((7,7),'same') over the raw 'vision_embedding_layer' output + Flatten."""
    vision_model.get_layer("vision_embedding_layer")
    m = EmbeddingModel(vision_model._owner, "vision", None)
    return m, x_i, TensorSpec("vision_embedding", (None, m.output_dim))


def load_embedding(weights_path, model_type, embedding_type, pooling_type, src_num_gpus=0, tgt_num_gpus=None,
                   return_io=False):
    """Loads an embedding model (model.py:131-181)."""
    m, inputs, output = load_model(weights_path, model_type, src_num_gpus=src_num_gpus, tgt_num_gpus=tgt_num_gpus,
                                   return_io=True)
    x_i, x_a = inputs
    if embedding_type == "vision":
        m_embed, x_embed, y_embed = construct_cnn_l3_orig_vision_embedding_model(m.get_layer("vision_model"), x_i)
    elif embedding_type == "audio":
        m_embed, x_embed, y_embed = convert_audio_model_to_embedding(m.get_layer("audio_model"), x_a, model_type,
                                                                     pooling_type)
    else:
        raise ValueError('Invalid embedding type: "{}"'.format(embedding_type))
    if return_io:
        return m_embed, x_embed, y_embed
    return m_embed
