"""Device-side state of one L3 model replica: flat arenas + workspace (torch tensors used as storage only) and
the calls into libl3b200.so that replace keras `train_on_batch` / `predict` (l3embedding/train.py:408-414,
data/usc/features.py:304).  All arithmetic happens in the CUDA library; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import L3Error
from .weights_io import he_normal_weights

SR = 48000


def _as_device(x, device, dtypes):
    """numpy / torch (host or device) -> contiguous device tensor of one of `dtypes` (no value conversion)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        raise TypeError("expected numpy array or torch tensor, got %r" % type(x))
    if x.dtype not in dtypes:
        raise TypeError("unsupported dtype %s (want one of %s)" % (x.dtype, dtypes))
    return x.to(device, non_blocking=True).contiguous()


class Engine:
    """One replica on one GPU.  dtype 'f32' = parity mode (fp32 storage, fp32 SIMT convolutions);
    'bf16' = throughput mode (bf16 activations, tcgen05 convolutions with fp32 accumulation);
    'f32tc' = parity mode on tensor cores (fp32 storage, split 16-bit operands on tcgen05; include/l3b200.h)."""

    def __init__(self, model_type: str, max_batch: int, dtype: str = "f32", training: bool = True,
                 towers=("vision", "audio"), device: Optional[torch.device] = None, host_staging: bool = True,
                 weights: Optional[Dict[str, np.ndarray]] = None, seed: int = 20180123):
        self.lib = _lib.load()
        self.model_type = model_type
        self.mid = _lib.model_id(model_type)
        if dtype not in _lib.DTYPES:
            raise ValueError("dtype must be one of %s" % sorted(_lib.DTYPES))
        if not torch.cuda.is_available():
            raise L3Error("no CUDA device: the L3 B200 path has no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.dtype = dtype
        self.training = bool(training)
        self.max_batch = int(max_batch)
        self.table = _lib.tensor_table(model_type)
        self.index = {n: (a, o, s) for n, a, o, s in self.table}
        self.n_params = int(self.lib.l3_param_count(self.mid))
        self.n_l2 = int(self.lib.l3_l2_count(self.mid))
        self.n_state = int(self.lib.l3_state_count(self.mid))
        self.flags = (_lib.WS_TRAINING if training else 0) | (_lib.WS_HOST_STAGING if host_staging else 0)
        if "vision" in towers:
            self.flags |= _lib.WS_VISION
        if "audio" in towers:
            self.flags |= _lib.WS_AUDIO
        dt = _lib.DTYPES[dtype]
        with torch.cuda.device(self.device):
            f32 = dict(dtype=torch.float32, device=self.device)
            self.params = torch.zeros(self.n_params, **f32)
            self.bn_state = torch.zeros(self.n_state, **f32)
            self.grads = torch.zeros(self.n_params, **f32) if training else None
            self.adam_m = torch.zeros(self.n_params, **f32) if training else None
            self.adam_v = torch.zeros(self.n_params, **f32) if training else None
            ws_bytes = _lib.check(self.lib.l3_workspace_bytes(self.mid, self.max_batch, dt, self.flags), "l3_workspace_bytes")
            self.workspace = torch.empty(ws_bytes + 256, dtype=torch.uint8, device=self.device)
            base = self.workspace.data_ptr()
            self._ws_ptr = (base + 255) // 256 * 256
            self.stream = torch.cuda.current_stream(self.device)
            ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
            self.ctx = self.lib.l3_ctx_create(self.mid, self.max_batch, dt, self.flags, ptr(self.params), ptr(self.grads),
                                              ptr(self.adam_m), ptr(self.adam_v), ptr(self.bn_state),
                                              C.c_void_p(self._ws_ptr), ws_bytes, C.c_void_p(self.stream.cuda_stream))
            if not self.ctx:
                raise L3Error("l3_ctx_create failed: " + _lib.last_error())
        self.set_weights(weights if weights is not None else he_normal_weights(model_type, seed))
        n_out, n_frames = C.c_int(), C.c_int()
        self.lib.l3_frontend_shape(self.mid, C.byref(n_out), C.byref(n_frames))
        self.frontend_shape = (n_out.value, n_frames.value)
        eh, ew = C.c_int(), C.c_int()
        self.lib.l3_embedding_map_shape(self.mid, C.byref(eh), C.byref(ew))
        self.embedding_map_shape = (eh.value, ew.value)

    def close(self):
        if getattr(self, "ctx", None):
            torch.cuda.synchronize(self.device)
            self.lib.l3_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -----------------------------------------------------------------------------------------
    def _arena(self, arena: int, grads: bool = False):
        if arena == 1:
            return self.bn_state
        return self.grads if grads else self.params

    def set_weights(self, weights: Dict[str, np.ndarray]):
        for name, (arena, off, shape) in self.index.items():
            if name not in weights:
                raise KeyError("missing weight '%s'" % name)
            w = np.asarray(weights[name], dtype=np.float32)
            if tuple(w.shape) != tuple(shape):
                raise ValueError("weight '%s' has shape %s, expected %s" % (name, w.shape, shape))
            n = int(np.prod(shape))
            self._arena(arena)[off:off + n].copy_(torch.from_numpy(np.ascontiguousarray(w).reshape(-1)))

    def get_weights(self) -> Dict[str, np.ndarray]:
        torch.cuda.synchronize(self.device)
        p, s = self.params.cpu().numpy(), self.bn_state.cpu().numpy()
        out = {}
        for name, (arena, off, shape) in self.index.items():
            src = s if arena == 1 else p
            out[name] = src[off:off + int(np.prod(shape))].reshape(shape).copy()
        return out

    def get_grads(self) -> Dict[str, np.ndarray]:
        """Gradients of the mean cross-entropy (the l2 term is applied inside l3_adam_step)."""
        torch.cuda.synchronize(self.device)
        g = self.grads.cpu().numpy()
        return {name: g[off:off + int(np.prod(shape))].reshape(shape).copy()
                for name, (arena, off, shape) in self.index.items() if arena == 0}

    # ---- hot path ----------------------------------------------------------------------------------------
    def _inputs(self, video, audio, labels=None):
        v = _as_device(video, self.device, (torch.uint8, torch.float32)) if video is not None else None
        a = _as_device(audio, self.device, (torch.int16, torch.float32)) if audio is not None else None
        lab = _as_device(labels, self.device, (torch.float32,)) if labels is not None else None
        B = (v if v is not None else a).shape[0]
        if v is not None and tuple(v.shape) != (B, 224, 224, 3):
            raise ValueError("video must be (B,224,224,3), got %s" % (tuple(v.shape),))
        if a is not None and a.numel() != B * SR:
            raise ValueError("audio must be (B,1,48000), got %s" % (tuple(a.shape),))
        if lab is not None and tuple(lab.shape) != (B, 2):
            raise ValueError("labels must be (B,2), got %s" % (tuple(lab.shape),))
        if B > self.max_batch:
            raise ValueError("batch %d exceeds max_batch %d" % (B, self.max_batch))
        vf = _lib.VIDEO_U8 if (v is not None and v.dtype == torch.uint8) else _lib.VIDEO_F32
        af = _lib.AUDIO_I16 if (a is not None and a.dtype == torch.int16) else _lib.AUDIO_F32
        return v, vf, a, af, lab, B

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def forward_backward(self, video, audio, labels, global_batch: Optional[int] = None):
        with torch.cuda.device(self.device):
            v, vf, a, af, lab, B = self._inputs(video, audio, labels)
            self._keep = (v, a, lab)  # keep inputs alive until the stream has consumed them
            _lib.check(self.lib.l3_forward_backward(self.ctx, self._p(v), vf, self._p(a), af, self._p(lab), B,
                                                    int(global_batch or B)), "l3_forward_backward")
        return B

    def adam_step(self, lr: float):
        _lib.check(self.lib.l3_adam_step(self.ctx, float(lr)), "l3_adam_step")

    def set_adam_t(self, t: int):
        _lib.check(self.lib.l3_adam_set_t(self.ctx, int(t)), "l3_adam_set_t")

    def get_adam_t(self) -> int:
        return int(_lib.check(self.lib.l3_adam_get_t(self.ctx), "l3_adam_get_t"))

    # ---- optimizer state (keras keeps ONE Adam state for the whole fit; so must every rebuild / resume) -----
    def get_adam_state(self) -> Dict[str, np.ndarray]:
        """{'t': step count, 'm': first moments, 'v': second moments} as flat host arrays in arena order."""
        if not self.training:
            raise L3Error("inference engine holds no optimizer state")
        torch.cuda.synchronize(self.device)
        return {"t": np.int64(self.get_adam_t()), "m": self.adam_m.cpu().numpy(), "v": self.adam_v.cpu().numpy()}

    def set_adam_state(self, state) -> None:
        if not self.training:
            raise L3Error("inference engine holds no optimizer state")
        m, v = np.asarray(state["m"], np.float32).reshape(-1), np.asarray(state["v"], np.float32).reshape(-1)
        if m.size != self.n_params or v.size != self.n_params:
            raise ValueError("optimizer state of %d / %d scalars does not fit a %s model (%d parameters)"
                             % (m.size, v.size, self.model_type, self.n_params))
        self.adam_m.copy_(torch.from_numpy(m))
        self.adam_v.copy_(torch.from_numpy(v))
        self.set_adam_t(int(state["t"]))

    def adopt_state(self, other: "Engine") -> None:
        """Take over weights, BN statistics and -- when both sides train -- the Adam moments and step count of another
        engine of the same model (device-to-device), e.g. before replacing it with a larger one."""
        if other.model_type != self.model_type:
            raise ValueError("cannot adopt the state of a %s engine" % other.model_type)
        torch.cuda.synchronize(other.device)
        self.params.copy_(other.params)
        self.bn_state.copy_(other.bn_state)
        if self.training and other.training:
            self.adam_m.copy_(other.adam_m)
            self.adam_v.copy_(other.adam_v)
            self.set_adam_t(other.get_adam_t())

    def metrics(self) -> Dict[str, float]:
        """Synchronises.  ce = mean cross-entropy over the local batch; loss = ce + l2 penalty (keras `loss`)."""
        out = (C.c_float * 4)()
        _lib.check(self.lib.l3_get_metrics(self.ctx, out), "l3_get_metrics")
        n = max(out[3], 1.0)
        return dict(ce_sum=out[0], correct=out[1], l2=out[2], batch=out[3], ce=out[0] / n, acc=out[1] / n,
                    loss=out[0] / n + out[2])

    def train_step_host(self, video: np.ndarray, audio: np.ndarray, labels: np.ndarray, lr: float) -> Dict[str, float]:
        """One keras train_on_batch from HOST buffers: H2D + forward + backward + Adam, metrics read back."""
        video = np.ascontiguousarray(video)
        audio = np.ascontiguousarray(audio)
        labels = np.ascontiguousarray(labels, dtype=np.float32)
        B = video.shape[0]
        vf = _lib.VIDEO_U8 if video.dtype == np.uint8 else _lib.VIDEO_F32
        af = _lib.AUDIO_I16 if audio.dtype == np.int16 else _lib.AUDIO_F32
        if video.dtype not in (np.uint8, np.float32) or audio.dtype not in (np.int16, np.float32):
            raise TypeError("video must be uint8|float32 and audio int16|float32")
        out = (C.c_float * 4)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.l3_train_step_host(self.ctx, video.ctypes.data_as(C.c_void_p), vf,
                                                   audio.ctypes.data_as(C.c_void_p), af,
                                                   labels.ctypes.data_as(C.c_void_p), B, float(lr), out),
                       "l3_train_step_host")
        n = max(out[3], 1.0)
        return dict(ce=out[0] / n, acc=out[1] / n, l2=out[2], loss=out[0] / n + out[2], batch=out[3])

    def upload_host(self, video: np.ndarray, audio: np.ndarray, labels: Optional[np.ndarray]) -> int:
        """Asynchronous H2D of one batch from host (ideally pinned) memory into one of the engine's two staging slots, on
        the library's copy stream.  Returns the batch size.  At most two uploads may be pending; they are consumed in
        order by train_step_staged / forward_backward_staged.  Safe to call from a prefetch thread while another
        thread runs a step.  The arrays must stay alive (and unmodified) until the consuming step has returned."""
        if video.dtype not in (np.uint8, np.float32) or audio.dtype not in (np.int16, np.float32):
            raise TypeError("video must be uint8|float32 and audio int16|float32")
        if not (video.flags.c_contiguous and audio.flags.c_contiguous):
            raise ValueError("upload_host needs C-contiguous arrays")
        B = video.shape[0]
        lab = None
        if labels is not None:
            if labels.dtype != np.float32 or not labels.flags.c_contiguous:
                raise ValueError("labels must be a C-contiguous float32 array")
            lab = labels.ctypes.data_as(C.c_void_p)
        vf = _lib.VIDEO_U8 if video.dtype == np.uint8 else _lib.VIDEO_F32
        af = _lib.AUDIO_I16 if audio.dtype == np.int16 else _lib.AUDIO_F32
        _lib.check(self.lib.l3_upload_batch_host(self.ctx, video.ctypes.data_as(C.c_void_p), vf,
                                                 audio.ctypes.data_as(C.c_void_p), af, lab, B), "l3_upload_batch_host")
        return B

    def train_step_staged(self, batch: int, lr: float) -> Dict[str, float]:
        """keras train_on_batch on the oldest staged batch (forward + backward + Adam; one synchronisation)."""
        out = (C.c_float * 4)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.l3_train_step_staged(self.ctx, int(batch), float(lr), out), "l3_train_step_staged")
        n = max(out[3], 1.0)
        return dict(ce=out[0] / n, acc=out[1] / n, l2=out[2], loss=out[0] / n + out[2], batch=out[3],
                    ce_sum=out[0], correct=out[1])

    def forward_backward_staged(self, batch: int, global_batch: Optional[int] = None):
        """forward + backward on the oldest staged batch (the data-parallel step adds the all-reduce and Adam)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.l3_forward_backward(self.ctx, None, 0, None, 0, None, int(batch),
                                                    int(global_batch or batch)), "l3_forward_backward")

    # ---- data parallelism (l3_dp_*: NCCL inside the library, gradient buckets overlapped with backward) -------
    @staticmethod
    def dp_unique_id() -> bytes:
        """128-byte NCCL unique id (rank 0 creates it and hands it to the other ranks)."""
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().l3_dp_unique_id(buf), "l3_dp_unique_id")
        return buf.raw

    def dp_init(self, unique_id: bytes, rank: int, world_size: int):
        """Collective over all ranks: joins this engine's context to the data-parallel job."""
        if len(unique_id) != 128:
            raise ValueError("the NCCL unique id has 128 bytes")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.l3_dp_init(self.ctx, unique_id, int(rank), int(world_size)), "l3_dp_init")

    @property
    def dp_world(self):
        """(rank, world size) once dp_init has run, else None."""
        r, n = C.c_int(), C.c_int()
        on = _lib.check(self.lib.l3_dp_info(self.ctx, C.byref(r), C.byref(n)), "l3_dp_info")
        return (r.value, n.value) if on else None

    def dp_train_step_staged(self, batch: int, global_batch: int, lr: float) -> Dict[str, float]:
        """train_on_batch of the GLOBAL batch on this rank's staged slice; metrics are sums over the global batch."""
        out = (C.c_float * 4)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.l3_dp_train_step_staged(self.ctx, int(batch), int(global_batch), float(lr), out),
                       "l3_dp_train_step_staged")
        n = max(out[3], 1.0)
        return dict(ce=out[0] / n, acc=out[1] / n, l2=out[2], loss=out[0] / n + out[2], batch=out[3],
                    ce_sum=out[0], correct=out[1])

    def dp_average_bn_state(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.l3_dp_average_bn_state(self.ctx), "l3_dp_average_bn_state")

    def train_stream(self, batches, lr: float):
        """Pipelined training over an iterable of host batches (video, audio, labels): the upload of batch k+1 is in
        flight on the copy stream while step k computes.  Yields the metrics of every step."""
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        n_cur = self.upload_host(*cur)
        while cur is not None:
            nxt = next(it, None)
            n_nxt = self.upload_host(*nxt) if nxt is not None else 0
            yield self.train_step_staged(n_cur, lr)
            cur, n_cur = nxt, n_nxt

    def predict(self, video, audio, labels=None):
        """Inference-mode forward (BN moving statistics): returns (probs, logits) as (B,2) numpy arrays."""
        with torch.cuda.device(self.device):
            v, vf, a, af, lab, B = self._inputs(video, audio, labels)
            probs = torch.empty(B, 2, dtype=torch.float32, device=self.device)
            logits = torch.empty(B, 2, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.l3_predict(self.ctx, self._p(v), vf, self._p(a), af, self._p(lab), B, self._p(probs),
                                           self._p(logits)), "l3_predict")
            return probs.cpu().numpy(), logits.cpu().numpy()

    def embed_audio(self, audio, pooling: str = "original", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """audio (n,1,48000) -> (n, 6144|512) device tensor (load_embedding 'audio' + predict)."""
        if pooling not in ("original", "short"):
            raise KeyError(pooling)
        with torch.cuda.device(self.device):
            _, _, a, af, _, n = self._inputs(None, audio)
            eh, ew = self.embedding_map_shape
            full = {"cnn_L3_melspec1": {"original": (4, 8), "short": (16, 24)}}.get(
                self.model_type, {"original": (8, 8), "short": (32, 24)})[pooling]
            dim = (eh // full[0]) * (ew // full[1]) * 512
            if out is None:
                out = torch.empty(n, dim, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.l3_embed_audio(self.ctx, self._p(a), af, n,
                                               _lib.POOL_ORIGINAL if pooling == "original" else _lib.POOL_SHORT,
                                               self._p(out)), "l3_embed_audio")
            self._keep = (a,)
            return out

    def embedding_dim(self, pooling: str) -> int:
        eh, ew = self.embedding_map_shape
        ph, pw = {"cnn_L3_melspec1": {"original": (4, 8), "short": (16, 24)}}.get(
            self.model_type, {"original": (8, 8), "short": (32, 24)})[pooling]
        return (eh // ph) * (ew // pw) * 512

    def embed_audio_frames(self, signal, hop: int, pooling: str = "original") -> torch.Tensor:
        """Embeddings of every 1 s window of a 1-D signal at `hop` samples (get_l3_frames_uniform,
        data/usc/features.py:279-304) -- the overlapping windows are read in place on the device."""
        with torch.cuda.device(self.device):
            sig = _as_device(signal, self.device, (torch.int16, torch.float32)).reshape(-1)
            n = sig.numel()
            if n < SR:
                raise ValueError("signal shorter than one 1 s frame")
            n_frames = 1 + (n - SR) // hop
            out = torch.empty(n_frames, self.embedding_dim(pooling), dtype=torch.float32, device=self.device)
            af = _lib.AUDIO_I16 if sig.dtype == torch.int16 else _lib.AUDIO_F32
            _lib.check(self.lib.l3_embed_audio_frames(self.ctx, self._p(sig), af, n, int(hop), n_frames,
                                                      _lib.POOL_ORIGINAL if pooling == "original" else _lib.POOL_SHORT,
                                                      self._p(out)), "l3_embed_audio_frames")
            self._keep = (sig,)
            return out

    def embed_vision(self, video, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        with torch.cuda.device(self.device):
            v, vf, _, _, _, n = self._inputs(video, None)
            if out is None:
                out = torch.empty(n, 4 * 4 * 512, dtype=torch.float32, device=self.device)
            _lib.check(self.lib.l3_embed_vision(self.ctx, self._p(v), vf, n, self._p(out)), "l3_embed_vision")
            self._keep = (v,)
            return out

    def frontend(self, audio) -> torch.Tensor:
        with torch.cuda.device(self.device):
            _, _, a, af, _, n = self._inputs(None, audio)
            out = torch.empty(n, self.frontend_shape[0], self.frontend_shape[1], dtype=torch.float32, device=self.device)
            _lib.check(self.lib.l3_frontend_fwd(self.ctx, self._p(a), af, n, self._p(out)), "l3_frontend_fwd")
            self._keep = (a,)
            return out

    def debug_read(self, which: str, batch: int, cap: int = 1 << 26) -> np.ndarray:
        """Copy an internal activation ('audio/z3', 'vision/a1', 'audio/x0', 'concat', ...) to the host as float32."""
        buf = np.empty(cap, dtype=np.float32)
        n = self.lib.l3_debug_read(self.ctx, which.encode(), int(batch), buf.ctypes.data_as(C.POINTER(C.c_float)), cap)
        _lib.check(int(n), "l3_debug_read(%s)" % which)
        return buf[:n].copy()

    def set_two_streams(self, enable: bool):
        """Overlap the two towers on two streams (default) or serialise them (per-kernel timing)."""
        _lib.check(self.lib.l3_ctx_set_two_streams(self.ctx, int(bool(enable))), "l3_ctx_set_two_streams")

    def profile(self, enable: bool):
        _lib.check(self.lib.l3_ctx_profile_enable(self.ctx, int(bool(enable))), "l3_ctx_profile_enable")

    def profile_read(self):
        """{class: (milliseconds, launches)} accumulated since the last read (synchronises)."""
        ms, n = (C.c_float * 4)(), (C.c_int * 4)()
        _lib.check(self.lib.l3_ctx_profile_read(self.ctx, ms, n), "l3_ctx_profile_read")
        return {k: (ms[i], n[i]) for i, k in enumerate(("conv_fwd", "conv_dgrad", "conv_wgrad", "frontend"))}

    @property
    def uses_tensor_cores(self) -> bool:
        return bool(self.lib.l3_ctx_uses_tensor_cores(self.ctx))

    def set_fused_inference(self, enable: bool):
        """Inference on the tensor-core path: BN + ReLU in the convolution epilogue (default) or layer by layer."""
        _lib.check(self.lib.l3_ctx_set_fused_inference(self.ctx, int(bool(enable))), "l3_ctx_set_fused_inference")

    def set_use_tensor_cores(self, enable: bool):
        _lib.check(self.lib.l3_ctx_set_use_tensor_cores(self.ctx, int(bool(enable))), "l3_ctx_set_use_tensor_cores")
