"""ctypes binding of libl3b200.so (include/l3b200.h).  There is no CPU fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libl3b200.so")

MODEL_IDS = {"cnn_L3_orig": 0, "cnn_L3_kapredbinputbn": 1, "cnn_L3_melspec1": 2, "cnn_L3_melspec2": 3}
DTYPE_F32, DTYPE_BF16, DTYPE_F32TC = 0, 1, 2
DTYPES = {"f32": DTYPE_F32, "bf16": DTYPE_BF16, "f32tc": DTYPE_F32TC}
VIDEO_U8, VIDEO_F32 = 0, 1
AUDIO_I16, AUDIO_F32 = 0, 1
WS_TRAINING, WS_VISION, WS_AUDIO, WS_HOST_STAGING = 1, 2, 4, 8
POOL_ORIGINAL, POOL_SHORT = 0, 1


class L3Error(RuntimeError):
    pass


_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_fp = C.POINTER(C.c_float)

# name -> (restype, argtypes); every symbol include/l3b200.h declares
SIGNATURES = {
    "l3_version": (_i, []),
    "l3_last_error": (C.c_char_p, []),
    "l3_param_count": (_i64, [_i]),
    "l3_l2_count": (_i64, [_i]),
    "l3_state_count": (_i64, [_i]),
    "l3_num_tensors": (_i, [_i]),
    "l3_tensor_info": (_i, [_i, _i, C.c_char_p, _i, C.POINTER(_i), C.POINTER(_i64), C.POINTER(_i), C.POINTER(_i64)]),
    "l3_frontend_shape": (_i, [_i, C.POINTER(_i), C.POINTER(_i)]),
    "l3_embedding_map_shape": (_i, [_i, C.POINTER(_i), C.POINTER(_i)]),
    "l3_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "l3_ctx_create": (_vp, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "l3_ctx_destroy": (None, [_vp]),
    "l3_ctx_set_use_tensor_cores": (_i, [_vp, _i]),
    "l3_ctx_uses_tensor_cores": (_i, [_vp]),
    "l3_ctx_set_fused_inference": (_i, [_vp, _i]),
    "l3_upload_batch_host": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i]),
    "l3_forward_backward": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i]),
    "l3_adam_step": (_i, [_vp, _f]),
    "l3_adam_set_t": (_i, [_vp, _i64]),
    "l3_adam_get_t": (_i64, [_vp]),
    "l3_get_metrics": (_i, [_vp, _fp]),
    "l3_train_step_staged": (_i, [_vp, _i, _f, _fp]),
    "l3_train_step_host": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _f, _fp]),
    "l3_dp_unique_id": (_i, [C.c_char_p]),
    "l3_dp_init": (_i, [_vp, C.c_char_p, _i, _i]),
    "l3_dp_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "l3_dp_nccl_version": (_i, []),
    "l3_dp_train_step_staged": (_i, [_vp, _i, _i, _f, _fp]),
    "l3_dp_average_bn_state": (_i, [_vp]),
    "l3_predict": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _vp]),
    "l3_embed_audio": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "l3_embed_audio_frames": (_i, [_vp, _vp, _i, _i64, _i, _i, _i, _vp]),
    "l3_embed_vision": (_i, [_vp, _vp, _i, _i, _vp]),
    "l3_frontend_fwd": (_i, [_vp, _vp, _i, _i, _vp]),
    "l3_conv3x3_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "l3_conv3x3_fwd_stats": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "l3_conv3x3_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "l3_conv3x3_dgrad_stats": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "l3_conv3x3_wgrad": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "l3_launch_count": (C.c_uint64, []),
    "l3_ctx_set_two_streams": (_i, [_vp, _i]),
    "l3_ctx_profile_enable": (_i, [_vp, _i]),
    "l3_ctx_profile_read": (_i, [_vp, _fp, C.POINTER(_i)]),
    "l3_act_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "l3_bn_act_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "l3_gmaxpool_fwd": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "l3_gmaxpool_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp]),
    "l3_debug_read": (_i64, [_vp, C.c_char_p, _i, _fp, _i64]),
}

_lib = None


def load():
    """Load the shared library and type every entry point.  Raises L3Error when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise L3Error("%s not found: run `python -m l3embedding_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().l3_last_error().decode("utf-8", "replace")


def check(rc, what: str):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise L3Error("%s failed: %s" % (what, last_error()))
    return rc


def model_id(model_type: str) -> int:
    if model_type not in MODEL_IDS:
        raise ValueError('Invalid model type: "{}"'.format(model_type))
    return MODEL_IDS[model_type]


def tensor_table(model_type: str):
    """[(name, arena, offset, shape)] in keras layer order (vision tower, audio tower, dense_1, dense_2)."""
    lib = load()
    m = model_id(model_type)
    out = []
    for i in range(check(lib.l3_num_tensors(m), "l3_num_tensors")):
        name = C.create_string_buffer(96)
        arena, off, nd = C.c_int(), C.c_int64(), C.c_int()
        dims = (C.c_int64 * 4)()
        check(lib.l3_tensor_info(m, i, name, 96, C.byref(arena), C.byref(off), C.byref(nd), dims), "l3_tensor_info")
        out.append((name.value.decode(), arena.value, off.value, tuple(int(dims[k]) for k in range(nd.value))))
    return out
