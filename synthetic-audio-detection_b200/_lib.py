"""ctypes binding of libsad_b200.so (include/sad_b200.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAD_LIB") or os.path.join(_HERE, "libsad_b200.so")   # SAD_LIB: A/B a differently built library
LIB_PATHS = {"bf16": LIB_PATH, "fp16": os.path.join(_HERE, "libsad_b200_f16.so")}   # one build per activation dtype

SAD_OK, SAD_EINVAL, SAD_ENODEVICE, SAD_ECUDA, SAD_ESTATE = 0, -1, -2, -3, -4
_CODES = {SAD_EINVAL: "SAD_EINVAL", SAD_ENODEVICE: "SAD_ENODEVICE", SAD_ECUDA: "SAD_ECUDA", SAD_ESTATE: "SAD_ESTATE"}

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> (restype, argtypes): every symbol include/sad_b200.h declares
SIGNATURES = {
    "sad_create": (_i, [C.POINTER(_vp), _i, _i, _i]),
    "sad_create_ex": (_i, [C.POINTER(_vp), _i, _i, _i, C.c_char_p]),
    "sad_backbone": (C.c_char_p, [_vp]),
    "sad_destroy": (_i, [_vp]),
    "sad_last_error": (C.c_char_p, [_vp]),
    "sad_version": (C.c_char_p, []),
    "sad_act_dtype": (C.c_char_p, []),
    "sad_weight_count": (_i, []),
    "sad_weight_name": (C.c_char_p, [_i]),
    "sad_weight_numel": (_ll, [_i]),
    "sad_backbone_weight_count": (_i, [C.c_char_p]),
    "sad_backbone_weight_name": (C.c_char_p, [C.c_char_p, _i]),
    "sad_backbone_weight_numel": (_ll, [C.c_char_p, _i]),
    "sad_load_weights": (_i, [_vp, _i, C.POINTER(_vp), _i]),
    "sad_set_frontend_constants": (_i, [_vp, _vp, _vp]),
    "sad_frontend_logmel": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "sad_frontend_image": (_i, [_vp, _vp, _i, _vp, _vp]),
    "sad_slice_count": (_ll, [_ll, _ll, _ll]),
    "sad_slice_gate": (_i, [_vp, _vp, _ll, _ll, _ll, _f, _vp, _vp]),
    "sad_gather_windows": (_i, [_vp, _vp, _vp, _i, _ll, _vp, _vp]),
    "sad_forward": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp, _vp]),
    "sad_forward_images": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp, _vp]),
    "sad_forward_host": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp]),
    "sad_clip_reduce": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "sad_n_heads": (_i, [_vp]),
    "sad_max_batch": (_i, [_vp]),
    "sad_launch_count": (_ll, [_vp]),
    "sad_profile_enable": (_i, [_vp, _i]),
    "sad_profile_read": (_i, [_vp, C.POINTER(C.c_double), C.POINTER(_ll)]),
    "sad_synth_segments": (_i, [_vp, _vp, _ll, _i, C.c_ulonglong, _vp]),
    "sad_ingest_length": (_ll, [_ll, _i]),
    "sad_ingest": (_i, [_vp, _vp, _i, _ll, _i, _i, _vp, _vp]),
    "sad_debug_conv": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "sad_debug_block": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp]),
    "sad_debug_stem": (_i, [_vp, _vp, _i, _vp, _vp]),
    "sad_debug_read": (_ll, [_vp, _i, _vp, _ll, _vp]),
}

_libs = {}


class SadError(RuntimeError):
    pass


def load(dtype: str = "bf16") -> C.CDLL:
    """dlopen the library built for activation dtype `dtype` ("bf16" default, "fp16") -- no GPU needed for this -- and
    declare the prototypes."""
    if dtype in _libs:
        return _libs[dtype]
    if dtype not in LIB_PATHS:
        raise ValueError(f"activation dtype {dtype!r}: choose 'bf16' or 'fp16'")
    path = LIB_PATHS[dtype]
    if not os.path.exists(path):
        raise SadError(
            f"{path} is missing: build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
            "There is no CPU or PyTorch fallback for this path.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.sad_act_dtype().decode() != dtype and not os.environ.get("SAD_LIB"):
        raise SadError(f"{path} was built for {lib.sad_act_dtype().decode()} activations, expected {dtype}")
    _libs[dtype] = lib
    return lib


def check(ctx, code: int, what: str):
    if code >= 0:
        return code
    msg = next(iter(_libs.values())).sad_last_error(ctx) if _libs else b""   # ctx-only accessor: any loaded build serves
    raise SadError(f"{what}: {_CODES.get(code, code)}: {msg.decode() if msg else ''}")
