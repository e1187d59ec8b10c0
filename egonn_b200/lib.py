"""ctypes binding of libegonn_b200.so (include/egonn_b200.h).  Fails loudly when the library is missing:
there is no CPU or PyTorch fallback for any operator of this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libegonn_b200.so")

EGN_MAX_LEVELS = 8
EGN_PYR_LEVELS = 10
EGN_MAX_HEAD_LEVELS = 4
EGN_MAX_EXTRA_BLOCKS = 3


class EgnError(RuntimeError):
    pass


class CoordsInfo(C.Structure):
    _fields_ = [("n_batches", C.c_int32), ("n_input", C.c_int32), ("n_rows", C.c_int32 * EGN_PYR_LEVELS),
                ("status", C.c_int32)]


class Layer(C.Structure):
    _fields_ = [("cin", C.c_int32), ("cout", C.c_int32), ("w", C.c_int64), ("scale", C.c_int64), ("shift", C.c_int64),
                ("wtc", C.c_int64)]


class Head(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("levels", C.c_int32 * EGN_MAX_HEAD_LEVELS), ("out_channels", C.c_int32),
                ("conv1x1", Layer * EGN_MAX_LEVELS), ("tconv", Layer * EGN_MAX_LEVELS)]


class Net(C.Structure):
    _fields_ = [("n_levels", C.c_int32), ("conv0_ksize", C.c_int32), ("conv0", Layer),
                ("down", Layer * EGN_MAX_LEVELS), ("conv1", Layer * EGN_MAX_LEVELS), ("conv2", Layer * EGN_MAX_LEVELS),
                ("res", Layer * EGN_MAX_LEVELS), ("eca_k", C.c_int32 * EGN_MAX_LEVELS), ("eca_w", C.c_int64 * EGN_MAX_LEVELS),
                ("global_head", Head), ("local_head", Head), ("global_mlp", Layer * 2),
                ("pool_method", C.c_int32), ("gem_p", C.c_float), ("gem_eps", C.c_float),
                ("desc_mlp", Layer * 2), ("kp_mlp", Layer * 2), ("sigma_mlp", Layer * 2),
                ("polar", C.c_int32), ("quant_step", C.c_float * 3), ("ignore_keypoint_regressor", C.c_int32),
                ("kpsig_mlp", Layer * 2),
                ("n_extra", C.c_int32 * EGN_MAX_LEVELS), ("xconv1", (Layer * EGN_MAX_EXTRA_BLOCKS) * EGN_MAX_LEVELS),
                ("xconv2", (Layer * EGN_MAX_EXTRA_BLOCKS) * EGN_MAX_LEVELS), ("xeca_k", (C.c_int32 * EGN_MAX_EXTRA_BLOCKS) * EGN_MAX_LEVELS),
                ("xeca_w", (C.c_int64 * EGN_MAX_EXTRA_BLOCKS) * EGN_MAX_LEVELS)]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_int64), ("ms", C.c_double), ("alg_bytes", C.c_double),
                ("flops", C.c_double)]


_P = C.c_void_p
_SIGNATURES = {
    "egn_last_error": (C.c_char_p, []),
    "egn_version": (C.c_int, []),
    "egn_ctx_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "egn_ctx_destroy": (C.c_int, [_P]),
    "egn_quantize": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_float), C.c_int, _P, _P, C.POINTER(C.c_int64), _P]),
    "egn_coords_build": (C.c_int, [_P, _P, C.c_int64, C.POINTER(CoordsInfo), _P]),
    "egn_coords_build_points": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(CoordsInfo), _P]),
    "egn_coords_get": (C.c_int, [_P, C.c_int, _P, _P]),
    "egn_coords_input_rows": (C.c_int, [_P, _P, _P]),
    "egn_coords_batch_offsets": (C.c_int, [_P, C.c_int, _P, _P]),
    "egn_coords_neighbors": (C.c_int, [_P, C.c_int, _P, _P]),
    "egn_weights_resident": (C.c_int, [_P, _P, C.c_size_t]),
    "egn_forward": (C.c_int, [_P, C.POINTER(Net), _P, _P, _P, _P, _P, _P, _P]),
    "egn_forward_tap": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "egn_conv": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "egn_conv_tc": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, C.c_int, _P, _P]),
    "egn_set_tensor_cores": (C.c_int, [_P, C.c_int]),
    "egn_global_pool": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "egn_broadcast_mul": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "egn_knn_l2": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "egn_match_mutual": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "egn_filter_points": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, _P, C.POINTER(C.c_int64), _P]),
    "egn_profile_enable": (C.c_int, [_P, C.c_int]),
    "egn_profile_read": (C.c_int, [_P, C.POINTER(ProfileEntry), C.c_int, C.POINTER(C.c_int), C.c_int]),
    "egn_launch_count": (C.c_int64, [_P]),
    "egn_debug_trace": (C.c_int, [_P, _P]),
    "egn_topk_smallest": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "egn_pack_topk": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, _P]),
    "egn_comm_unique_id": (C.c_int, [_P]),
    "egn_comm_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, _P]),
    "egn_comm_destroy": (C.c_int, [_P]),
    "egn_allgather_global": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library (built in-tree by egonn_b200/build.py or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EgnError(f"{LIB_PATH} is missing: run `python -m egonn_b200.build` (nvcc, sm_100a). "
                           "egonn_b200 has no CPU / PyTorch fallback.")
        if "EGN_NCCL_LIB" not in os.environ:       # egn_comm_* dlopen NCCL at first use: point them at PyTorch's bundled copy
            try:
                import nvidia.nccl as _nccl
                cand = os.path.join(list(_nccl.__path__)[0], "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["EGN_NCCL_LIB"] = cand
            except Exception:
                pass
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(status: int):
    if status != 0:
        msg = load().egn_last_error().decode("utf-8", "replace")
        raise EgnError(f"egonn_b200 error {status}: {msg}")
