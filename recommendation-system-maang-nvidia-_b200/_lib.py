"""ctypes binding of librecsys_b200.so (include/recsys_b200.h) with DLPack tensor exchange.

There is no CPU fallback: if the shared library cannot be built/loaded, or no sm_100 device is usable,
the first call raises.  Tensors are handed to C as borrowed `DLTensor*` taken from the objects' own
`__dlpack__()` capsules (torch tensors and numpy arrays both provide one); the capsule is kept alive
for the duration of the call and then dropped, which runs the producer's deleter.
"""

import ctypes
import os
from pathlib import Path

import numpy as np
import torch

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "librecsys_b200.so"

RS_ACT_NONE, RS_ACT_RELU, RS_ACT_SIGMOID = 0, 1, 2
RS_EXACT_FP32, RS_TENSOR_BF16, RS_AUTO = 0, 1, 2
_ACT = {None: RS_ACT_NONE, "linear": RS_ACT_NONE, "relu": RS_ACT_RELU, "sigmoid": RS_ACT_SIGMOID}
_MODE = {"exact": RS_EXACT_FP32, "tensor": RS_TENSOR_BF16, "auto": RS_AUTO}


class DLDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32)]


class DLDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("device", DLDevice), ("ndim", ctypes.c_int32),
                ("dtype", DLDataType), ("shape", ctypes.POINTER(ctypes.c_int64)),
                ("strides", ctypes.POINTER(ctypes.c_int64)), ("byte_offset", ctypes.c_uint64)]


_DLP = ctypes.POINTER(DLTensor)
_capsule_get = ctypes.pythonapi.PyCapsule_GetPointer
_capsule_get.restype = _DLP                    # DLManagedTensor starts with its DLTensor
_capsule_get.argtypes = [ctypes.py_object, ctypes.c_char_p]
# torch's C++ exporter: the same capsule tensor.__dlpack__() returns, without the Python-level stream handling
# (we pass stream=-1 there anyway: the library is stream-ordered behind the caller's current stream)
_torch_to_dlpack = getattr(torch._C, "_to_dlpack", None)

class X3Layer(ctypes.Structure):
    """rs_x3_layer_t of include/recsys_b200.h"""
    _fields_ = [(n, _DLP) for n in ("x", "ids", "x_hi", "x_lo", "wt_hi", "wt_lo", "bias", "y", "y_hi", "y_lo")]


# rs_allgather_fn of include/recsys_b200.h: (ctx, send, recv, bytes per rank) -> 0 on success; host memory
ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64)

_lib = None


def _needs_build() -> bool:
    if not _LIB_PATH.exists():
        return True
    t = _LIB_PATH.stat().st_mtime
    srcs = list((_HERE / "csrc").glob("*.cu")) + list((_HERE / "csrc").glob("*.cuh"))
    srcs.append(_HERE.parent / "include" / "recsys_b200.h")
    return any(s.exists() and s.stat().st_mtime > t for s in srcs)


def load():
    """Load (building first if the in-tree .so is missing or older than its sources)."""
    global _lib
    if _lib is not None:
        return _lib
    if _needs_build():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if os.path.exists(nvcc):
            from .build import build
            build()
        elif not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} is missing and nvcc is not available to build it; "
                               "run recommendation-system-maang-nvidia-_b200/build.py where nvcc exists")
    lib = ctypes.CDLL(str(_LIB_PATH))
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.rs_version.restype = i32
    lib.rs_last_error.restype = ctypes.c_char_p
    lib.rs_launch_count.restype = i64
    sigs = {
        "rs_device_sm_count": [ctypes.POINTER(i32)],
        "rs_vocab_create": [vp, i64, i64, ctypes.POINTER(vp)],
        "rs_vocab_free": [vp],
        "rs_vocab_size": [vp, ctypes.POINTER(i64)],
        "rs_vocab_lookup": [vp, _DLP, _DLP, vp],
        "rs_gather_concat": [i32, ctypes.POINTER(_DLP), ctypes.POINTER(_DLP), _DLP, vp],
        "rs_a2a_count": [_DLP, i32, _DLP, vp],
        "rs_a2a_scatter": [_DLP, i32, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_hash_table_fill": [_DLP, i64, i64, ctypes.c_uint64, vp],
        "rs_dense_forward": [_DLP, _DLP, _DLP, i32, _DLP, vp],
        "rs_dense_forward_bf16": [_DLP, _DLP, _DLP, i32, _DLP, _DLP, vp],
        "rs_gather_bf16": [_DLP, _DLP, _DLP, vp],
        "rs_dense_x3_prepare": [_DLP, _DLP, _DLP, vp],
        "rs_dense_x3_supported": [i64, i64, i32],
        "rs_dense_forward_x3": [i32, ctypes.POINTER(X3Layer), i32, vp],
        "rs_ranker_create": [_DLP, _DLP, ctypes.POINTER(_DLP), ctypes.POINTER(_DLP), ctypes.POINTER(_DLP),
                             ctypes.POINTER(_DLP), _DLP, _DLP, _DLP, _DLP, ctypes.POINTER(vp)],
        "rs_ranker_free": [vp],
        "rs_ranker_forward": [vp, _DLP, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_ranker_forward_rows": [vp, _DLP, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_chain_forward_bf16": [i32, i32, ctypes.POINTER(_DLP), ctypes.POINTER(_DLP), ctypes.POINTER(i32),
                                  ctypes.POINTER(_DLP), ctypes.POINTER(_DLP), ctypes.POINTER(_DLP), ctypes.POINTER(_DLP),
                                  ctypes.POINTER(_DLP), vp],
        "rs_dcn_cross_forward": [_DLP, _DLP, _DLP, _DLP, vp],
        "rs_heads_forward": [_DLP, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_cross_v2_combine": [_DLP, _DLP, _DLP, _DLP, vp],
        "rs_cross_heads_forward": [_DLP, _DLP, _DLP, _DLP, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_normalize_l2": [_DLP, vp],
        "rs_index_create": [i32, ctypes.POINTER(vp)],
        "rs_index_free": [vp],
        "rs_index_add": [vp, _DLP, vp],
        "rs_index_ntotal": [vp, ctypes.POINTER(i64)],
        "rs_index_dim": [vp, ctypes.POINTER(i32)],
        "rs_index_reconstruct": [vp, i64, _DLP, vp],
        "rs_index_search": [vp, _DLP, i32, i64, i32, _DLP, _DLP, vp],
        "rs_index_search_deferred": [vp, _DLP, i32, i64, i32, _DLP, _DLP, _DLP, vp],
        "rs_tensor_plan": [i32, i64, i64, i32, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32)],
        "rs_tensor_threshold_rank": [i32, ctypes.c_double, ctypes.POINTER(ctypes.c_int32)],
        "rs_index_shared_begin": [vp, _DLP, i32, _DLP, ctypes.POINTER(ctypes.c_void_p), vp],
        "rs_index_shared_finish": [vp, _DLP, i32, i64, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_tc_session_free": [vp],
        "rs_topk_verify": [_DLP, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_index_fallbacks": [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)],
        "rs_debug_tensor_scores": [vp, _DLP, _DLP, vp],
        "rs_matvec": [_DLP, _DLP, _DLP, vp],
        "rs_topk_merge": [_DLP, _DLP, _DLP, _DLP, vp],
        "rs_peer_create": [i32, i32, i64, ctypes.POINTER(vp), vp],
        "rs_peer_connect": [vp, vp],
        "rs_peer_local": [vp, ctypes.POINTER(vp), ctypes.POINTER(i64)],
        "rs_peer_publish": [vp, _DLP, vp],
        "rs_peer_wait": [vp, vp],
        "rs_peer_ack": [vp, vp],
        "rs_peer_free": [vp],
        "rs_peer_disconnect": [vp],
        "rs_a2a_plan_peer": [_DLP, i32, i32, _DLP, _DLP, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_peer_a2a_ids": [vp, _DLP, _DLP, vp],
        "rs_peer_a2a_rows": [vp, vp, _DLP, vp],
        "rs_peer_a2a_restore": [vp, _DLP, _DLP, _DLP, _DLP, _DLP, vp],
        "rs_comm_init": [i32, i32, ALLGATHER_FN, vp, ctypes.POINTER(vp)],
        "rs_comm_rank": [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)],
        "rs_comm_free": [vp],
        "rs_index_search_sharded": [vp, vp, i64, _DLP, i32, i32, _DLP, _DLP, _DLP, vp],
        "rs_sharded_table_create": [vp, _DLP, i64, i32, ctypes.POINTER(vp)],
        "rs_sharded_table_free": [vp],
        "rs_embedding_a2a_lookup": [vp, _DLP, _DLP, vp],
        "rs_sharded_table_last_counts": [vp, ctypes.POINTER(i64), vp],
        "rs_prof_enable": [i32],
        "rs_prof_read": [i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)],
    }
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = argtypes
    _lib = lib
    return lib


EXPORTED = ["rs_version", "rs_last_error", "rs_launch_count", "rs_device_sm_count", "rs_vocab_create", "rs_vocab_free", "rs_vocab_size", "rs_vocab_lookup", "rs_gather_concat", "rs_a2a_count", "rs_a2a_scatter", "rs_hash_table_fill",
            "rs_dense_forward", "rs_dense_forward_bf16", "rs_dense_x3_prepare", "rs_dense_x3_supported", "rs_dense_forward_x3", "rs_gather_bf16", "rs_chain_forward_bf16", "rs_ranker_create", "rs_ranker_free", "rs_ranker_forward", "rs_ranker_forward_rows", "rs_dcn_cross_forward", "rs_heads_forward", "rs_cross_v2_combine", "rs_cross_heads_forward", "rs_normalize_l2", "rs_index_create",
            "rs_index_free", "rs_index_add", "rs_index_ntotal", "rs_index_dim", "rs_index_reconstruct", "rs_index_search", "rs_index_search_deferred",
            "rs_tensor_plan", "rs_tensor_threshold_rank", "rs_index_shared_begin", "rs_index_shared_finish", "rs_tc_session_free", "rs_topk_verify",
            "rs_index_fallbacks", "rs_debug_tensor_scores",
            "rs_matvec", "rs_topk_merge", "rs_prof_enable", "rs_prof_read",
            "rs_peer_create", "rs_peer_connect", "rs_peer_local", "rs_peer_publish", "rs_peer_wait", "rs_peer_ack", "rs_peer_free", "rs_peer_disconnect",
            "rs_a2a_plan_peer", "rs_peer_a2a_ids", "rs_peer_a2a_rows", "rs_peer_a2a_restore",
            "rs_comm_init", "rs_comm_rank", "rs_comm_free", "rs_index_search_sharded", "rs_sharded_table_create", "rs_sharded_table_free",
            "rs_embedding_a2a_lookup", "rs_sharded_table_last_counts"]

PROF_SCAN, PROF_MERGE, PROF_FILTER, PROF_RESCORE, PROF_RANK, PROF_SAMPLE, PROF_DEEP, PROF_THRESHOLD, PROF_EXCHANGE, PROF_TOPS = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9


def check(status: int) -> None:
    if status != 0:
        msg = load().rs_last_error()
        raise RuntimeError("librecsys_b200: " + (msg.decode() if msg else f"status {status}"))


class _Borrowed:
    """A DLTensor* borrowed from obj.__dlpack__(); keeps the capsule (and obj) alive."""

    __slots__ = ("obj", "capsule", "ptr")

    def __init__(self, obj):
        self.obj = obj
        if isinstance(obj, torch.Tensor):
            if _torch_to_dlpack is not None:
                self.capsule = _torch_to_dlpack(obj)
            else:
                self.capsule = obj.__dlpack__(stream=-1) if obj.is_cuda else obj.__dlpack__()
        elif isinstance(obj, np.ndarray):
            if not obj.flags.writeable:      # DLPack (unversioned) cannot flag read-only: export a private copy
                obj = self.obj = obj.copy()
            self.capsule = obj.__dlpack__()
        else:
            raise TypeError(f"cannot export {type(obj)} through DLPack")
        self.ptr = _capsule_get(self.capsule, b"dltensor")


def dl(obj):
    """None -> NULL pointer; tensor/array -> borrowed DLTensor*."""
    if obj is None:
        return None
    return _Borrowed(obj)


def ptr(b):
    return b.ptr if b is not None else None


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def sm_count() -> int:
    out = ctypes.c_int(0)
    check(load().rs_device_sm_count(ctypes.byref(out)))
    return out.value


def launch_count() -> int:
    return int(load().rs_launch_count())


def prof_enable(on: bool) -> None:
    check(load().rs_prof_enable(1 if on else 0))


def prof_read(slot: int):
    """-> (total_ms, launches) of the event-bracketed launches of one kernel family since the last read."""
    ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
    check(load().rs_prof_read(slot, ctypes.byref(ms), ctypes.byref(n)))
    return ms.value, n.value


def act_code(activation) -> int:
    if activation not in _ACT:
        raise ValueError(f"unsupported activation {activation!r}")
    return _ACT[activation]


def mode_code(mode) -> int:
    if mode not in _MODE:
        raise ValueError(f"mode must be one of {sorted(_MODE)}")
    return _MODE[mode]
