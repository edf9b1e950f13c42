"""ctypes binding of libkanpyo_b200.so (include/kanpyo_b200.h).  No CPU fallback: if the library
cannot be built or loaded, importing a compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_lib = None


class KanpyoB200Error(RuntimeError):
    def __init__(self, status: int, detail: str):
        self.status = status
        super().__init__("kanpyo_b200 status %d: %s" % (status, detail))


KP_OK, KP_ERR_ARG, KP_ERR_CUDA, KP_ERR_DICT, KP_ERR_UTF8, KP_ERR_NOMEM, KP_ERR_TOO_LARGE, KP_ERR_BLOB = (
    0, -1, -2, -3, -4, -5, -6, -7)


class DictArrays(C.Structure):
    _fields_ = [("da", C.c_void_p), ("da_len", C.c_uint64),
                ("dup_ids", C.c_void_p), ("dup_counts", C.c_void_p), ("n_dup", C.c_uint64),
                ("morphs", C.c_void_p), ("n_morphs", C.c_uint64),
                ("conn_row", C.c_uint64), ("conn_col", C.c_uint64), ("conn", C.c_void_p),
                ("char_category", C.c_void_p), ("n_char_category", C.c_uint64),
                ("invoke_list", C.c_void_p), ("n_invoke", C.c_uint64),
                ("group_list", C.c_void_p), ("n_group", C.c_uint64),
                ("unk_cat", C.c_void_p), ("unk_first_id", C.c_void_p), ("unk_count", C.c_void_p),
                ("n_unk_map", C.c_uint64),
                ("unk_morphs", C.c_void_p), ("n_unk_morphs", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("n_sent", C.c_uint64), ("n_tokens", C.c_uint64), ("tok_off", C.c_void_p),
                ("tokens", C.c_void_p), ("eos_cost", C.c_void_p)]


class Result8(C.Structure):
    _fields_ = [("n_sent", C.c_uint64), ("n_tokens", C.c_uint64), ("tok_off", C.c_void_p),
                ("tokens", C.c_void_p), ("eos_cost", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("bytes", "chars", "nodes", "tokens", "sentences", "probes", "probes_ok",
                                          "pairs")]


class Profile(C.Structure):
    _fields_ = [(k, C.c_float) for k in ("h2d_ms", "prep_ms", "lattice_ms", "bucket_ms", "viterbi_ms", "backtrace_ms",
                                         "d2h_ms", "total_ms")] + [("kernel_launches", C.c_uint32),
                                                                   ("chunks", C.c_uint32), ("fused_ms", C.c_float),
                                                                   ("fused_sentences", C.c_uint32)]


class Lattice(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("nodes", C.c_void_p)]


# every symbol include/kanpyo_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "kp_abi_version": (C.c_int, []),
    "kp_strerror": (C.c_char_p, [C.c_int]),
    "kp_last_error": (C.c_char_p, []),
    "kp_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "kp_dict_create": (C.c_int, [C.POINTER(DictArrays), C.c_int, C.POINTER(_P)]),
    "kp_dict_pack": (C.c_int, [C.POINTER(DictArrays), _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "kp_dict_blob": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "kp_dict_device_blob": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "kp_dict_create_from_blob": (C.c_int, [_P, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "kp_dict_create_from_device_blob": (C.c_int, [_P, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "kp_dict_destroy": (None, [_P]),
    "kp_tokenizer_create": (C.c_int, [_P, C.POINTER(_P)]),
    "kp_tokenizer_destroy": (None, [_P]),
    "kp_tokenizer_set_chunk_bytes": (C.c_int, [_P, C.c_uint64]),
    "kp_tokenizer_set_count_work": (C.c_int, [_P, C.c_int]),
    "kp_tokenize": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(Result)]),
    "kp_tokenize_batch": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(Result)]),
    "kp_tokenize_batch_device": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Result)]),
    "kp_tokenize_batch8": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(Result8)]),
    "kp_tokenize_batch_device8": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(Result8)]),
    "kp_expand_tokens8": (C.c_int, [C.POINTER(Result8), _P, _P]),
    "kp_tokenizer_set_path": (C.c_int, [_P, C.c_int]),
    "kp_queue_create": (C.c_int, [_P, C.c_uint32, C.POINTER(_P)]),
    "kp_queue_set_path": (C.c_int, [_P, C.c_int]),
    "kp_queue_set_blocking_sync": (C.c_int, [_P, C.c_int]),
    "kp_tokenizer_set_blocking_sync": (C.c_int, [_P, C.c_int]),
    "kp_queue_submit": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "kp_queue_wait": (C.c_int, [_P, C.c_uint64, C.POINTER(Result8)]),
    "kp_queue_destroy": (None, [_P]),
    "kp_shards_create": (C.c_int, [C.POINTER(DictArrays), C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "kp_shards_tokenize": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(Result8)]),
    "kp_shards_tokenize_gather": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(Result8)]),
    "kp_shards_times": (C.c_int, [_P, C.POINTER(C.c_float * 4)]),
    "kp_shards_copy_to_host": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "kp_shards_destroy": (None, [_P]),
    "kp_gather_unique_id": (C.c_int, [_P]),
    "kp_gather_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.c_uint64, C.c_uint64, C.POINTER(_P)]),
    "kp_gather_tokens": (C.c_int, [_P, C.POINTER(Result8), C.POINTER(Result8)]),
    "kp_gather_last_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "kp_gather_copy_to_host": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "kp_gather_destroy": (None, [_P]),
    "kp_last_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "kp_last_profile": (C.c_int, [_P, C.POINTER(Profile)]),
    "kp_tokenizer_sync": (C.c_int, [_P]),
    "kp_copy_to_host": (C.c_int, [_P, _P, _P, C.c_uint64]),
    "kp_lattice_dump": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(Lattice)]),
    "kp_da_build": (C.c_int, [_P, _P, C.c_uint64, _P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "kp_da_free": (None, [_P]),
    "kp_da_common_prefix": (C.c_int, [_P, _P, C.c_uint64, C.c_int, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
}


def lib_path() -> str:
    return _build.LIB


def load():
    """Build (if stale) and load the C-ABI library.  Raises if it cannot be produced."""
    global _lib
    if _lib is None:
        path = os.environ.get("KANPYO_B200_LIB") or _build.LIB
        if path == _build.LIB and _build.stale():
            try:
                _build.build()
            except Exception:
                if not os.path.exists(path):
                    raise
        L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)     # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int):
    if status != 0:
        L = load()
        detail = L.kp_last_error().decode("utf-8", "replace") or L.kp_strerror(status).decode()
        raise KanpyoB200Error(status, detail)
