"""ctypes binding of libtkbo.so (C ABI: include/tkbo.h).

This is the only place the package touches native code.  There is no CPU / PyTorch / Triton fallback on
the hot path: a missing library, a missing symbol or a non-sm_100 device raises ``TkboError`` loudly.
PyTorch is used by callers purely as the owner of device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# TKBO_LIB selects another build of the same library (tools/trace_k1.py uses libtkbo_trace.so, compiled with -DTKBO_TRACE)
LIB_PATH = os.environ.get("TKBO_LIB") or os.path.join(_HERE, "libtkbo.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "tkbo.h")

MPES_RANK, MPES_PES = 0, 1


class TkboError(RuntimeError):
    pass


c_p = ctypes.c_void_p
c_i, c_i64, c_d = ctypes.c_int, ctypes.c_int64, ctypes.c_double

_SIGNATURES = {
    "tkbo_version": (c_i, []),
    "tkbo_last_error": (ctypes.c_char_p, []),
    "tkbo_create": (c_i, [ctypes.POINTER(c_p), c_i]),
    "tkbo_destroy": (c_i, [c_p]),
    "tkbo_sm_count": (c_i, [c_p]),
    "tkbo_launch_count": (c_i64, [c_p]),
    "tkbo_debug_set_trace": (c_i, [c_p, c_p, c_i]),
    "tkbo_rff_basis_create": (c_i, [c_p, ctypes.POINTER(c_p), c_i, c_i, c_i]),
    "tkbo_rff_basis_create_ex": (c_i, [c_p, ctypes.POINTER(c_p), c_i, c_i, c_i, c_i]),
    "tkbo_rff_basis_destroy": (c_i, [c_p, c_p]),
    "tkbo_rff_basis_set": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_p]),
    "tkbo_rff_basis_config": (c_i, [c_p, c_p, c_i64, ctypes.POINTER(ctypes.c_int32)]),
    "tkbo_rff_argmax": (c_i, [c_p, c_p, c_p, c_i64, c_i64, c_p, c_p, c_p]),
    "tkbo_rff_argmax_key": (c_i, [c_p, c_p, c_p, c_i64, c_i64, c_p, c_p]),
    "tkbo_peer_buffer_bytes": (c_i64, [c_i, c_i]),
    "tkbo_rff_argmax_peer": (c_i, [c_p, c_p, c_p, c_i64, c_i64, ctypes.POINTER(c_p), c_i, c_i, ctypes.c_uint64, c_p, c_p, c_p]),
    "tkbo_peer_status": (c_i64, [c_p]),
    "tkbo_two_pass_items": (c_i, [c_p, c_p, c_p]),
    "tkbo_rff_argmax_host_peer": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_i64, c_i, c_i64, ctypes.POINTER(c_p), c_i, c_i,
                                        ctypes.c_uint64, c_p, c_p]),
    "tkbo_merge_peer_keys": (c_i, [c_p, c_p, c_i, c_i, ctypes.c_uint64, c_p, c_p, c_p]),
    "tkbo_peer_push_keys": (c_i, [c_p, c_p, c_i, ctypes.POINTER(c_p), c_i, c_i, ctypes.c_uint64, c_p]),
    "tkbo_rff_duel_copeland": (c_i, [c_p, c_p, c_p, c_i, c_p, c_p, c_p]),
    "tkbo_rff_duel_copeland_partial": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_p]),
    "tkbo_copeland_finish": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p, c_p]),
    "tkbo_unpack_argmax_key": (c_i, [c_p, c_p, c_i, c_p, c_p, c_p]),
    "tkbo_rff_eval": (c_i, [c_p, c_p, c_p, c_i64, c_p, c_i64, c_p]),
    "tkbo_merge_argmax": (c_i, [c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p]),
    "tkbo_rff_argmax_host": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p, c_i64, c_i, c_p, c_p]),
    "tkbo_rff_eval_batched": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "tkbo_rff_features_batched": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p]),
    "tkbo_rff_ascent_batched": (c_i, [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_d, c_d, c_i, c_d, c_d, c_d,
                                      ctypes.POINTER(c_i), c_p]),
    "tkbo_mpes_topk": (c_i, [c_p, c_p, c_p, c_i, c_i64, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_p]),
    "tkbo_mpes_indiff": (c_i, [c_p, c_p, c_p, c_i, c_i64, c_i, c_i, c_d, c_p, c_p]),
    "tkbo_argmax_rows": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p]),
    "tkbo_soft_copeland": (c_i, [c_p, c_p, c_i, c_p, c_p, c_p]),
}

_lib = None
_lib_lock = threading.Lock()


def declared_symbols():
    """Function names declared in include/tkbo.h (used by the ABI test)."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tkbo_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load libtkbo.so once; raises TkboError if it is missing or lacks a declared symbol."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise TkboError(
                f"{LIB_PATH} not found. Build it with `python __graft_entry__.py build` "
                "(needs nvcc; sm_100a only). There is no CPU fallback for this path.")
        try:
            handle = ctypes.CDLL(LIB_PATH)
        except OSError as err:
            raise TkboError(f"cannot load {LIB_PATH}: {err}") from err
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as err:
                raise TkboError(f"{LIB_PATH} does not export {name}") from err
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        return _lib


def last_error():
    return lib().tkbo_last_error().decode("utf-8", "replace")


def check(rc, what):
    if rc != 0:
        raise TkboError(f"{what} failed (status {rc}): {last_error()}")


_handles = {}
_handles_lock = threading.Lock()


def handle(device_index):
    """One tkbo handle per CUDA device, created on first use."""
    with _handles_lock:
        h = _handles.get(device_index)
        if h is None:
            out = c_p()
            check(lib().tkbo_create(ctypes.byref(out), int(device_index)), "tkbo_create")
            h = out
            _handles[device_index] = h
        return h


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise TkboError("no CUDA device visible: this package runs its hot path only on B200 (sm_100a) "
                        "through libtkbo.so and has no CPU fallback")
    return torch


def current_stream_ptr(device):
    import torch
    return c_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a contiguous torch tensor (or None)."""
    if t is None:
        return c_p(0)
    assert t.is_contiguous(), "tkbo: tensor must be contiguous"
    return c_p(t.data_ptr())
