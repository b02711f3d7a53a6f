"""ctypes binding of ``libb200gcn.so`` (the C ABI declared in ``include/b200gcn.h``).

There is no CPU implementation behind this module: if the library cannot be loaded, or a call is made
with tensors that are not on a CUDA device, the call raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200gcn.so")

OK, ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_RANGE = 0, 1, 2, 3, 4
ABI_VERSION = 10


class EngineError(RuntimeError):
    """A CUDA runtime / launch failure reported by libb200gcn."""


class SpmmArgs(C.Structure):
    """Mirror of ``b200gcn_spmm_args`` (include/b200gcn.h)."""

    _fields_ = [
        ("n_rows", C.c_int64), ("dim", C.c_int32), ("flags", C.c_int32),
        ("rowptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
        ("x", C.c_void_p), ("x2", C.c_void_p), ("x_split", C.c_int64), ("ldx", C.c_int64),
        ("y", C.c_void_p), ("ldy", C.c_int64),
        ("noise", C.c_void_p), ("ldn", C.c_int64),
        ("eps", C.c_float), ("acc_scale", C.c_float), ("seed", C.c_uint64),
        ("acc_in", C.c_void_p), ("acc_in2", C.c_void_p), ("acc_split", C.c_int64), ("ld_acc_in", C.c_int64),
        ("acc_out", C.c_void_p), ("ld_acc_out", C.c_int64),
        ("y_peers", C.c_void_p), ("y_mc", C.c_void_p), ("y_peer_row0", C.c_int64), ("ld_peer", C.c_int64),
        ("n_peers", C.c_int32), ("n_acc_extra", C.c_int32),
        ("acc_extra", C.c_void_p * 3), ("ld_acc_extra", C.c_int64), ("peer_need", C.c_void_p),
    ]


CHAIN_MAX_PHASES, CHAIN_MAX_RANKS, CHAIN_SCRATCH_BYTES = 12, 16, 256


class ChainSync(C.Structure):
    """Mirror of ``b200gcn_chain_sync`` (include/b200gcn.h)."""

    _fields_ = [("n_ranks", C.c_int32), ("rank", C.c_int32), ("epoch", C.c_uint32), ("start_wait_phase", C.c_int32),
                ("flags", C.c_void_p), ("flags_peers", C.c_void_p), ("scratch", C.c_void_p),
                ("wait_phase", C.c_int8 * CHAIN_MAX_PHASES), ("wait_local", C.c_int8 * CHAIN_MAX_PHASES),
                ("merge_next", C.c_int8 * CHAIN_MAX_PHASES)]


class HubPlan(C.Structure):
    """Mirror of ``b200gcn_hub_plan`` (include/b200gcn.h)."""

    _fields_ = [("n_hubs", C.c_int32), ("n_chunks", C.c_int32), ("hub_rows", C.c_void_p),
                ("hub_chunk_ptr", C.c_void_p), ("chunk_beg", C.c_void_p), ("chunk_end", C.c_void_p),
                ("scratch", C.c_void_p)]


# name -> (restype, argtypes); the single source of truth for the symbol-export test
_P, _I64, _I32, _F, _SZP = C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.POINTER(C.c_size_t)
SIGNATURES = {
    "b200gcn_abi_version": (C.c_int, []),
    "b200gcn_last_error": (C.c_char_p, []),
    "b200gcn_device_info": (C.c_int, [C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I32), C.POINTER(_I32)]),
    "b200gcn_csr_from_coo_workspace": (C.c_int, [_I64, _I64, _I64, _SZP]),
    "b200gcn_csr_from_coo": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _P, _P, _P, _P, C.c_size_t, C.c_int, _P]),
    "b200gcn_csr_from_interactions_workspace": (C.c_int, [_I64, _I64, _I64, _SZP]),
    "b200gcn_csr_from_interactions": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _P, C.c_size_t, C.c_int, _P]),
    "b200gcn_gcn_norm_csr": (C.c_int, [_P, _P, _P, _P, _P, _I64, _P]),
    "b200gcn_bipartite_norm_coo": (C.c_int, [_P, _P, _I64, _I64, _I64, C.c_int, _P, _P, C.c_size_t, _P]),
    "b200gcn_csr_transpose_workspace": (C.c_int, [_I64, _I64, _I64, _SZP]),
    "b200gcn_csr_transpose": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P, _P, _P, _P, C.c_size_t, _P]),
    "b200gcn_csr_row_ids": (C.c_int, [_P, _I64, _I64, _P, _P]),
    "b200gcn_csr_mask_workspace": (C.c_int, [_I64, _I64, _SZP]),
    "b200gcn_csr_mask": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _P, _P, _P, C.POINTER(_I64), _P, C.c_size_t, _P]),
    "b200gcn_spmm": (C.c_int, [C.POINTER(SpmmArgs), _P]),
    "b200gcn_plan_hubs": (C.c_int, [_P, _I64, _I64, _P, _I32, C.POINTER(_I32), _P]),
    "b200gcn_spmm_planned": (C.c_int, [C.POINTER(SpmmArgs), _I64, _P, _I32, _P]),
    "b200gcn_spmm_hubs": (C.c_int, [C.POINTER(SpmmArgs), C.POINTER(HubPlan), _P]),
    "b200gcn_spmm_chain": (C.c_int, [C.POINTER(SpmmArgs), _I32, C.POINTER(ChainSync), _P]),
    "b200gcn_bpr_loss_workspace": (C.c_int, [_I64, _SZP]),
    "b200gcn_bpr_loss": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, _P, _I64, _I32, _F, _F, C.c_int,
                                   _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _P, C.c_size_t, _P]),
    "b200gcn_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I64, _P]),
    "b200gcn_bignn_tail_backward": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _P, _I64, _P, _F, _F, C.c_int, _P, _I64, _I64,
                                              _I32, _I32, _P, _P, _P, _P, _P]),
    "b200gcn_fullsort_topk_workspace": (C.c_int, [_I64, _I64, _I32, _SZP]),
    "b200gcn_fullsort_topk": (C.c_int, [_P, _I64, _I64, _P, _I64, _I64, _I32, _I32, _I64, _P, _P, _P, _P, _P,
                                        C.c_size_t, _P]),
    "b200gcn_fullsort_scores": (C.c_int, [_P, _I64, _I64, _P, _I64, _I64, _I32, _P, _I64, _P]),
    "b200gcn_inter_open": (C.c_int, [C.c_char_p, C.POINTER(_P), C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]),
    "b200gcn_inter_read": (C.c_int, [_P, _P, _P]),
    "b200gcn_inter_close": (None, [_P]),
    "b200gcn_bignn_tail": (C.c_int, [_P, _I64, _P, _I64, _P, _P, _P, _P, _I64, _I32, _I32, _F, _P, _F, C.c_int,
                                     _P, _I64, _P, _I64, _P, _I64, _P]),
}

_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m recbole_gnn_b200.build` "
                "(__graft_entry__.build()).  recbole_gnn_b200 has no CPU or PyTorch fallback.")
        import torch  # noqa: F401  (makes sure torch's libcudart.so.12 is the one already mapped)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.b200gcn_abi_version() != ABI_VERSION:
            raise ImportError("libb200gcn.so ABI version mismatch; rebuild it")
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = (load().b200gcn_last_error() or b"").decode("utf-8", "replace")
    if rc in (ERR_INVALID, ERR_WORKSPACE):
        raise ValueError(f"b200gcn: {msg}")
    if rc == ERR_RANGE:
        raise IndexError(f"b200gcn: {msg}")
    raise EngineError(f"b200gcn: {msg}")


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors, what: str = "input") -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                f"recbole_gnn_b200: {what} is on {t.device}; the engine only runs on CUDA devices "
                "(there is deliberately no CPU fallback).")
