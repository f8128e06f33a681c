"""ctypes binding of libdcb_b200.so (the C ABI of include/dcb.h).

There is NO CPU fallback: if the library is missing or a call fails, a DcbError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdcb_b200.so")

ENV_IDS = {"cube3": 0, "puzzle15": 1, "puzzle24": 2, "puzzle35": 3, "puzzle48": 4, "lightsout7": 5, "cube4": 6}


class DcbError(RuntimeError):
    pass


class SearchInst(ctypes.Structure):
    """dcb_search_inst == dcb_open_state (include/dcb.h): one problem instance's record, 128 bytes."""
    _fields_ = ([(n, c_uint32) for n in ("open_size", "n_popped", "n_expand", "min_key", "goal_id", "goal_key", "done", "n_goals",
                                         "next_slot", "base_slot", "iterations", "tile_off")] +
                [("nodes_generated", ctypes.c_uint64), ("nodes_expanded", ctypes.c_uint64)] +
                [(n, c_uint32) for n in ("thr_key", "thr_id", "need", "prefix", "cand_count", "n_holes", "n_surv", "take_all", "n_at_pop",
                                         "n_take", "resting", "overflow", "thr_lo")] + [("reserved", c_uint32 * 3)])

    @property
    def size(self) -> int:          # OPEN entries (the stand-alone queue's name for it)
        return self.open_size


OpenState = SearchInst
INST_WORDS = ctypes.sizeof(SearchInst) // 4


class StepPlan(ctypes.Structure):
    """dcb_step_plan (include/dcb.h), 64 bytes."""
    _fields_ = ([(n, c_uint32) for n in ("n_tiles", "n_parents", "n_kept", "n_ambiguous", "closed_entries", "n_running", "error", "budget")] +
                [("total_kept", ctypes.c_uint64), ("total_expanded", ctypes.c_uint64), ("reserved1", ctypes.c_uint64 * 2)])


PLAN_WORDS = ctypes.sizeof(StepPlan) // 4


class SearchCtx(ctypes.Structure):
    """dcb_search_ctx (include/dcb.h): geometry + device buffers of one engine."""
    _fields_ = ([(n, c_int32) for n in ("env", "n_inst", "batch", "semantics")] +
                [("slots_per_inst", c_uint32), ("open_per_inst", c_uint32), ("closed_capacity", c_int64)] +
                [(n, c_void_p) for n in ("d_arena", "d_node_g", "d_node_solved", "d_slot_parent", "d_closed", "d_open_key", "d_open_id", "d_open_key_lo", "d_inst",
                                         "d_plan", "d_weights", "d_popped_ids", "d_tiles", "d_hash", "d_kept_ids", "d_pop_scratch",
                                         "d_closed_scratch")])


_lib = None

_P = c_void_p  # every device/host buffer is passed as a raw address
_SIGS = {
    "dcb_abi_version": (c_int, []),
    "dcb_error_string": (c_char_p, [c_int]),
    "dcb_last_cuda_error": (c_char_p, []),
    "dcb_env_num_moves": (c_int, [c_int]),
    "dcb_env_state_bytes": (c_int, [c_int]),
    "dcb_env_slot_align": (c_int, [c_int]),
    "dcb_env_goal_state": (c_int, [c_int, _P]),
    "dcb_env_move_table": (c_int, [c_int, _P, c_int64]),
    "dcb_expand": (c_int, [c_int, _P, c_int64, _P, _P, _P, _P]),
    "dcb_expand_indexed": (c_int, [c_int, _P, _P, c_int64, _P, _P, _P, _P]),
    "dcb_next_state": (c_int, [c_int, _P, c_int64, c_int, _P, _P]),
    "dcb_is_solved": (c_int, [c_int, _P, c_int64, _P, _P]),
    "dcb_hash_states": (c_int, [c_int, _P, c_int64, _P, _P]),
    "dcb_nnet_input": (c_int, [c_int, _P, c_int64, _P, _P]),
    "dcb_expand_host": (c_int, [c_int, _P, c_int64, _P, _P, _P, c_int]),
    "dcb_next_state_host": (c_int, [c_int, _P, c_int64, c_int, _P, c_int]),
    "dcb_is_solved_host": (c_int, [c_int, _P, c_int64, _P, c_int]),
    "dcb_closed_bytes": (c_int64, [c_int64]),
    "dcb_closed_clear": (c_int, [_P, c_int64, _P]),
    "dcb_closed_scratch_bytes": (c_int64, [c_int64]),
    "dcb_closed_insert": (c_int, [c_int, _P, c_int64, _P, _P, _P, _P, c_uint32, c_int64, _P, _P, _P, _P]),
    "dcb_closed_rehash": (c_int, [_P, c_int64, _P, c_int64, _P]),
    "dcb_open_clear": (c_int, [_P, _P]),
    "dcb_open_push": (c_int, [_P, _P, _P, c_int64, _P, _P, c_uint32, _P, c_int64, _P]),
    "dcb_open_scratch_bytes": (c_int64, [c_int64, c_int64]),
    "dcb_open_pop": (c_int, [_P, _P, _P, c_int64, c_int32, c_int, _P, _P, _P, _P]),
    "dcb_child_meta": (c_int, [c_int, _P, c_int64, c_uint32, _P, _P, _P]),
    "dcb_compact_kept": (c_int, [_P, c_uint32, c_int64, _P, _P, _P]),
    "dcb_gather_nnet_input": (c_int, [c_int, _P, _P, c_int64, _P, _P]),
    "dcb_compute_cost": (c_int, [_P, _P, _P, _P, c_float, c_int64, _P, _P]),
    "dcb_reconstruct_path": (c_int, [c_int, _P, c_uint32, c_int32, _P, _P, _P]),
    "dcb_search_pop_scratch_bytes": (c_int64, [c_int32, c_int64, c_int32]),
    "dcb_search_reset": (c_int, [_P, _P, _P]),
    "dcb_search_pop": (c_int, [_P, c_int, _P]),
    "dcb_search_expand": (c_int, [_P, _P]),
    "dcb_search_closed": (c_int, [_P, _P]),
    "dcb_search_push": (c_int, [_P, _P, _P, c_int32, c_float, _P]),
    "dcb_search_path": (c_int, [_P, c_uint32, c_int32, _P, _P, _P]),
    "dcb_resnet_gemm": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, c_float, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, _P]),
    "dcb_resnet_gemm_ex": (c_int, [_P, _P, c_int64, _P, _P, c_int64, _P, c_float, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32,
                                   _P, c_int32, c_int32, _P, _P]),
    "dcb_resnet_gemm_scratch_bytes": (c_int64, []),
    "dcb_onehot_fp16": (c_int, [_P, c_int64, c_int32, c_int32, c_int32, _P, _P]),
    "dcb_onehot_fp16_nodes": (c_int, [c_int, _P, _P, c_int64, c_int32, c_int32, _P, _P]),
    "dcb_onehot_fp16_nodes_ex": (c_int, [c_int, _P, _P, c_int64, c_int32, c_int32, _P, _P, c_int32, _P]),
    "dcb_rowdot": (c_int, [_P, _P, _P, c_float, c_int64, c_int32, c_int32, _P, _P]),
}


def exported_symbols():
    """Every symbol include/dcb.h declares (used by the CPU-tier ABI test)."""
    return sorted(_SIGS)


def load():
    """Load the library once; raise DcbError (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DcbError("libdcb_b200.so is not built (%s). Run `python -m deepcubea_b200.build`; "
                       "there is no CPU fallback for this path." % LIB_PATH)
    try:
        import torch  # noqa: F401  (loads the CUDA runtime the library links against)
    except Exception:  # pragma: no cover
        pass
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise DcbError("cannot load %s: %s" % (LIB_PATH, e))
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.dcb_abi_version() != 1:
        raise DcbError("libdcb_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    lib = load()
    msg = lib.dcb_error_string(rc).decode()
    cuda = lib.dcb_last_cuda_error().decode()
    raise DcbError("%s failed: %s (%d)%s" % (what or "dcb call", msg, rc, (" -- CUDA: " + cuda) if cuda else ""))


def ptr(t) -> int:
    """Raw address of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
