"""The raw C ABI of libblance_b200.so in ctypes (struct blance_plan_in / blance_plan_out of
include/blance_b200.h, the exported symbols, and the lazily loaded library handle).  Importing this module loads
NO native code - the table builders (synth.py, tables.py) and bench.py's CPU reference arm use it without
mapping the CUDA library; capi() loads the library on first use."""
import ctypes
import os

from . import build as _build


class BlanceError(RuntimeError):
    """A negative blance_status from the C ABI (the message is blance_last_error())."""


_I32_FIELDS = ("n_nodes", "n_node_ids", "n_states", "n_parts", "n_slots", "max_iters", "top_state", "booster_kind",
               "add_is_nil", "has_part_weights", "has_node_weights", "has_hier_rules")
_PTR_FIELDS = ("state_priority", "state_constraints", "state_slot_off", "state_stickiness", "state_has_stickiness",
               "node_removed", "node_added", "node_weight", "node_has_weight", "part_in_prev", "part_in_assign",
               "part_weight", "part_has_weight", "part_name_rank", "prev_rows", "prev_shape", "cur_rows", "cur_shape",
               "extra_tot_first", "extra_tot_rest")


class _PlanIn(ctypes.Structure):        # struct blance_plan_in, include/blance_b200.h
    _fields_ = ([(n, ctypes.c_int32) for n in _I32_FIELDS] + [(n, ctypes.c_void_p) for n in _PTR_FIELDS] +
                [("n_rules", ctypes.c_int32), ("n_hier_bits", ctypes.c_int32), ("rule_off", ctypes.c_void_p),
                 ("ie_mask", ctypes.c_void_p), ("engine", ctypes.c_int32)])


class _PlanOut(ctypes.Structure):
    _fields_ = [("next_rows", ctypes.c_void_p), ("next_shape", ctypes.c_void_p), ("warn", ctypes.c_void_p),
                ("iters_run", ctypes.c_int32), ("converged", ctypes.c_int32), ("steps", ctypes.c_int64),
                ("device_ms", ctypes.c_float), ("kernel_ms", ctypes.c_float), ("pass_ms", ctypes.c_float),
                ("sticky_steps", ctypes.c_int64)]


_CAPI = None
EXPORTS = ("blance_ctx_create", "blance_ctx_create_multi", "blance_ctx_device_count", "blance_ctx_destroy", "blance_last_error", "blance_version", "blance_ctx_kernel_launches", "blance_plan_in_check", "blance_plan_next_map",
           "blance_plan_next_map_batch", "blance_plan_upload", "blance_plan_run", "blance_plan_fetch", "blance_plan_free", "blance_plan_timing",
           "blance_calc_partition_moves", "blance_moves_create", "blance_moves_fetch", "blance_moves_available", "blance_moves_free")


def capi():
    """ctypes handle of libblance_b200.so with argtypes set (the same symbols a cgo shim binds)."""
    global _CAPI
    if _CAPI is None:
        lib = ctypes.CDLL(os.environ.get("BLANCE_B200_LIB", _build.lib_path()))   # override: instrumented builds
        vp, i32 = ctypes.c_void_p, ctypes.c_int32
        lib.blance_ctx_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
        lib.blance_ctx_create_multi.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
        lib.blance_ctx_device_count.argtypes = [vp]
        lib.blance_ctx_destroy.argtypes = [vp]
        lib.blance_ctx_destroy.restype = None
        lib.blance_last_error.argtypes = [vp]
        lib.blance_last_error.restype = ctypes.c_char_p
        lib.blance_ctx_kernel_launches.argtypes = [vp]
        lib.blance_ctx_kernel_launches.restype = ctypes.c_int64
        lib.blance_plan_in_check.argtypes = [vp, ctypes.c_char_p, i32]
        lib.blance_plan_next_map.argtypes = [vp, vp, vp]
        lib.blance_plan_next_map_batch.argtypes = [vp, i32, vp, vp]
        lib.blance_plan_upload.argtypes = [vp, vp, ctypes.POINTER(vp)]
        lib.blance_plan_run.argtypes = [vp, vp]
        lib.blance_plan_fetch.argtypes = [vp, vp, vp]
        lib.blance_plan_timing.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float),
                                           ctypes.POINTER(ctypes.c_int32)]
        lib.blance_plan_free.argtypes = [vp, vp]
        lib.blance_plan_free.restype = None
        lib.blance_calc_partition_moves.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp]
        lib.blance_moves_create.argtypes = [vp, i32, i32, i32, vp, vp, vp, i32, i32, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int64)]
        lib.blance_moves_fetch.argtypes = [vp, vp, vp, vp, vp, vp]
        lib.blance_moves_available.argtypes = [vp, vp, vp, vp, vp, vp]
        lib.blance_moves_free.argtypes = [vp, vp]
        lib.blance_moves_free.restype = None
        _CAPI = lib
    return _CAPI


PlanIn = _PlanIn
PlanOut = _PlanOut
