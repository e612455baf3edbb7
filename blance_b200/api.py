"""Python face of the host API.  Mirrors the reference's exported names:

    PlanNextMapEx(prevMap, partitionsToAssign, nodesAll, nodesToRemove, nodesToAdd, model, options)
        -> (nextMap, warnings)                                    api.go:147-157
    PlanNextMap(..., modelStateConstraints, partitionWeights, stateStickiness, nodeWeights,
                nodeHierarchy, hierarchyRules) -> (nextMap, warnings)   api.go:109-132 (deprecated wrapper)
    CalcPartitionMoves(states, begNodesByState, endNodesByState, favorMinNodes) -> [NodeStateOp]   moves.go:41-46

A PartitionMap is {partitionName: {stateName: [nodeName, ...] | None}}; a
PartitionModel is {stateName: (priority, constraints)}; HierarchyRules is
{stateName: [(includeLevel, excludeLevel), ...]}.  prevMap and
partitionsToAssign are mutated in place exactly as plan.go:49-52 mutates them.
"""
import collections
import ctypes
import dataclasses
import typing

from . import _host
from . import build as _build

BlanceError = _host.BlanceError
BOOSTER_NONE = 0
BOOSTER_CBGT_MAX = 1     # cbgt's max(float64(-w), stickiness), control_test.go:19-26

NodeStateOp = collections.namedtuple("NodeStateOp", ["Node", "State", "Op"])   # moves.go:17-21


@dataclasses.dataclass
class PlanNextMapOptions:          # api.go:183-190 (+ the package-level hooks of plan.go:21,693)
    ModelStateConstraints: typing.Optional[dict] = None
    PartitionWeights: typing.Optional[dict] = None
    StateStickiness: typing.Optional[dict] = None
    NodeWeights: typing.Optional[dict] = None
    NodeHierarchy: typing.Optional[dict] = None
    HierarchyRules: typing.Optional[dict] = None
    MaxIterationsPerPlan: int = 10
    NodeScoreBooster: int = BOOSTER_NONE
    Engine: int = 0


def _replace(dst, src):
    dst.clear()
    dst.update(src)


def PlanNextMapEx(prevMap, partitionsToAssign, nodesAll, nodesToRemove, nodesToAdd, model, options=None,
                  stats=None):
    o = options or PlanNextMapOptions()
    same = prevMap is partitionsToAssign
    r = _host.PlanNextMapEx(prevMap, None if same else partitionsToAssign, list(nodesAll),
                            None if nodesToRemove is None else list(nodesToRemove),
                            None if nodesToAdd is None else list(nodesToAdd),
                            {k: tuple(v) for k, v in model.items()},
                            o.ModelStateConstraints, o.PartitionWeights, o.StateStickiness, o.NodeWeights,
                            o.NodeHierarchy,
                            None if o.HierarchyRules is None else {k: [tuple(x) for x in v] for k, v in o.HierarchyRules.items()},
                            o.NodeScoreBooster, o.MaxIterationsPerPlan, o.Engine)
    _replace(prevMap, r["prev_map"])                 # plan.go:49-52
    if not same:
        _replace(partitionsToAssign, r["partitions_to_assign"])
    if stats is not None:
        stats.update({k: r[k] for k in ("iterations", "converged", "steps", "device_ms", "kernel_ms", "pass_ms")})
    return r["next_map"], r["warnings"]


def PlanNextMap(prevMap, partitionsToAssign, nodesAll, nodesToRemove, nodesToAdd, model,
                modelStateConstraints=None, partitionWeights=None, stateStickiness=None, nodeWeights=None,
                nodeHierarchy=None, hierarchyRules=None):
    return PlanNextMapEx(prevMap, partitionsToAssign, nodesAll, nodesToRemove, nodesToAdd, model,
                         PlanNextMapOptions(modelStateConstraints, partitionWeights, stateStickiness, nodeWeights,
                                            nodeHierarchy, hierarchyRules))


def CalcPartitionMoves(states, begNodesByState, endNodesByState, favorMinNodes):
    return [NodeStateOp(*t) for t in _host.CalcPartitionMoves(list(states), begNodesByState, endNodesByState,
                                                              bool(favorMinNodes))]


def CalcPartitionMovesMap(states, begMap, endMap, favorMinNodes):
    """Vectorised CalcPartitionMoves over two PartitionMaps: {partitionName: [NodeStateOp]}."""
    r = _host.CalcPartitionMovesMap(list(states), begMap, endMap, bool(favorMinNodes))
    return {k: [NodeStateOp(*t) for t in v] for k, v in r.items()}


# ---- the raw C ABI (ctypes), for tests and bench.py -------------------------------------------

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
EXPORTS = ("blance_ctx_create", "blance_ctx_create_multi", "blance_ctx_device_count", "blance_ctx_destroy", "blance_last_error", "blance_version", "blance_ctx_kernel_launches", "blance_plan_next_map",
           "blance_plan_next_map_batch", "blance_plan_upload", "blance_plan_run", "blance_plan_fetch", "blance_plan_free", "blance_plan_timing",
           "blance_calc_partition_moves")


def capi():
    """ctypes handle of libblance_b200.so with argtypes set (the same symbols a cgo shim binds)."""
    global _CAPI
    if _CAPI is None:
        import os
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
        _CAPI = lib
    return _CAPI


PlanIn = _PlanIn
PlanOut = _PlanOut
