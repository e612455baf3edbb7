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

from .abi import BlanceError     # base class; _host.BlanceError derives from it
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


# ---- the raw C ABI (ctypes) lives in abi.py; re-exported here for callers of the Python face -----------
from .abi import EXPORTS, PlanIn, PlanOut, _I32_FIELDS, _PTR_FIELDS, capi  # noqa: E402,F401
