"""blance_b200 — a B200-native (sm_100a) drop-in for the hot path of couchbase/blance:
PlanNextMapEx (api.go:147-157; plan.go:23-331) and CalcPartitionMoves (moves.go:41-119).

The product is the C-ABI library `blance_b200/lib/libblance_b200.so`
(include/blance_b200.h) plus the C++ host mirror of blance's api.go
(blance_b200/csrc/host_api.hpp).  This Python package is only a thin face over
that host mirror so tests and bench.py can call it with dicts; names follow the
reference (PlanNextMapEx, PlanNextMap, CalcPartitionMoves, ...).

There is no CPU fallback: importing works anywhere the shared objects were built,
but every compute call raises BlanceError without a CUDA device.
"""
import os

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

if not (os.path.exists(_build.lib_path()) and os.path.exists(_build.host_module_path())):
    _build.build_all()

from . import _host  # noqa: E402
from .api import (BOOSTER_CBGT_MAX, BOOSTER_NONE, BlanceError, CalcPartitionMoves, CalcPartitionMovesMap,  # noqa: E402,F401
                  NodeStateOp, PlanNextMap, PlanNextMapEx, PlanNextMapOptions, capi)

__all__ = ["PlanNextMap", "PlanNextMapEx", "PlanNextMapOptions", "CalcPartitionMoves", "CalcPartitionMovesMap",
           "NodeStateOp", "BlanceError", "BOOSTER_NONE", "BOOSTER_CBGT_MAX", "capi"]
