"""BASELINE.json configs[4]: 1 024 independent PlanNextMap instances (1 024 partitions x 64 nodes,
rack rule on the replica state) as ONE blance_plan_next_map_batch call per stage, checked against the
CPU oracle on a sample of instances.  Prints one JSON line."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from blance_b200 import synth, tables
from oracle_loader import fast_lib_path

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
fast = ctypes.CDLL(fast_lib_path())
fast.oracle_fast_plan_next_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
ctx = tables.Context()
fresh = [synth.make_fresh(5, seed_offset=i) for i in range(n)]
ctx.plan_next_map_batch(fresh[:8])                     # warm-up
t0 = time.perf_counter()
r1 = ctx.plan_next_map_batch(fresh)
t_fresh = time.perf_counter() - t0
reb = [synth.make_rebalance(5, r.next_rows, seed_offset=i) for i, r in enumerate(r1)]
t0 = time.perf_counter()
r2 = ctx.plan_next_map_batch(reb)
t_reb = time.perf_counter() - t0
ok = True
for i in range(0, n, max(1, n // 16)):
    for t, r in ((fresh[i], r1[i]), (reb[i], r2[i])):
        ref = tables.PlanResult(t)
        s = t.struct()
        fast.oracle_fast_plan_next_map(ctypes.byref(s), ctypes.byref(ref.out))
        ok &= bool(np.array_equal(ref.next_rows, r.next_rows)) and ref.iters_run == r.iters_run
parts = sum(t.n_parts for t in fresh)
print(json.dumps({"workload": "cfg5: %d instances x (1024 partitions x 64 nodes), k=(1,1), rack rule" % n,
                  "fresh": {"partitions_per_s": parts / t_fresh, "seconds": t_fresh, "device_ms": r1[0].device_ms,
                            "iterations_max": max(r.iters_run for r in r1)},
                  "rebalance": {"partitions_per_s": parts / t_reb, "seconds": t_reb, "device_ms": r2[0].device_ms,
                                "iterations_max": max(r.iters_run for r in r2)},
                  "steps_total": int(sum(r.steps for r in r1) + sum(r.steps for r in r2)),
                  "sample_equals_oracle": ok}))
