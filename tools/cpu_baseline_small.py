"""CPU arm for BASELINE.json configs[0..2] ("time it on cfg-1/2 fully, cfg-3 fully if
< 30 min"): the literal C++ restatement of the Go planner (oracle/literal.cpp) and the array-form oracle
(oracle/fast.c) on the complete cfg-1, cfg-2 and cfg-3 rebalance plans, one core (the reference planner is
single-goroutine).  No GPU.  Prints one JSON line.  Oracle code is test infrastructure: this tool only
measures it."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from blance_b200 import synth, tables
from oracle_loader import fast_lib_path, literal

fast = ctypes.CDLL(fast_lib_path())
fast.oracle_fast_plan_next_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
L = literal()
out = {"cores": 1, "kind": "port", "maps_equal": True, "configs": {}}
for cfg in (1, 2, 3):
    t = synth.make_fresh(cfg)
    if cfg != 1:                        # cfg-2/3: prev = output of a fresh plan, then remove / add nodes
        fr = tables.PlanResult(t)
        s0 = t.struct()
        fast.oracle_fast_plan_next_map(ctypes.byref(s0), ctypes.byref(fr.out))
        t = synth.make_rebalance(cfg, fr.next_rows)
    ref = tables.PlanResult(t)
    s = t.struct()
    t0 = time.perf_counter()
    fast.oracle_fast_plan_next_map(ctypes.byref(s), ctypes.byref(ref.out))
    t_fast = time.perf_counter() - t0
    kw = synth.to_dicts(t, cfg)
    t0 = time.perf_counter()
    r = L.plan_next_map_ex(**kw)
    t_lit = time.perf_counter() - t0
    assert r["iterations"] == ref.iters_run and int(r["steps"]) == int(ref.steps)
    # the two oracles agree on the complete plan (the same check tests/test_synth.py makes at reduced sizes)
    want = synth.to_dicts(t, cfg)["partitions_to_assign"]
    nodes = ["n%04d" % i for i in range(t.n_nodes)]
    states = ["primary", "replica", "standby"][:t.n_states]
    rows, shape = np.asarray(ref.next_rows).reshape(t.n_parts, -1), np.asarray(ref.next_shape).reshape(t.n_parts, -1)
    for p in range(t.n_parts):
        got = {}
        for si in range(t.n_states):
            if shape[p, si] == 0:
                continue
            lo, hi = int(t.state_slot_off[si]), int(t.state_slot_off[si + 1])
            got[states[si]] = None if shape[p, si] == 1 else [nodes[x] for x in rows[p, lo:hi] if x >= 0]
        assert got == r["next_map"][str(p)], (cfg, p)
    assert len(r["next_map"]) == len(want)
    out["configs"]["cfg%d" % cfg] = {
        "partitions": int(t.n_parts), "nodes": int(t.n_nodes), "iterations": int(ref.iters_run), "steps": int(ref.steps),
        "literal": {"seconds": t_lit, "partitions_per_s": t.n_parts / t_lit, "findBestNodes_steps_per_s": ref.steps / t_lit},
        "fast_oracle": {"seconds": t_fast, "partitions_per_s": t.n_parts / t_fast, "findBestNodes_steps_per_s": ref.steps / t_fast}}
    print("cfg%d done: literal %.1f s, fast %.3f s" % (cfg, t_lit, t_fast), file=sys.stderr, flush=True)
print(json.dumps(out))
