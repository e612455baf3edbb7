"""CPU arm of BASELINE.json configs[4] ("cfg-5 additionally on all host cores, one
instance per thread"): the literal C++ restatement of the Go planner (oracle/literal.cpp - string maps and
sort, the reference's data structures) on a bounded sample of the 1 024 cfg-5 instances (1 024 partitions x
64 nodes, rack rule), one instance per process over all host cores, fresh placement and rebalance; and the
array-form oracle (oracle/fast.c) the same way as the "best CPU" line.  No GPU.  Prints one JSON line.
Oracle code is test infrastructure: this tool only measures it."""
import ctypes
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _tables(i, stage):
    from blance_b200 import synth, tables
    from oracle_loader import fast_lib_path
    fast = ctypes.CDLL(fast_lib_path())
    fast.oracle_fast_plan_next_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    t = synth.make_fresh(5, seed_offset=i)
    if stage == "fresh":
        return t, fast, synth, tables
    ref = tables.PlanResult(t)
    s = t.struct()
    fast.oracle_fast_plan_next_map(ctypes.byref(s), ctypes.byref(ref.out))
    return synth.make_rebalance(5, ref.next_rows, seed_offset=i), fast, synth, tables


def run_literal(arg):
    i, stage = arg
    from oracle_loader import literal
    t, _, synth, _ = _tables(i, stage)
    kw = synth.to_dicts(t, 5)
    L = literal()
    t0 = time.perf_counter()
    r = L.plan_next_map_ex(**kw)
    return time.perf_counter() - t0, int(r["steps"]), t.n_parts


def run_fast(arg):
    i, stage = arg
    t, fast, _, tables = _tables(i, stage)
    ref = tables.PlanResult(t)
    s = t.struct()
    t0 = time.perf_counter()
    fast.oracle_fast_plan_next_map(ctypes.byref(s), ctypes.byref(ref.out))
    return time.perf_counter() - t0, int(ref.steps), t.n_parts


def arm(fn, stage, n, cores):
    with mp.Pool(cores) as pool:
        pool.map(fn, [(0, stage)] * cores)                    # load the libraries, warm the caches
        t0 = time.perf_counter()
        res = pool.map(fn, [(i, stage) for i in range(n)], chunksize=1)
        wall = time.perf_counter() - t0
    parts = sum(r[2] for r in res)
    return {"partitions_per_s": parts / wall, "instances_per_s": n / wall, "wall_s": wall, "instances": n,
            "cpu_s_per_instance": sum(r[0] for r in res) / n, "steps": sum(r[1] for r in res)}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    cores = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    out = {"workload": "cfg5 sample: %d of 1024 instances x (1024 partitions x 64 nodes), k=(1,1), rack rule" % n,
           "cores": cores, "kind": "port",
           "literal": {st: arm(run_literal, st, n, cores) for st in ("fresh", "rebalance")},
           "fast_oracle": {st: arm(run_fast, st, max(16 * n, 64 * cores), cores) for st in ("fresh", "rebalance")}}
    print(json.dumps(out))
