"""Computes the array-form oracle's next map of the headline workload (cfg 4, 1 048 576 x 1 024, seed offset 0) and
stores its sha256 in profiles/parity_cfg4.json, so that bench.py can print the GPU result's digest next to it
without spending two CPU minutes in every bench run.  tests/test_gpu_parity.py::test_cfg4_full_size_bit_exact
compares the full arrays on the GPU box; this file only carries the digest."""
import ctypes, hashlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from blance_b200 import synth, tables
from oracle_loader import fast_lib_path

fast = ctypes.CDLL(fast_lib_path())
fast.oracle_fast_plan_next_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
t = synth.make_rebalance(4)
r = tables.PlanResult(t)
s = t.struct()
t0 = time.time()
assert fast.oracle_fast_plan_next_map(ctypes.byref(s), ctypes.byref(r.out)) == 0
d = {"workload": "cfg4", "seed_offset": 0, "n_parts": t.n_parts, "n_nodes": t.n_nodes, "iters_run": int(r.iters_run),
     "converged": int(r.converged), "steps": int(r.steps), "sha256_next_rows": hashlib.sha256(r.next_rows.tobytes()).hexdigest(),
     "oracle": "oracle/fast.c", "oracle_seconds": round(time.time() - t0, 1)}
with open(os.path.join(ROOT, "profiles", "parity_cfg4.json"), "w") as f:
    json.dump(d, f, indent=1)
print(d)
