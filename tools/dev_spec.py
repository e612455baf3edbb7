"""Development driver for the speculative pass kernel (run on the GPU box): parity against the array oracle
on reduced cfg-4 clusters and a few random tables, with BLANCE_SPEC_STATS counters and the pass time."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["BLANCE_SPEC_STATS"] = "1"
import numpy as np
from blance_b200 import synth, tables
from oracle_loader import fast_lib_path

FAST = ctypes.CDLL(fast_lib_path())
FAST.oracle_fast_plan_next_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p]

def oracle(t):
    r = tables.PlanResult(t); s = t.struct()
    assert FAST.oracle_fast_plan_next_map(ctypes.byref(s), ctypes.byref(r.out)) == 0
    return r

def check(ctx, t, name, engines=(0,)):
    ref = oracle(t)
    for e in engines:
        t.engine = e
        t0 = time.time()
        got = ctx.plan_next_map(t)
        dt = time.time() - t0
        ok = (np.array_equal(got.next_rows, ref.next_rows) and np.array_equal(got.next_shape, ref.next_shape)
              and np.array_equal(got.warn, ref.warn) and (got.iters_run, got.converged, got.steps) == (ref.iters_run, ref.converged, ref.steps))
        nbad = int((got.next_rows != ref.next_rows).any(axis=1).sum()) if got.next_rows.shape == ref.next_rows.shape else -1
        print("%s engine %d: %s  (bad rows %d, iters %d/%d, steps %d, accepted %d, pass %.1f ms, kernel %.1f ms, wall %.2f s)" % (
            name, e, "OK" if ok else "MISMATCH", nbad, got.iters_run, ref.iters_run, got.steps, got.sticky_steps, got.pass_ms, got.kernel_ms, dt), flush=True)
    return ok

if __name__ == "__main__":
    sizes = [int(x) for x in sys.argv[1:]] or [4096, 32768]
    ctx = tables.Context()
    allok = True
    for P in sizes:
        allok &= check(ctx, synth.make_rebalance(4, P=P), "cfg4 P=%d" % P, engines=(0, 2))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_parity as T
    for seed in range(12):
        allok &= check(ctx, T.random_tables(seed), "random %d" % seed)
    print("ALL OK" if allok else "FAILURES")
