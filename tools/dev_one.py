"""One cfg-4 plan of a given size through the C ABI (for ncu captures of the pass kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blance_b200 import synth, tables
P = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
t = synth.make_rebalance(4, P=P)
ctx = tables.Context()
r = ctx.plan_next_map(t)
print("steps", r.steps, "accepted", r.sticky_steps, "pass_ms", r.pass_ms)
