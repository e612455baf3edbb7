"""Small driver for ncu captures of the assign-pass kernel: one resident plan of a
reduced cfg-4 cluster (P partitions x 1024 nodes), run twice."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blance_b200 import synth, tables
P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
t = synth.make_rebalance(4, P=P)
t.max_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = tables.Context()
plan = ctx.upload(t)
ctx.run(plan)
ctx.run(plan)
print(ctx.timing(plan), ctx.fetch(plan, tables.PlanResult(t)).steps)
