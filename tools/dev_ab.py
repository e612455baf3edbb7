"""A/B harness for kernel variants (run on the GPU box): for every library given, plans the same cfg-4 cluster in a
fresh process, prints the pass time, the leader's cycle breakdown (timing builds) and a digest of the result - all
variants must print the same digest.  usage: dev_ab.py P lib1.so lib2.so ..."""
import hashlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] != "--child":
    P = sys.argv[1]
    for lib in sys.argv[2:]:
        env = dict(os.environ, BLANCE_B200_LIB=lib, BLANCE_SPEC_STATS="1")
        r = subprocess.run([sys.executable, __file__, "--child", P], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=90)
        print("=== %s\n%s" % (lib, r.stdout[-1500:]), flush=True)
    sys.exit(0)
sys.path.insert(0, ROOT)
from blance_b200 import synth, tables
P = int(sys.argv[2])
t = synth.make_rebalance(4, P=P)
ctx = tables.Context()
ctx.plan_next_map(t)
r = ctx.plan_next_map(t)
print("P %d pass_ms %.1f kernel_ms %.1f accepted %d sha %s" % (P, r.pass_ms, r.kernel_ms, r.sticky_steps, hashlib.sha256(r.next_rows.tobytes()).hexdigest()[:16]))
