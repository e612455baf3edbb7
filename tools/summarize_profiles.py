"""Turns the raw captures that a gpurun call leaves in gpurun_out/ into the text summaries kept under profiles/:
  gpurun_out/r2_launches.csv            -> profiles/r2_launches_summary.txt   (share of the step per kernel)
  gpurun_out/r2_spec_full.ncu-rep       -> profiles/r2_ncu_assign_pass_spec.txt, profiles/ncu_traffic.json
  gpurun_out/r2_fullsize_pass_times.log -> profiles/r2_fullsize_pass_times.txt
  blance_b200/lib/libblance_b200.so     -> profiles/r2_sass_hot.txt           (TMA / mbarrier / redux / red instructions)
Run here (no GPU needed): ncu and cuobjdump only read files."""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
STEPS_IN_CAPTURE = 32768          # tools/dev_one.py 32768: steps of one pass

def launches():
    with open("gpurun_out/r2_launches.csv") as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0.0, 0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
        agg[row["Kernel Name"]][0] += ms; agg[row["Kernel Name"]][1] += 1
    tot = sum(a[0] for a in agg.values())
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 1 --warmup 3 --parts 131072 --no-cpu-baseline --no-string-api --no-batch",
           "# (reduced partition count so the serialised capture stays short; the SHARE of the step is what matters; the first 900 launches = 1.6 plans)",
           "# share   total_ms  launches  kernel"]
    for k, (ms, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
        out.append("%7.3f%% %9.3f %6d  %s" % (100 * ms / tot, ms, n, k[:150]))
    out.append("# total %.3f ms over %d launches" % (tot, sum(a[1] for a in agg.values())))
    open("profiles/r2_launches_summary.txt", "w").write("\n".join(out) + "\n")

def full_capture():
    raw = subprocess.run(["ncu", "-i", "gpurun_out/r2_spec_full.ncu-rep", "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
    rd = float(d["dram__bytes_read.sum"][0]) * mult[d["dram__bytes_read.sum"][1]]
    wr = float(d["dram__bytes_write.sum"][0]) * mult[d["dram__bytes_write.sum"][1]]
    per = (rd + wr) / STEPS_IN_CAPTURE
    ms = float(d["gpu__time_duration.sum"][0])
    stall = lambda k: float(d["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % k][0])
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__shared_mem_per_block_static", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]
    keys += [h for h in rows[0] if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    out = ["# ncu --set full --clock-control none --import-source on -k regex:k_assign_pass_spec --launch-skip 3 -c 1 python tools/dev_one.py 32768",
           "# (gpurun_out/r2_spec_full.ncu-rep; one B200; the captured launch = %s, the replica pass of iteration 2 of a" % d["Kernel Name"][0].replace("void ", ""),
           "#  32 768-partition x 1 024-node cfg-4 plan: 32 768 steps)",
           "#",
           "# reading: grid %s x %s threads (leader warp + 9 scout warps + 2 warps that exit), %s registers, %.0f KB dynamic shared memory" % (
               d["launch__grid_size"][0], d["launch__block_size"][0], d["launch__registers_per_thread"][0], float(d["launch__shared_mem_per_block_dynamic"][0])),
           "# (mirror + ring).  %.2f ms for 32 768 steps = %.0f ns per step.  DRAM: %.1f MB read, %.1f MB written = %.0f B per step (the step records" % (ms, ms * 1e6 / STEPS_IN_CAPTURE, rd / 1e6, wr / 1e6, per),
           "# and qstat words, once, through the TMA ring); the algorithmic figure of the roofline line is 16 548 B per step - the kernel proves most",
           "# results without touching those bytes, and what it touches sits in shared memory / L2.  sm__throughput %.2f %%: one SM of 148, and on" % float(d["sm__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
           "# that SM the critical path is ONE warp (the leader): stall reasons per issue are 'wait' (fixed-latency dependency) %.1f, short" % stall("wait"),
           "# scoreboard (shared-memory loads) %.1f, long scoreboard (L2 loads) %.1f, barrier %.2f (round 1's sequencer kernel: 7.1) - dependent-issue" % (stall("short_scoreboard"), stall("long_scoreboard"), stall("barrier")),
           "# latency, not bandwidth.",
           "#"]
    out += ["%-100s %s %s" % (k, d[k][0], d[k][1]) for k in keys if k in d]
    open("profiles/r2_ncu_assign_pass_spec.txt", "w").write("\n".join(out) + "\n")
    json.dump({"source": "profiles/r2_ncu_assign_pass_spec.txt: (dram__bytes_read.sum + dram__bytes_write.sum) / 32768 steps of the captured k_assign_pass_spec<2> launch",
               "dram_bytes_per_step": round(per, 1)}, open("profiles/ncu_traffic.json", "w"))

def pass_times():
    lines = open("gpurun_out/r2_fullsize_pass_times.log").read().strip().split("\n")
    t = [float(re.search(r": ([0-9.]+) ms", l).group(1)) for l in lines if "assign pass" in l]
    hdr = ["# BLANCE_PASS_TIMES=1 python tools/dev_one.py 1048576   (one B200; the production library; cfg-4 shape: 1 048 576 partitions x 1 024 nodes)",
           "# pass 2k = state 'primary' of iteration k+1 (K = 1), pass 2k+1 = state 'replica' (K = 2); 1 048 576 findBestNodes steps per pass",
           "# iteration 1 (%.1f + %.1f ms) = %.0f %% of the %.0f ms; primary passes of the later iterations (~50 movers each): %.2f ms = %.1f ns per" % (
               t[0], t[1], 100 * (t[0] + t[1]) / sum(t), sum(t), t[2], t[2] * 1e6 / 1048576),
           "# step; replica passes of the later iterations (~13 k movers each): %.1f-%.1f ms" % (min(t[3::2]), max(t[3::2]))]
    open("profiles/r2_fullsize_pass_times.txt", "w").write("\n".join(hdr + lines) + "\n")

def sass():
    txt = subprocess.run(["cuobjdump", "-sass", "blance_b200/lib/libblance_b200.so"], stdout=subprocess.PIPE, text=True).stdout.split("\n")
    s = [i for i, l in enumerate(txt) if "specILi2" in l and "Function" in l][0]
    e = [i for i, l in enumerate(txt) if "Function" in l and i > s][0]
    ins = []
    for l in txt[s:e]:
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(1), m.group(2).strip()))
    cnt = collections.Counter()
    for a, t in ins:
        op = t.split()[0] if not t.startswith("@") else t.split()[1]
        cnt[op.split(".")[0]] += 1
    keys = ["UBLKCP", "SYNCS", "CREDUX", "REDG", "ATOMG", "BAR", "LDS", "STS", "LDG", "STG", "DFMA", "DMUL", "DADD", "DSETP", "VOTE", "SHFL", "POPC", "NANOSLEEP", "MEMBAR", "FENCE"]
    out = ["# SASS of k_assign_pass_spec<2> (the replica pass of the headline workload) in blance_b200/lib/libblance_b200.so",
           "# cuobjdump -sass blance_b200/lib/libblance_b200.so, function _ZN10blance_dev18k_assign_pass_specILi2EEEv5DPooliiji",
           "# %d instructions; opcode counts of the ones that matter for the design:" % len(ins)]
    out += ["#   %-10s %5d" % (k, cnt[k]) for k in keys if cnt[k]]
    out += ["#", "# TMA bulk copies (global -> shared, completion on an mbarrier), mbarrier operations, warp reductions, value-less atomics, named barriers, fences:"]
    out += ["  /*%s*/  %s" % (a, t) for a, t in ins if re.search(r"UBLKCP|SYNCS|CREDUX|REDG|NANOSLEEP|BAR\.|MEMBAR|FENCE", t)]
    open("profiles/r2_sass_hot.txt", "w").write("\n".join(out) + "\n")

for f in (launches, full_capture, pass_times, sass):
    try:
        f()
    except Exception as ex:       # a missing capture leaves its summary as it was
        print("%s: skipped (%s)" % (f.__name__, ex), file=sys.stderr)
