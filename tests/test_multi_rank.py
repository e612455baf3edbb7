"""World-size-2 host logic on CPU (gloo): the multi-GPU path is "replicas / instance
sharding with no data-path collective" (DESIGN.md section 5), so what the ranks share is
only (a) the round-robin assignment of a batch's instances to ranks and (b) the
barrier + max-over-ranks of the per-rank times that bench.py reports.  The per-rank
planner is stood in for by the CPU oracle here (no GPU on this box)."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shard(n_instances, rank, world):
    """instance i -> rank i mod world."""
    return [i for i in range(n_instances) if i % world == rank]


def _worker(rank, world, port, n_inst, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_loader import fast_lib_path
    from blance_b200 import synth, tables
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    fast = ctypes.CDLL(fast_lib_path())
    fast.oracle_fast_plan_next_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    mine = shard(n_inst, rank, world)
    sums = {}
    for i in mine:
        t = synth.make_fresh(5, seed_offset=i, P=64 + i)
        r = tables.PlanResult(t)
        s = t.struct()
        fast.oracle_fast_plan_next_map(ctypes.byref(s), ctypes.byref(r.out))
        sums[i] = int(np.asarray(r.next_rows, np.int64).sum())
    local_time = torch.tensor([1.0 + rank], dtype=torch.float64)     # stand-in for the per-rank device time
    dist.barrier()
    dist.all_reduce(local_time, op=dist.ReduceOp.MAX)
    parts = torch.tensor([sum(64 + i for i in mine)], dtype=torch.int64)
    dist.all_reduce(parts, op=dist.ReduceOp.SUM)
    out_q.put((rank, sums, float(local_time[0]), int(parts[0])))
    dist.destroy_process_group()


def test_instance_sharding_world_size_2():
    n_inst, world = 7, 2
    assert sorted(shard(n_inst, 0, world) + shard(n_inst, 1, world)) == list(range(n_inst))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_inst, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = {}
    for rank, sums, tmax, total in res:
        assert tmax == 2.0                       # max over ranks, identical on every rank
        assert total == sum(64 + i for i in range(n_inst))
        assert set(sums) == set(shard(n_inst, rank, world))
        merged.update(sums)
    assert set(merged) == set(range(n_inst))
