"""The marshalling layer (PartitionMap <-> flat rows) splits big maps over
host threads.  These tests drive maps above the 32 768-partition threshold through it: the
tables must not depend on the thread count, lists longer than the constraints must widen the
slot layout (second pass), names outside nodesAll and non-model states must survive, and the
result must still equal the literal oracle's.  CPU only."""
import copy
import random

import numpy as np
import pytest

from oracle_loader import literal
from test_fast_oracle import FAST, _host

L = literal()
P_BIG = 40000


def big_instance(seed, wide_lists):
    rnd = random.Random(seed)
    nodes = ["n%02d" % i for i in range(9)]
    names = [str(i) for i in range(P_BIG - 3)] + ["007", "p-x", "+5"]      # numeric, padded-tie and raw names
    prev = {}
    for n in names:
        a = rnd.sample(nodes, 3)
        row = {"primary": [a[0]], "replica": [a[1]] if rnd.random() < 0.9 else None}
        if rnd.random() < 0.001:
            row["replica"] = ["ghost%d" % rnd.randint(0, 3)]                # nodes outside nodesAll
        if wide_lists and rnd.random() < 0.0005:
            row["replica"] = [a[1], a[2], a[0]]                             # longer than constraints: widens the layout
        if rnd.random() < 0.0005:
            row["dead"] = [a[2]]                                            # non-model state (feeds the totals only)
        prev[n] = row
    assign = {n: {s: (None if v is None else list(v)) for s, v in prev[n].items() if s != "dead"} for n in names
              if rnd.random() < 0.97}
    return dict(prev_map=prev, partitions_to_assign=assign, nodes_all=nodes, nodes_to_remove=[nodes[0]],
                nodes_to_add=[nodes[8]], model={"primary": (0, 1), "replica": (1, 1)},
                partition_weights={n: rnd.randint(1, 4) for n in names[::7]}, state_stickiness={"primary": 2})


def decoded(ip):
    t = ip.tables()
    names = np.array(ip.node_names + [""], dtype=object)

    def rows(key):
        return names[np.where(t[key] < 0, len(names) - 1, t[key])]

    out = {k: v for k, v in t.items() if k not in ("prev_rows", "cur_rows", "node_removed", "node_added")}
    out["prev_rows"], out["cur_rows"] = rows("prev_rows"), rows("cur_rows")
    out["part_names"] = np.array(ip.part_names, dtype=object)
    out["removed"] = sorted(n for n, f in zip(ip.node_names, t["node_removed"]) if f)
    out["added"] = sorted(n for n, f in zip(ip.node_names, t["node_added"]) if f)
    return out


@pytest.mark.parametrize("wide_lists", [False, True])
def test_tables_do_not_depend_on_thread_count(wide_lists):
    kw = big_instance(11 + wide_lists, wide_lists)
    got = {}
    try:
        for threads in (1, 2, 7):
            _host.set_host_threads(threads)
            assert _host.host_threads() == threads
            got[threads] = decoded(_host.intern_plan(**copy.deepcopy(kw)))
    finally:
        _host.set_host_threads(0)
    assert list(got[1]["state_slot_off"]) == ([0, 1, 4] if wide_lists else [0, 1, 2])
    for threads in (2, 7):
        assert got[threads].keys() == got[1].keys()
        for k in got[1]:
            assert np.array_equal(got[threads][k], got[1][k]), (threads, k)


def test_big_map_through_the_threads_equals_literal_oracle():
    kw = big_instance(5, True)
    lit = L.plan_next_map_ex(**copy.deepcopy(kw))
    try:
        _host.set_host_threads(5)
        ip = _host.intern_plan(**copy.deepcopy(kw))
        out = _host.plan_out(ip)
        assert FAST.oracle_fast_plan_next_map(ip.in_ptr, out.out_ptr) == 0
        next_map, warnings = _host.unintern_plan(ip, out)
    finally:
        _host.set_host_threads(0)
    assert next_map == lit["next_map"]
    assert warnings == lit["warnings"]
    assert out.iters_run == lit["iterations"]
