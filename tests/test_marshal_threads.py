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


def test_name_order_of_a_big_map_with_mixed_names():
    """plan.go:519-528 on 40 000 names: non-negative integers order by value ("%10d"), everything else by its
    raw bytes, ties on the padded form by the raw name - also when the sort is split over threads."""
    rnd = random.Random(3)
    names = set()
    while len(names) < P_BIG:
        r = rnd.random()
        v = rnd.randint(0, 99999)
        names.add(str(v) if r < 0.5 else "%07d" % v if r < 0.6 else "+%d" % v if r < 0.65 else "-%d" % v if r < 0.7
                  else "p%05d" % v if r < 0.9 else " %d" % v)
    names.add(str(12345678901))          # wider than ten digits: no longer ordered like the shorter numbers
    names.add("9999999999")
    names = sorted(names)
    rnd.shuffle(names)
    nodes = ["a", "b"]
    prev = {n: {"primary": [nodes[i & 1]]} for i, n in enumerate(names)}

    def key(n):
        try:
            ok = n.lstrip("+-").isdigit() and n.isascii() and len(n.lstrip("+-")) == len(n) - (n[0] in "+-")
            v = int(n) if ok else None
        except ValueError:
            v = None
        return ("%10d" % v if v is not None and v >= 0 else n, n)

    expect = sorted(names, key=key)
    try:
        for threads in (1, 6):
            _host.set_host_threads(threads)
            ip = _host.intern_plan(prev_map=prev, partitions_to_assign=None, nodes_all=nodes, nodes_to_remove=[],
                                   nodes_to_add=[], model={"primary": (0, 1)})
            assert ip.part_names == expect, threads
    finally:
        _host.set_host_threads(0)


def test_errors_raised_inside_the_threads_surface():
    kw = big_instance(9, False)
    kw["partitions_to_assign"]["17"]["bogus"] = ["n01"]            # a state outside the model: the reference panics
    try:
        _host.set_host_threads(4)
        with pytest.raises(Exception, match="not in the model"):
            _host.intern_plan(**kw)
    finally:
        _host.set_host_threads(0)
