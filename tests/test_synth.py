"""The flat-table workload generator (blance_b200/synth.py) agrees with the string
API's interning layer, and the two CPU oracles agree on the BASELINE.json
configurations that fit a CPU test (cfg 1, cfg 2, a reduced cfg 3 and cfg 4).  CPU only."""
import ctypes

import numpy as np
import pytest

from oracle_loader import literal
from test_fast_oracle import FAST, _host

from blance_b200 import synth, tables


def fast_on_tables(t):
    r = tables.PlanResult(t)
    s = t.struct()
    assert FAST.oracle_fast_plan_next_map(ctypes.byref(s), ctypes.byref(r.out)) == 0
    return r


def interned_equals_tables(t, kw):
    ip = _host.intern_plan(**kw)
    d = ip.tables()
    assert (ip.n_nodes, ip.n_states, ip.n_parts, ip.n_slots) == (t.n_nodes, t.n_states, t.n_parts, t.n_slots)
    assert np.array_equal(d["prev_rows"], np.asarray(t.prev_rows).reshape(-1))
    assert np.array_equal(d["cur_rows"], np.asarray(t.cur_rows).reshape(-1))
    assert np.array_equal(d["prev_shape"], np.asarray(t.prev_shape).reshape(-1))
    assert np.array_equal(d["node_removed"], t.node_removed) and np.array_equal(d["node_added"], t.node_added)
    assert np.array_equal(d["part_weight"] * (np.asarray(t.part_has_weight) > 0), t.part_weight * (np.asarray(t.part_has_weight) > 0))
    if t.n_rules:
        assert ip.n_hier_bits >= t.n_hier_bits
        if ip.n_hier_bits == t.n_hier_bits:
            assert np.array_equal(d["ie_mask"], np.asarray(t.ie_mask).reshape(-1))
    return ip


@pytest.mark.parametrize("cfg,P,N", [(1, None, None), (2, 512, None), (2, None, None), (3, 512, None), (4, 96, None)])
def test_two_stage_scenario_literal_vs_fast(cfg, P, N):
    L = literal()
    fresh = synth.make_fresh(cfg, P=P, N=N)
    kw = synth.to_dicts(fresh, cfg)
    ip = interned_equals_tables(fresh, kw)
    r1 = fast_on_tables(fresh)
    lit = L.plan_next_map_ex(**kw)
    out = _host.plan_out(ip)
    assert FAST.oracle_fast_plan_next_map(ip.in_ptr, out.out_ptr) == 0
    nm, wn = _host.unintern_plan(ip, out)
    assert nm == lit["next_map"] and wn == lit["warnings"]
    assert np.array_equal(out.next_rows, r1.next_rows.reshape(-1))
    if cfg == 1:
        return
    reb = synth.make_rebalance(cfg, None if cfg == 4 else r1.next_rows, P=P, N=N)
    kw = synth.to_dicts(reb, cfg)
    ip = interned_equals_tables(reb, kw)
    r2 = fast_on_tables(reb)
    lit = L.plan_next_map_ex(**kw)
    out = _host.plan_out(ip)
    assert FAST.oracle_fast_plan_next_map(ip.in_ptr, out.out_ptr) == 0
    nm, wn = _host.unintern_plan(ip, out)
    assert nm == lit["next_map"] and wn == lit["warnings"]
    assert np.array_equal(out.next_rows, r2.next_rows.reshape(-1))
    assert out.iters_run == lit["iterations"] == r2.iters_run
    # the rebalance drains every removed node
    removed = np.nonzero(reb.node_removed)[0]
    assert not np.isin(r2.next_rows, removed).any()


def test_splitmix64_known_answer():
    # splitmix64 with seed 0: the published first outputs
    v = synth.splitmix64(0, 3)
    assert [int(x) for x in v] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
