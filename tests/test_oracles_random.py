"""Literal oracle (string maps + sort, oracle/literal.cpp) vs array-form oracle
(oracle/fast.c through the product's interning layer) on randomised instances:
same next map (incl. nil/absent shape), same warnings, same iteration count, same
caller-map mutation.  This is what entitles the fast oracle to serve as the checker
for the GPU path at sizes the literal one cannot reach.  CPU only."""
import copy

import pytest

from oracle_loader import literal
from randgen import random_instance
from test_fast_oracle import FAST, _host

L = literal()


def run_both(kw):
    lit = L.plan_next_map_ex(**copy.deepcopy(kw))
    ip = _host.intern_plan(**copy.deepcopy(kw))
    out = _host.plan_out(ip)
    assert FAST.oracle_fast_plan_next_map(ip.in_ptr, out.out_ptr) == 0
    next_map, warnings = _host.unintern_plan(ip, out)
    return lit, next_map, warnings, out


@pytest.mark.parametrize("chunk", range(20))
def test_literal_equals_fast_on_random_instances(chunk):
    for seed in range(chunk * 150, (chunk + 1) * 150):
        kw = random_instance(seed)
        lit, next_map, warnings, out = run_both(kw)
        assert next_map == lit["next_map"], seed
        assert warnings == lit["warnings"], seed
        assert out.iters_run == lit["iterations"], seed
        assert out.steps == lit["steps"], seed
