"""The array-form CPU oracle (oracle/fast.c), driven through the product's own
interning layer (blance_b200/csrc/host_api.cpp), must reproduce every golden
vector of the reference; this also pins the interning / un-interning code that
the GPU path shares.  CPU only."""
import ctypes

import pytest

import golden_util as G
from oracle_loader import fast_lib_path

from blance_b200 import _host

FAST = ctypes.CDLL(fast_lib_path())
FAST.oracle_fast_plan_next_map.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
FAST.oracle_fast_plan_next_map_capped.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]


def fast_plan(kwargs):
    ip = _host.intern_plan(**kwargs)
    out = _host.plan_out(ip)
    assert FAST.oracle_fast_plan_next_map(ip.in_ptr, out.out_ptr) == 0
    next_map, warnings = _host.unintern_plan(ip, out)
    return next_map, warnings, out


@pytest.mark.parametrize("c", G.plan_cases(), ids=G.case_id)
def test_fast_oracle_golden(c):
    next_map, warnings, _ = fast_plan(G.plan_kwargs(c))
    assert next_map == G.pmap(c["exp"])
    assert G.count_warnings(c, warnings) == c["expNumWarnings"]
