"""The JSON wire form of a PartitionMap (api.go:30,35 tags `name` / `nodesByState`):
the host twin's encoder against Python's json with Go's encoding/json conventions (sorted keys, nil slice ->
null, HTML-safe escapes, U+2028/9, invalid UTF-8 -> U+FFFD), on the literal oracle's maps.  CPU only."""
import copy
import json

import blance_b200
from oracle_loader import literal
from randgen import random_instance
from test_fast_oracle import FAST, _host


def go_json(pmap):
    """json.Marshal(map[string]*Partition) as Go writes it."""
    obj = {k: {"name": k, "nodesByState": v} for k, v in pmap.items()}
    s = json.dumps(obj, sort_keys=True, separators=(",", ":"), ensure_ascii=False)
    # Go puts struct fields in declaration order (name, nodesByState) - sort_keys agrees ("name" < "nodesByState") -
    # and escapes <, >, &, U+2028, U+2029
    for a, b in (("<", "\\u003c"), (">", "\\u003e"), ("&", "\\u0026"), (" ", "\\u2028"), (" ", "\\u2029")):
        s = s.replace(a, b)
    return s.encode("utf-8")


def test_partition_map_json_equals_go_encoding_on_oracle_maps():
    L = literal()
    for seed in range(300):
        kw = random_instance(seed)
        lit = L.plan_next_map_ex(**copy.deepcopy(kw))
        assert _host.PartitionMapToJSON(lit["next_map"]) == go_json(lit["next_map"]), seed


def test_json_escapes_and_nil_lists():
    m = {'p"1\\': {"primary": ["a<b>&c", "tab\there", " x", "café", "\x01"], "replica": None, "z": []},
         "0": {}}
    got = _host.PartitionMapToJSON(m)
    assert got == go_json(m)
    assert json.loads(got) == {k: {"name": k, "nodesByState": v} for k, v in m.items()}


def test_plan_result_to_json_equals_map_encoder():
    """rows -> JSON directly == encoder(UninternPlan(rows)), on oracle-computed rows."""
    for seed in range(200):
        kw = random_instance(seed)
        ip = _host.intern_plan(**copy.deepcopy(kw))
        out = _host.plan_out(ip)
        assert FAST.oracle_fast_plan_next_map(ip.in_ptr, out.out_ptr) == 0
        next_map, _ = _host.unintern_plan(ip, out)
        assert _host.plan_result_to_json(ip, out) == _host.PartitionMapToJSON(next_map) == go_json(next_map), seed
