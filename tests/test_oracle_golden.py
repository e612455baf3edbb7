"""Pins the literal CPU oracle (oracle/literal.cpp) to the reference's own golden
vectors: plan_test.go:392-2863, control_test.go:18-416, moves_test.go:19-486 and
the helper tables (plan_test.go:21-390, misc_test.go:18-89), as transcribed by
tests/golden/make_fixtures.py.  Comparison is reflect.DeepEqual-shaped: same
keys, same list order, nil vs empty list distinguished."""
import pytest

import golden_util as G
from oracle_loader import literal

L = literal()


@pytest.mark.parametrize("c", G.plan_cases(), ids=G.case_id)
def test_plan_next_map_golden(c):
    r = L.plan_next_map_ex(**G.plan_kwargs(c))
    assert r["next_map"] == G.pmap(c["exp"])
    assert G.count_warnings(c, r["warnings"]) == c["expNumWarnings"]


@pytest.mark.parametrize("c", [c for c in G.plan_cases(True) if c["ignore"]], ids=G.case_id)
def test_plan_next_map_ignored_cases_weak(c):
    """The reference skips these (plan_test.go:1953,2424,2447) because its harness
    cannot spell the expected ORDER inside a state list; the node SETS it documents
    in comments are still checked where the case says the result is as drawn."""
    r = L.plan_next_map_ex(**G.plan_kwargs(c))
    assert set(r["next_map"]) == set(c["exp"])


def test_find_state_changes_golden():
    for i, c in enumerate(G.load("moves_cases.json")["findStateChanges"]):
        got = L.find_state_changes(c["begStateIdx"], c["endStateIdx"], c["state"], c["states"],
                                   c["begNodesByState"], c["endNodesByState"])
        assert got == (c["expected"] or []), i


def test_calc_partition_moves_golden():
    for c in G.load("moves_cases.json")["calcPartitionMoves"]:
        got = L.calc_partition_moves(c["states"], c["before"], c["after"], c["favorMinNodes"])
        assert len(got) == len(c["exp"]), c["index"]
        for (node, state, op), e in zip(got, c["exp"]):
            assert node == e["node"] and state == e["state"] and op in e["op"], (c["index"], got, c["exp"])


def test_orchestrate_derived_move_sequences():
    """orchestrate_test.go:1796-1808 compares the recorded (node, state) of each assignment with the
    expected list, entry by entry (the recorded list may be longer)."""
    for c in G.load("moves_cases.json")["orchestrateDerived"]:
        got = L.calc_partition_moves(c["states"], c["before"], c["after"], c["favorMinNodes"])
        assert len(got) >= len(c["exp"]), (c["label"], c["partition"])
        for (node, state, _op), e in zip(got, c["exp"]):
            assert node == e["node"] and state == e["state"], (c["label"], c["partition"], got, c["exp"])


def test_unit_tables():
    u = G.load("unit_cases.json")
    for c in u["flattenNodesByState"]:
        assert L.flatten_nodes_by_state(c["a"], ["primary", "replica"]) == c["exp"]
    for c in u["removeNodesFromNodesByState"]:
        assert L.remove_nodes_from_nodes_by_state(c["nodesByState"], c["removeNodes"]) == c["exp"]
    for c in u["stateNameSorter"]:
        m = {k: (v["Priority"], v["Constraints"]) for k, v in c["m"].items()}
        assert L.state_name_sort(m, c["s"]) == c["exp"]
    for c in u["countStateNodes"]:
        pm = {k: v["NodesByState"] for k, v in c["m"].items()}
        assert L.count_state_nodes(pm, c["w"]) == c["exp"]
    for c in u["findAncestor"]:
        assert L.find_ancestor("a", c["mapParents"], c["level"]) == c["exp"]
    for c in u["findLeaves"]:
        assert L.find_leaves("a", c["mapChildren"]) == c["exp"]
    for c in u["mapParentsToMapChildren"]:
        assert L.map_parents_to_map_children(c["in"]) == c["exp"]
    for c in u["stringsRemoveStrings"]:
        assert L.strings_remove_strings(c["a"], c["b"]) == c["exp"]
    for c in u["stringsIntersectStrings"]:
        assert L.strings_intersect_strings(c["a"], c["b"]) == c["exp"]
