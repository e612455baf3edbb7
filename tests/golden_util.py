"""Helpers shared by the golden-vector tests: load tests/golden/*.json and turn a
fixture into oracle / product call arguments."""
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def plan_cases(include_ignored=False):
    return [c for c in load("plan_cases.json") if include_ignored or not c["ignore"]]


def case_id(c):
    return "%s[%d]" % (c["group"], c["index"])


def pmap(m):
    """fixture PartitionMap -> {name: {state: [nodes]|None}} (Partition.Name == key in every fixture)."""
    if m is None:
        return None
    out = {}
    for k, v in m.items():
        assert v["name"] == k
        out[k] = {s: (list(n) if n is not None else None) for s, n in v["nodesByState"].items()}
    return out


def plan_kwargs(c):
    return dict(
        prev_map=pmap(c["prevMap"]),
        partitions_to_assign=pmap(c["partitionsToAssign"]),
        nodes_all=c["nodes"],
        nodes_to_remove=c["nodesToRemove"],
        nodes_to_add=c["nodesToAdd"],
        model={k: (v["priority"], v["constraints"]) for k, v in c["model"].items()},
        model_state_constraints=c["modelStateConstraints"],
        partition_weights=c["partitionWeights"],
        state_stickiness=c["stateStickiness"],
        node_weights=c["nodeWeights"],
        node_hierarchy=c["nodeHierarchy"],
        hierarchy_rules=(None if c["hierarchyRules"] is None else
                         {k: [(r["includeLevel"], r["excludeLevel"]) for r in v]
                          for k, v in c["hierarchyRules"].items()}),
        booster={"none": 0, "cbgt": 1}[c["booster"]],
    )


def count_warnings(c, warnings):
    """plan_test.go:1599-1603 counts warning strings, :1738 counts partitions."""
    if c["warnCount"] == "strings":
        return sum(len(v) for v in warnings.values())
    return len(warnings)
