"""Synthetic clusters for the five configurations of BASELINE.json (BASELINE.md
section 4), generated straight into flat tables (blance_b200.tables.PlanTables).

Language-neutral recipe so any implementation can rebuild the same inputs: PRNG =
splitmix64, seed 0xB1A9CE00 + cfg (+ instance for cfg 5); node names "n%04d" in
nodesAll order; partition names decimal "0".."P-1" (so the name rank of
plan.go:525-528 is the number itself); states primary(0) / replica(1) /
standby(2).

Every configuration is a two-stage scenario: stage "fresh" assigns all partitions
to the initial nodes from an empty map; stage "rebalance" takes a previous map,
removes some nodes and adds others.  The rebalance stage of cfg 2/3/5 needs the
output of the fresh stage as its prevMap (pass `prev_rows`); cfg 4 starts from a
closed-form round-robin map.
"""
import numpy as np

from .tables import NO_NODE, SHAPE_LIST, PlanTables

MASK64 = (1 << 64) - 1
SEED_BASE = 0xB1A9CE00


def splitmix64(seed, n):
    """n consecutive outputs of splitmix64 started at `seed` (numpy uint64)."""
    with np.errstate(over="ignore"):
        x = (np.uint64(seed & MASK64) + np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


CONFIGS = {
    # cfg: P, N, constraints, tree fan-outs (leaf -> root), rules per state, initial nodes, removed, added
    1: dict(P=64, N=8, k=(1, 1), levels=(), rules={}, weights=False),
    2: dict(P=4096, N=64, k=(1, 1), levels=(8, 8), rules={1: [(2, 1)]}, weights=False),
    3: dict(P=65536, N=256, k=(1, 2, 1), levels=(8, 8, 4), rules={1: [(2, 1)], 2: [(3, 2)]}, weights=False),
    4: dict(P=1048576, N=1024, k=(1, 2), levels=(), rules={}, weights=True),
    5: dict(P=1024, N=64, k=(1, 1), levels=(8, 8), rules={1: [(2, 1)]}, weights=False),
}


def _group_of(nodes, levels, level):
    """id of the level-`level` ancestor group of each node in a regular tree with the
    given fan-outs (level 0 = the node itself); -1 past the root."""
    if level == 0:
        return nodes.copy()
    if level > len(levels):
        return np.full_like(nodes, -1)
    div = 1
    for f in levels[:level]:
        div *= f
    return nodes // div


def hierarchy_masks(N, levels, rules_by_state, n_states):
    """ie_mask[rule][anchor][word] for a regular tree: leaves(anc(a, inc)) minus
    leaves(anc(a, exc)) (plan.go:723-734).  The "" anchor (index N) and ancestors past
    the root give the set {""}, which never intersects nodesAll, hence an empty mask."""
    rule_off = [0]
    rules = []
    for s in range(n_states):
        rules += rules_by_state.get(s, [])
        rule_off.append(len(rules))
    HW = (N + 31) // 32
    mask = np.zeros((len(rules), N + 1, HW), np.uint32)
    nodes = np.arange(N)
    for r, (inc, exc) in enumerate(rules):
        gi, ge = _group_of(nodes, levels, inc), _group_of(nodes, levels, exc)
        for a in range(N):
            member = np.zeros(N, bool)
            if gi[a] >= 0:
                member = gi == gi[a]
                if ge[a] >= 0:
                    member &= ~(ge == ge[a])
            bits = np.nonzero(member)[0]
            np.bitwise_or.at(mask[r, a], bits >> 5, (np.uint32(1) << (bits & 31).astype(np.uint32)))
    return np.asarray(rule_off, np.int32), mask


def node_hierarchy_dict(N, levels):
    """The NodeHierarchy map (child -> parent names) of the regular tree, for the string API."""
    names = ["n%04d" % i for i in range(N)]
    parents = {}
    cur = names
    for lv, f in enumerate(levels):
        nxt = ["L%d_%03d" % (lv + 1, g) for g in range((len(cur) + f - 1) // f)]
        for i, c in enumerate(cur):
            parents[c] = nxt[i // f]
        cur = nxt
    return parents


def _base(cfg, P=None, N=None):
    c = dict(CONFIGS[cfg])
    if P is not None:
        c["P"] = P
    if N is not None:
        c["N"] = N
    S = len(c["k"])
    t = PlanTables(c["N"], S, c["P"], list(range(S)), c["k"])
    if c["rules"]:
        t.has_hier_rules = 1
        t.rule_off, mask = hierarchy_masks(c["N"], c["levels"], c["rules"], S)
        t.n_rules = mask.shape[0]
        t.ie_mask = mask.reshape(-1)
        t.n_hier_bits = c["N"]
    return c, t


def removed_added(cfg, N, levels):
    """Which nodes the rebalance stage removes / adds."""
    if cfg == 1:
        return np.array([], int), np.arange(N)
    if cfg in (2, 5):          # -4 / +4 nodes: the last 4 are the new ones, the first of 4 racks lose one each
        return np.array([0, 8, 16, 24]) % max(N - 4, 1), np.arange(N - 4, N)
    if cfg == 3:               # -1 rack / +1 rack
        return np.arange(0, 8), np.arange(N - 8, N)
    if cfg == 4:               # -16 / +16
        return np.arange(0, 16), np.arange(N - 16, N)
    raise ValueError(cfg)


def make_fresh(cfg, seed_offset=0, P=None, N=None):
    """Stage 1: empty prevMap, every partition to assign, all initial nodes added."""
    c, t = _base(cfg, P, N)
    _, added = removed_added(cfg, c["N"], c["levels"])
    if cfg == 1:
        initial = np.arange(c["N"])
    else:
        initial = np.setdiff1d(np.arange(c["N"]), added)
    # nodes that only join later are treated as removed-from-consideration here: the fresh
    # stage simply lists them in nodesToRemove (they hold nothing, so nothing moves)
    t.node_removed[np.setdiff1d(np.arange(c["N"]), initial)] = 1
    t.node_added[initial] = 1
    t.part_in_prev[:] = 0
    # the reference needs prevMap entries whenever nodesToRemove is non-empty (plan.go:544):
    if t.node_removed.any():
        t.part_in_prev[:] = 1
    _apply_weights(cfg, c, t, seed_offset)
    return t


def _apply_weights(cfg, c, t, seed_offset):
    if not c["weights"]:
        return
    N, P = t.n_nodes, t.n_parts
    seed = SEED_BASE + cfg + seed_offset
    r = splitmix64(seed, N + 2 * P)
    u = (r[:N] >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    t.has_node_weights = 1
    t.node_has_weight[:] = 1
    t.node_weight[:] = np.where(u < 0.4, 1, np.where(u < 0.7, 2, np.where(u < 0.9, 3, 4)))
    t.has_part_weights = 1
    sel = (r[N:N + P] % np.uint64(4)) == 0                    # 25 % of the partitions carry a weight
    t.part_has_weight[:] = sel
    t.part_weight[:] = np.where(sel, (r[N + P:N + 2 * P] % np.uint64(8)).astype(np.int32) + 1, 1)
    t.state_has_stickiness[:] = 1                            # StateStickiness{primary:3, replica:2}
    t.state_stickiness[:] = [3, 2][:t.n_states] + [2] * max(0, t.n_states - 2)


def round_robin_rows(t, n_initial):
    """cfg 4's previous map: primary = p mod n_initial, replica j = (p + (j+1)*337) mod n_initial."""
    P = t.n_parts
    p = np.arange(P, dtype=np.int64)
    rows = np.full((P, t.n_slots), NO_NODE, np.int32)
    slot = 0
    for s in range(t.n_states):
        for j in range(int(t.state_constraints[s])):
            rows[:, slot] = (p + slot * 337) % n_initial
            slot += 1
    return rows


def make_rebalance(cfg, prev_rows=None, seed_offset=0, P=None, N=None):
    """Stage 2: prevMap = partitionsToAssign = `prev_rows` ([P][slots]; for cfg 4 the
    round-robin map when omitted), with the configuration's nodes removed / added."""
    c, t = _base(cfg, P, N)
    removed, added = removed_added(cfg, c["N"], c["levels"])
    if prev_rows is None:
        if cfg != 4:
            raise ValueError("cfg %d needs the fresh stage's rows as prev_rows" % cfg)
        prev_rows = round_robin_rows(t, c["N"] - len(added))
    prev_rows = np.ascontiguousarray(prev_rows, np.int32).reshape(t.n_parts, t.n_slots)
    t.prev_rows[:] = prev_rows
    t.cur_rows[:] = prev_rows
    t.prev_shape[:] = SHAPE_LIST
    t.cur_shape[:] = SHAPE_LIST
    t.part_in_prev[:] = 1
    t.node_removed[removed] = 1
    t.node_added[added] = 1
    _apply_weights(cfg, c, t, seed_offset)
    return t


def to_dicts(t, cfg=None, levels=None, rules_by_state=None):
    """Flat tables -> the string API's arguments (small sizes only): names as in the
    module docstring.  Returns kwargs for the literal oracle / _host.PlanNextMapEx."""
    states = ["primary", "replica", "standby"][:t.n_states]
    nodes = ["n%04d" % i for i in range(t.n_nodes)]

    def pmap(rows, shape, present):
        m = {}
        for p in range(t.n_parts):
            if not present[p]:
                continue
            nbs = {}
            for s in range(t.n_states):
                if shape[p, s] == 0:
                    continue
                lo, hi = int(t.state_slot_off[s]), int(t.state_slot_off[s + 1])
                nbs[states[s]] = None if shape[p, s] == 1 else [nodes[x] for x in rows[p, lo:hi] if x != NO_NODE]
            m[str(p)] = nbs
        return m

    kw = dict(
        prev_map=pmap(np.asarray(t.prev_rows).reshape(t.n_parts, -1), np.asarray(t.prev_shape).reshape(t.n_parts, -1), t.part_in_prev),
        partitions_to_assign=pmap(np.asarray(t.cur_rows).reshape(t.n_parts, -1), np.asarray(t.cur_shape).reshape(t.n_parts, -1), t.part_in_assign),
        nodes_all=nodes,
        nodes_to_remove=[nodes[i] for i in np.nonzero(t.node_removed)[0]],
        nodes_to_add=None if t.add_is_nil else [nodes[i] for i in np.nonzero(t.node_added)[0]],
        model={states[s]: (int(t.state_priority[s]), int(t.state_constraints[s])) for s in range(t.n_states)},
        booster=int(t.booster_kind), max_iterations=int(t.max_iters),
    )
    if t.has_part_weights:
        kw["partition_weights"] = {str(p): int(t.part_weight[p]) for p in range(t.n_parts) if t.part_has_weight[p]}
        kw["state_stickiness"] = {states[s]: int(t.state_stickiness[s]) for s in range(t.n_states) if t.state_has_stickiness[s]}
    if t.has_node_weights:
        kw["node_weights"] = {nodes[i]: int(t.node_weight[i]) for i in range(t.n_nodes) if t.node_has_weight[i]}
    if t.has_hier_rules:
        c = CONFIGS[cfg] if cfg is not None else dict(levels=levels, rules=rules_by_state)
        kw["node_hierarchy"] = node_hierarchy_dict(t.n_nodes, c["levels"])
        kw["hierarchy_rules"] = {states[s]: list(r) for s, r in c["rules"].items()}
    return kw
