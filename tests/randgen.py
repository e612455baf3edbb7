"""Random small planner instances in the string API's form, covering the option
space the reference's goldens leave unpinned: 3 states, k up
to 3, several rules per state, deeper hierarchies, stickiness with partition
weights, negative node weights with the booster, prevMap != partitionsToAssign."""
import random

STATES = ["primary", "replica", "standby"]


def random_instance(seed, max_nodes=10, max_parts=20):
    rnd = random.Random(seed)
    n_nodes = rnd.randint(1, max_nodes)
    nodes = ["n%02d" % i for i in range(n_nodes)]
    rnd.shuffle(nodes)
    n_states = rnd.randint(1, 3)
    states = STATES[:n_states]
    equal_prio = rnd.random() < 0.1
    model = {s: (0 if equal_prio else i, rnd.choice([0, 1, 1, 1, 2, 2, 3])) for i, s in enumerate(states)}
    n_parts = rnd.randint(0, max_parts)
    numeric = rnd.random() < 0.7
    names = [str(i) if numeric else "p%03d" % i for i in range(n_parts)]
    if rnd.random() < 0.2 and n_parts > 2:
        names[1] = "007"          # exercises the Atoi/padding rule against raw-name tie-breaks
        names[2] = "7"
        names = list(dict.fromkeys(names))

    def rand_row(allow_extra):
        nbs = {}
        pool = nodes + (["ghost"] if allow_extra and rnd.random() < 0.1 else [])
        avail = pool[:]
        rnd.shuffle(avail)
        for s in states:
            r = rnd.random()
            if r < 0.25:
                continue
            cnt = rnd.randint(0, min(3, len(avail)))
            lst = [avail.pop() for _ in range(cnt)]
            if rnd.random() < 0.05 and lst:
                lst.append(lst[0])       # duplicate inside a list (misc.go:45 dedupes the decrement)
            nbs[s] = lst if (lst or rnd.random() < 0.8) else None
        return nbs

    mode = rnd.random()
    prev = {n: rand_row(True) for n in names} if mode > 0.15 else {}
    if mode > 0.5:
        assign = None                      # same object
    elif mode > 0.3:
        assign = {n: {s: (None if v is None else list(v)) for s, v in prev[n].items()} for n in names if rnd.random() < 0.7}
    elif mode > 0.15:
        assign = {n: rand_row(False) for n in names}
    else:
        assign = {n: ({} if rnd.random() < 0.7 else rand_row(False)) for n in names}
    if mode > 0.15 and assign is not None and rnd.random() < 0.3:
        prev["other"] = {"primary": [rnd.choice(nodes)], "dead": [rnd.choice(nodes)]}   # not assigned; non-model state
    remove = rnd.sample(nodes, rnd.randint(0, max(0, n_nodes // 3))) if rnd.random() < 0.6 else []
    if assign is not None and remove and any(n not in prev for n in assign):
        remove = []                        # the reference panics otherwise (plan.go:544)
    if rnd.random() < 0.1:
        remove = None
    r = rnd.random()
    add = None if r < 0.2 else rnd.sample(nodes, rnd.randint(0, n_nodes)) if r < 0.8 else []
    kw = dict(prev_map=prev, partitions_to_assign=assign, nodes_all=nodes, nodes_to_remove=remove, nodes_to_add=add,
              model=model)
    if rnd.random() < 0.3:
        kw["model_state_constraints"] = {rnd.choice(states): rnd.randint(0, 3)}
    if rnd.random() < 0.5:
        kw["partition_weights"] = {n: rnd.randint(0, 9) for n in names if rnd.random() < 0.5}
        if rnd.random() < 0.7:
            kw["state_stickiness"] = {s: rnd.randint(0, 5) for s in states if rnd.random() < 0.7}
    if rnd.random() < 0.5:
        kw["node_weights"] = {n: rnd.choice([-3, -2, -1, 0, 1, 1, 2, 3, 5]) for n in nodes if rnd.random() < 0.7}
        kw["booster"] = 1 if rnd.random() < 0.5 else 0
    if rnd.random() < 0.5:
        depth = rnd.randint(1, 3)
        parents = {}
        cur = nodes[:] + (["spare"] if rnd.random() < 0.3 else [])
        if rnd.random() < 0.3 and len(cur) > 1:
            cur = cur[1:]                   # one node is missing from the hierarchy
        for lv in range(depth):
            groups = max(1, len(cur) // rnd.randint(1, 3))
            nxt = ["g%d_%d" % (lv, g) for g in range(groups)]
            for i, c in enumerate(cur):
                parents[c] = nxt[i % groups]
            cur = nxt
        kw["node_hierarchy"] = parents
        if rnd.random() < 0.85:
            kw["hierarchy_rules"] = {s: [(rnd.randint(0, 3), rnd.randint(0, 3)) for _ in range(rnd.randint(0, 2))]
                                     for s in states if rnd.random() < 0.7}
    # a prevMap entry of an ASSIGNED partition with a key outside the model: reflect.DeepEqual (plan.go:38) can
    # never match it, so the first iteration cannot converge (drawn from a separate stream: older seeds keep
    # their instances)
    rnd2 = random.Random(seed * 7919 + 13)
    if assign is not None and prev and rnd2.random() < 0.25:
        both = [n for n in names if n in prev and n in assign]
        if both:
            n = rnd2.choice(both)
            prev[n] = dict(prev[n])
            prev[n]["dead"] = rnd2.choice([[], None, [rnd2.choice(nodes)]])
    return kw
