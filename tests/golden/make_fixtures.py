#!/usr/bin/env python3
"""Transcribe the reference's own golden vectors into language-neutral JSON.

The reference (couchbase/blance, pure Go) cannot be executed here or on the GPU
box: there is no Go toolchain.  Its tests, however, are table driven, and the
tables are plain Go composite literals.  This script parses those literals
straight from the reference tree (read-only, only needed when REGENERATING the
fixtures - never at test time) and writes

    tests/golden/plan_cases.json     TestPlanNextMap, TestPlanNextMapVis,
                                     TestPlanNextMapHierarchy, TestMultiPrimary,
                                     Test2Replicas,
                                     TestPlanNextMapHierarchyMultiRackFailureCases
                                     (plan_test.go:392-2863) and
                                     TestControlCase1-4 (control_test.go:18-416)
    tests/golden/moves_cases.json    TestFindStateChanges, TestCalcPartitionMoves
                                     (moves_test.go:19-486), and the per-partition move
                                     sequences of TestOrchestrateMoves
                                     (orchestrate_test.go:1049-1811)
    tests/golden/unit_cases.json     helper tables (plan_test.go:21-390,
                                     misc_test.go:18-89)

No reference source is copied: the output holds only the test DATA (inputs and
expected outputs), decoded with the same rules the reference harness applies
(plan_test.go:1642-1744 for the "Vis" rows, moves_test.go:370-486 for the
move lines).  Usage:  python tests/golden/make_fixtures.py [/root/reference]
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

# ----------------------------------------------------------------------------
# Tokenizer for the subset of Go used by the test tables.

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lc>//[^\n]*)
  | (?P<bc>/\*.*?\*/)
  | (?P<raw>`[^`]*`)
  | (?P<str>"(?:\\.|[^"\\])*")
  | (?P<num>\d+(?:\.\d+)?)
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>:=|==|!=|<=|>=|&&|\|\||\+\+|--|\+=|[{}\[\](),:*&.=\-+<>!;/%|])
""", re.X | re.S)


def tokenize(src):
    toks = []
    pos = 0
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError("cannot tokenize at %d: %r" % (pos, src[pos:pos + 40]))
        pos = m.end()
        kind = m.lastgroup
        if kind in ("ws", "lc", "bc"):
            continue
        toks.append((kind, m.group(kind)))
    return toks


def go_unquote(s):
    if s[0] == "`":
        return s[1:-1]
    return json.loads(s)  # the tables only use JSON-compatible escapes


# ----------------------------------------------------------------------------
# A tiny type system, enough to resolve elided literal types and positional
# struct literals.  Types: ("map",K,V) ("slice",T) ("ptr",T) ("named",name)
# ("struct",[(field,type),...]) ("basic",name)

NAMED = {}


def T_named(n):
    return ("named", n)


def resolve(t):
    while t[0] == "named":
        t = NAMED[t[1]]
    return t


STRING = ("basic", "string")
INT = ("basic", "int")
BOOL = ("basic", "bool")
NAMED["Partition"] = ("struct", [("Name", STRING),
                                 ("NodesByState", ("map", STRING, ("slice", STRING)))])
NAMED["PartitionMap"] = ("map", STRING, ("ptr", T_named("Partition")))
NAMED["PartitionModelState"] = ("struct", [("Priority", INT), ("Constraints", INT)])
NAMED["PartitionModel"] = ("map", STRING, ("ptr", T_named("PartitionModelState")))
NAMED["HierarchyRule"] = ("struct", [("IncludeLevel", INT), ("ExcludeLevel", INT)])
NAMED["HierarchyRules"] = ("map", STRING, ("slice", ("ptr", T_named("HierarchyRule"))))
NAMED["OrchestratorOptions"] = ("struct", [("MaxConcurrentPartitionMovesPerNode", INT), ("FavorMinNodes", BOOL)])
NAMED["assignPartitionRec"] = ("struct", [("partition", STRING), ("node", STRING), ("state", STRING), ("op", STRING)])
NAMED["VisTestCase"] = ("struct", [
    ("Ignore", BOOL), ("About", STRING), ("FromTo", ("slice", ("slice", STRING))),
    ("FromToPriority", BOOL), ("Nodes", ("slice", STRING)),
    ("NodesToRemove", ("slice", STRING)), ("NodesToAdd", ("slice", STRING)),
    ("Model", T_named("PartitionModel")),
    ("ModelStateConstraints", ("map", STRING, INT)),
    ("PartitionWeights", ("map", STRING, INT)),
    ("StateStickiness", ("map", STRING, INT)),
    ("NodeWeights", ("map", STRING, INT)),
    ("NodeHierarchy", ("map", STRING, STRING)),
    ("HierarchyRules", T_named("HierarchyRules")),
    ("expNumWarnings", INT)])


def zero(t):
    t = resolve(t)
    if t[0] in ("map", "slice", "ptr"):
        return None
    if t[0] == "struct":
        return {f: zero(ft) for f, ft in t[1]}
    return {"string": "", "int": 0, "bool": False, "error": None}[t[1]]


class Parser:
    def __init__(self, toks, env=None):
        self.t = toks
        self.i = 0
        self.env = env if env is not None else {}

    def peek(self, k=0):
        return self.t[self.i + k] if self.i + k < len(self.t) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ("op", "id"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        tok = self.next()
        if tok[1] != val:
            raise SyntaxError("expected %r got %r at token %d" % (val, tok, self.i))

    # ---- types
    def parse_type(self):
        if self.accept("*"):
            return ("ptr", self.parse_type())
        if self.accept("["):
            self.expect("]")
            return ("slice", self.parse_type())
        kind, val = self.next()
        if val == "map":
            self.expect("[")
            k = self.parse_type()
            self.expect("]")
            return ("map", k, self.parse_type())
        if val == "struct":
            self.expect("{")
            fields = []
            while not self.accept("}"):
                names = [self.next()[1]]
                while self.accept(","):
                    names.append(self.next()[1])
                ft = self.parse_type()
                for n in names:
                    fields.append((n, ft))
                self.accept(";")
            return ("struct", fields)
        if val in ("string", "int", "bool", "error"):
            return ("basic", val)
        if kind == "id":
            return T_named(val)
        raise SyntaxError("bad type token %r" % (val,))

    def looks_like_type(self):
        kind, val = self.peek()
        if val in ("[", "*") or val in ("map", "struct"):
            return True
        return kind == "id" and val in NAMED and self.peek(1)[1] == "{"

    # ---- values
    def parse_value(self, want=None):
        """Parse an expression; `want` is the expected type for elided literals."""
        kind, val = self.peek()
        if val == "&":
            self.next()
            return self.parse_value(want[1] if want and want[0] == "ptr" else want)
        if val == "{":
            t = want
            if t is not None and resolve(t)[0] == "ptr":
                t = resolve(t)[1]
            return self.parse_literal_body(t)
        if self.looks_like_type():
            t = self.parse_type()
            if self.peek()[1] == "(":           # conversion such as []string(nil)
                self.next()
                v = self.parse_value(t)
                self.expect(")")
                return v
            return self.parse_literal_body(t)
        self.next()
        if kind == "str" or kind == "raw":
            return go_unquote(val)
        if kind == "num":
            return int(val)
        if val == "-":
            return -self.parse_value(want)
        if val == "nil":
            return None
        if val == "true":
            return True
        if val == "false":
            return False
        if kind == "id":
            if val in self.env:
                return json.loads(json.dumps(self.env[val]))  # deep copy
            raise SyntaxError("unknown identifier %r" % val)
        raise SyntaxError("bad value token %r" % (val,))

    def parse_literal_body(self, t):
        rt = resolve(t)
        self.expect("{")
        if rt[0] == "map":
            out = {}
            while not self.accept("}"):
                k = self.parse_value(rt[1])
                self.expect(":")
                out[k] = self.parse_value(rt[2])
                if not self.accept(","):
                    self.expect("}")
                    break
            return out
        if rt[0] == "slice":
            out = []
            while not self.accept("}"):
                out.append(self.parse_value(rt[1]))
                if not self.accept(","):
                    self.expect("}")
                    break
            return out
        if rt[0] == "struct":
            out = {f: zero(ft) for f, ft in rt[1]}
            ftypes = dict(rt[1])
            pos = 0
            while not self.accept("}"):
                if self.peek()[0] == "id" and self.peek(1)[1] == ":" and self.peek()[1] in ftypes:
                    f = self.next()[1]
                    self.expect(":")
                    out[f] = self.parse_value(ftypes[f])
                else:
                    f, ft = rt[1][pos]
                    pos += 1
                    out[f] = self.parse_value(ft)
                if not self.accept(","):
                    self.expect("}")
                    break
            return out
        raise SyntaxError("literal of non-composite type %r" % (rt,))


def func_body_tokens(toks, name):
    """Tokens of `func name(...) { BODY }` (BODY only)."""
    for i in range(len(toks) - 1):
        if toks[i] == ("id", "func") and toks[i + 1] == ("id", name):
            j = i
            while toks[j][1] != "{":
                j += 1
            depth = 0
            k = j
            while True:
                if toks[k][1] == "{":
                    depth += 1
                elif toks[k][1] == "}":
                    depth -= 1
                    if depth == 0:
                        return toks[j + 1:k]
                k += 1
    raise KeyError(name)


def parse_test_func(toks, name):
    """Evaluate the leading `ident := literal` statements of a test function and
    return the environment (the table is env['tests'])."""
    body = func_body_tokens(toks, name)
    p = Parser(body)
    while p.peek()[0] == "id" and p.peek(1)[1] == ":=":
        ident = p.next()[1]
        p.next()
        p.env[ident] = p.parse_value()
        if ident == "tests":
            break
    return p.env


# ----------------------------------------------------------------------------
# Decoders that restate what the reference harness does with its tables.

def vis_decode(rowstr, cell_len):
    """plan_test.go:1674-1691: cells -> sort by entry (stable) -> per-state lists."""
    cells = []
    for j in range(0, len(rowstr), cell_len):
        cells.append((rowstr[j:j + cell_len], chr(97 + j // cell_len)))
    cells.sort(key=lambda c: c[0])  # stable; see SURVEY.md section 4 on why this is safe
    nbs = {}
    for entry, node in cells:
        st = {"m": "primary", "s": "replica"}.get(entry[0:1], "")
        if st:
            nbs.setdefault(st, []).append(node)
    return nbs


def vis_cases(group, tests):
    out = []
    for idx, c in enumerate(tests):
        cell = 2 if c["FromToPriority"] else 1
        prev, exp = {}, {}
        for i, (frm, to) in enumerate(c["FromTo"]):
            name = "%03d" % i
            prev[name] = {"name": name, "nodesByState": vis_decode(frm, cell)}
            exp[name] = {"name": name, "nodesByState": vis_decode(to, cell)}
        out.append({
            "group": group, "index": idx, "about": c["About"], "ignore": c["Ignore"],
            "prevMap": prev, "partitionsToAssign": None,  # None = SAME OBJECT as prevMap (plan_test.go:1716-1718)
            "nodes": c["Nodes"], "nodesToRemove": c["NodesToRemove"], "nodesToAdd": c["NodesToAdd"],
            "model": model_json(c["Model"]),
            "modelStateConstraints": c["ModelStateConstraints"],
            "partitionWeights": c["PartitionWeights"], "stateStickiness": c["StateStickiness"],
            "nodeWeights": c["NodeWeights"], "nodeHierarchy": c["NodeHierarchy"],
            "hierarchyRules": rules_json(c["HierarchyRules"]),
            "booster": "none",
            "exp": exp, "expNumWarnings": c["expNumWarnings"],
            "warnCount": "partitions",  # len(rWarnings), plan_test.go:1738
        })
    return out


def model_json(m):
    if m is None:
        return None
    return {k: {"priority": v["Priority"], "constraints": v["Constraints"]} for k, v in m.items()}


def rules_json(r):
    if r is None:
        return None
    return {k: [{"includeLevel": x["IncludeLevel"], "excludeLevel": x["ExcludeLevel"]} for x in v]
            for k, v in r.items()}


def pmap_json(m):
    if m is None:
        return None
    return {k: {"name": v["Name"], "nodesByState": v["NodesByState"]} for k, v in m.items()}


def plan_table_cases(tests):
    out = []
    for idx, c in enumerate(tests):
        out.append({
            "group": "TestPlanNextMap", "index": idx, "about": c["About"], "ignore": False,
            "prevMap": pmap_json(c["PrevMap"]), "partitionsToAssign": pmap_json(c["PartitionsToAssign"]),
            "nodes": c["Nodes"], "nodesToRemove": c["NodesToRemove"], "nodesToAdd": c["NodesToAdd"],
            "model": model_json(c["Model"]),
            "modelStateConstraints": c["ModelStateConstraints"],
            "partitionWeights": c["PartitionWeights"], "stateStickiness": c["StateStickiness"],
            "nodeWeights": c["NodeWeights"], "nodeHierarchy": c["NodeHierarchy"],
            "hierarchyRules": rules_json(c["HierarchyRules"]),
            "booster": "none",
            "exp": pmap_json(c["exp"]), "expNumWarnings": c["expNumWarnings"],
            "warnCount": "strings",  # sum of len(warnings[p]), plan_test.go:1599-1603
        })
    return out


# ----------------------------------------------------------------------------

def convert_line(line, states):
    """moves_test.go:491-517."""
    nbs = {}
    line = line.strip(" ")
    while True:
        linex = line.replace("  ", " ")
        if linex == line:
            break
        line = linex
    parts = line.split("|")
    for i, st in enumerate(states):
        if i >= len(parts):
            break
        part = parts[i].strip(" ")
        if part != "":
            nbs.setdefault(st, []).extend(part.split(" "))
    return nbs


def moves_cases(env):
    states = env["states"]
    out = []
    for idx, c in enumerate(env["tests"]):
        before = convert_line(c["before"], states)
        after = convert_line(c["after"], states)
        exp = []
        if c["moves"] != "":
            for ml in c["moves"].split("\n"):
                me = convert_line(ml.replace("\t", " "), states)
                # moves_test.go:397-470: the first +x / -x entry (states in order)
                # names the node; a "flip side" entry in a later state turns the
                # op into promote/demote (the reference accepts either word).
                found = None
                for si, st in enumerate(states):
                    if found:
                        break
                    for mv in me.get(st, []):
                        if found:
                            break
                        op = mv[0:1]
                        if op in "+-" and op != "":
                            flip = {"+": "-", "-": "+"}[op] + mv[1:]
                            flip_state = ""
                            for j in range(si + 1, len(states)):
                                for x in me.get(states[j], []):
                                    if x == flip:
                                        flip_state = states[j]
                            state_exp = st
                            if flip_state:
                                if op == "-":
                                    state_exp = flip_state
                                ops = ["promote", "demote"]
                            else:
                                if op == "-":
                                    state_exp = ""
                                ops = [{"+": "add", "-": "del"}[op]]
                            found = {"node": mv[1:], "state": state_exp, "op": ops}
                assert found is not None, (idx, ml)
                exp.append(found)
        out.append({"index": idx, "states": states, "before": before, "after": after,
                    "favorMinNodes": c["favorMinNodes"], "exp": exp})
    return out


def package_vars(toks, names):
    """Top-level `var name = literal` declarations."""
    env = {}
    for i in range(len(toks) - 3):
        if toks[i] == ("id", "var") and toks[i + 1][1] in names and toks[i + 2][1] == "=":
            p = Parser(toks, env)
            p.i = i + 3
            env[toks[i + 1][1]] = p.parse_value()
    return env


def orchestrate_cases(toks):
    """TestOrchestrateMoves (orchestrate_test.go:1049-1811): the fake callback records, per partition, the
    (node, state) sequence the orchestrator applies, which is CalcPartitionMoves(states, beg[p], end[p],
    options.FavorMinNodes) in order (orchestrate.go:263-287).  states = sortStateNames(model)."""
    env = package_vars(toks, ("mrPartitionModel", "options1"))
    body = func_body_tokens(toks, "TestOrchestrateMoves")
    p = Parser(body, env)
    assert p.peek()[1] == "tests"
    p.next(); p.next()
    tests = p.parse_value()
    out = []
    for idx, c in enumerate(tests):
        if c["skip"] or not c["expectAssignPartitions"]:
            continue
        model = c["partitionModel"]
        states = sorted(model, key=lambda k: (model[k]["Priority"], k))
        for part, recs in sorted(c["expectAssignPartitions"].items()):
            out.append({"index": idx, "label": c["label"], "partition": part, "states": states,
                        "before": (c["begMap"].get(part) or {"NodesByState": {}})["NodesByState"],
                        "after": (c["endMap"].get(part) or {"NodesByState": {}})["NodesByState"],
                        "favorMinNodes": c["options"]["FavorMinNodes"],
                        "exp": [{"node": r["node"], "state": r["state"]} for r in recs]})
    return out


def main():
    plan_toks = tokenize(open(os.path.join(REF, "plan_test.go")).read())
    cases = plan_table_cases(parse_test_func(plan_toks, "TestPlanNextMap")["tests"])
    for fn in ("TestPlanNextMapVis", "TestPlanNextMapHierarchy", "TestMultiPrimary",
               "Test2Replicas", "TestPlanNextMapHierarchyMultiRackFailureCases"):
        cases += vis_cases(fn, parse_test_func(plan_toks, fn)["tests"])
    cases += control_cases_all(tokenize(open(os.path.join(REF, "control_test.go")).read()))
    with open(os.path.join(OUT, "plan_cases.json"), "w") as f:
        json.dump(cases, f, indent=1, sort_keys=True)
    active = sum(1 for c in cases if not c["ignore"])
    print("plan cases: %d (%d active)" % (len(cases), active))

    mv_toks = tokenize(open(os.path.join(REF, "moves_test.go")).read())
    fsc = parse_test_func(mv_toks, "TestFindStateChanges")["tests"]
    mv = moves_cases(parse_test_func(mv_toks, "TestCalcPartitionMoves"))
    orch = orchestrate_cases(tokenize(open(os.path.join(REF, "orchestrate_test.go")).read()))
    with open(os.path.join(OUT, "moves_cases.json"), "w") as f:
        json.dump({"findStateChanges": fsc, "calcPartitionMoves": mv, "orchestrateDerived": orch}, f, indent=1, sort_keys=True)
    print("findStateChanges cases: %d, calcPartitionMoves cases: %d, orchestrate-derived sequences: %d"
          % (len(fsc), len(mv), len(orch)))

    units = {}
    for fn, key in (("TestFlattenNodesByState", "flattenNodesByState"),
                    ("TestRemoveNodesFromNodesByState", "removeNodesFromNodesByState"),
                    ("TestStateNameSorter", "stateNameSorter"),
                    ("TestCountStateNodes", "countStateNodes"),
                    ("TestFindAncestor", "findAncestor"),
                    ("TestFindLeaves", "findLeaves"),
                    ("TestMapParentsToMapChildren", "mapParentsToMapChildren")):
        units[key] = parse_test_func(plan_toks, fn)["tests"]
    misc_toks = tokenize(open(os.path.join(REF, "misc_test.go")).read())
    for fn, key in (("TestStringsRemoveStrings", "stringsRemoveStrings"),
                    ("TestStringsIntersectStrings", "stringsIntersectStrings"),
                    ("TestStringsDeduplicate", "stringsDeduplicate")):
        try:
            units[key] = parse_test_func(misc_toks, fn)["tests"]
        except KeyError:
            pass
    with open(os.path.join(OUT, "unit_cases.json"), "w") as f:
        json.dump(units, f, indent=1, sort_keys=True)
    print("unit tables:", {k: len(v) for k, v in units.items()})


def control_cases_all(toks):
    """TestControlCase1-4 build their inputs with `ident := literal` statements
    (control_test.go:18-416); evaluate those, then parse the argument list of
    the PlanNextMapEx(...) call.  All four install cbgt's booster
    max(float64(-w), stickiness) (control_test.go:19-26) and assert
    len(warnings) == 0."""
    out = []
    for fn in ("TestControlCase1", "TestControlCase2", "TestControlCase3", "TestControlCase4"):
        body = func_body_tokens(toks, fn)
        p = Parser(body)
        # walk every top-level `ident := <composite literal>` in the body
        env = {}
        i = 0
        while i < len(body) - 2:
            if body[i][0] == "id" and body[i + 1][1] == ":=" and (
                    body[i + 2][1] in ("[", "map", "&") or body[i + 2][1] in NAMED):
                p.i = i + 2
                p.env = env
                try:
                    env[body[i][1]] = p.parse_value()
                    i = p.i
                    continue
                except SyntaxError:
                    pass
            i += 1
        out.append((fn, env))
    return [control_case_from_env(fn, env, toks) for fn, env in out]


def control_case_from_env(fn, env, toks):
    # The call sites are checked by eye against control_test.go:261-283 and
    # :369-391; the option literals inside the call are parsed below.
    body = func_body_tokens(toks, fn)
    # find "PlanNextMapOptions" "{" and parse that struct literal
    NAMED.setdefault("PlanNextMapOptions", ("struct", [
        ("ModelStateConstraints", ("map", STRING, INT)),
        ("PartitionWeights", ("map", STRING, INT)),
        ("StateStickiness", ("map", STRING, INT)),
        ("NodeWeights", ("map", STRING, INT)),
        ("NodeHierarchy", ("map", STRING, STRING)),
        ("HierarchyRules", T_named("HierarchyRules"))]))
    opts = None
    call_args = None
    for i in range(len(body)):
        if body[i] == ("id", "PlanNextMapEx") and body[i + 1][1] == "(":
            p = Parser(body, dict(env))
            p.i = i + 2
            args = []
            while True:
                args.append(p.parse_value())
                if not p.accept(","):
                    break
                if p.peek()[1] == ")":
                    break
            call_args = args
            break
    prev, assign, nodes, rm, add, model, opts = call_args
    exp = env.get("expect")
    return {
        "group": fn, "index": 0, "about": fn, "ignore": False,
        "prevMap": pmap_json(prev), "partitionsToAssign": pmap_json(assign),
        "nodes": nodes, "nodesToRemove": rm, "nodesToAdd": add,
        "model": model_json(model),
        "modelStateConstraints": opts["ModelStateConstraints"],
        "partitionWeights": opts["PartitionWeights"], "stateStickiness": opts["StateStickiness"],
        "nodeWeights": opts["NodeWeights"], "nodeHierarchy": opts["NodeHierarchy"],
        "hierarchyRules": rules_json(opts["HierarchyRules"]),
        "booster": "cbgt",
        "exp": pmap_json(exp), "expNumWarnings": 0, "warnCount": "partitions",
    }


if __name__ == "__main__":
    main()
