"""oracle/fast.c's CalcPartitionMoves (the array form that checks the vectorised GPU kernel at sizes the
goldens do not reach) against the literal restatement of moves.go:41-136 (which the reference's 29 + 9 + 21
goldens pin) on random partitions: up to 4 states of which only a prefix is walked, both traversal orders,
duplicates inside a list, nodes that change state, empty sides.  CPU only."""
import ctypes
import random

import numpy as np
import pytest

from oracle_loader import literal
from test_fast_oracle import FAST

L = literal()
FAST.oracle_fast_calc_partition_moves.argtypes = [ctypes.c_int32] * 3 + [ctypes.c_void_p] * 3 + [ctypes.c_int32] * 2 + [ctypes.c_void_p] * 4
KINDS = ["add", "del", "promote", "demote"]
STATE_NAMES = ["primary", "replica", "standby", "zombie"]


def random_case(rnd):
    S = rnd.randint(1, 4)
    names = STATE_NAMES[:S]
    n_visit = rnd.randint(1, S)
    caps = [rnd.randint(0, 3) for _ in range(S)]
    nodes = ["n%d" % i for i in range(rnd.randint(1, 7))]

    def side():
        avail = nodes[:]
        rnd.shuffle(avail)
        m = {}
        for s in range(S):
            if rnd.random() < 0.15:
                continue                                   # state absent on this side
            cnt = rnd.randint(0, min(caps[s], len(avail)))
            lst = [avail.pop() for _ in range(cnt)]
            if lst and caps[s] > len(lst) and rnd.random() < 0.1:
                lst.append(lst[0])                         # a duplicate inside a list
            m[names[s]] = lst
        return m

    return names, n_visit, caps, nodes, side(), side(), rnd.random() < 0.5


def run_fast(names, n_visit, caps, nodes, beg, end, favor):
    S = len(names)
    slot_off = np.concatenate([[0], np.cumsum([max(c, 1) for c in caps])]).astype(np.int32)
    SL = int(slot_off[-1])
    ids = {n: i for i, n in enumerate(nodes)}

    def rows(m):
        r = np.full((1, SL), -1, np.int32)
        for s, name in enumerate(names):
            for j, n in enumerate(m.get(name, [])):
                r[0, slot_off[s] + j] = ids[n]
        return r

    b, e = rows(beg), rows(end)
    max_ops = 2 * SL + 2
    on = np.zeros((1, max_ops), np.int32); os_ = np.zeros((1, max_ops), np.uint8)
    ok = np.zeros((1, max_ops), np.uint8); oc = np.zeros(1, np.int32)
    assert FAST.oracle_fast_calc_partition_moves(1, S, n_visit, slot_off.ctypes.data, b.ctypes.data, e.ctypes.data,
                                                 int(favor), max_ops, on.ctypes.data, os_.ctypes.data, ok.ctypes.data,
                                                 oc.ctypes.data) == 0
    return [(nodes[on[0, i]], "" if os_[0, i] == 0xFF else names[os_[0, i]], KINDS[ok[0, i]]) for i in range(int(oc[0]))]


@pytest.mark.parametrize("chunk", range(8))
def test_fast_moves_equal_literal_on_random_partitions(chunk):
    rnd = random.Random(1000 + chunk)
    for _ in range(1500):
        names, n_visit, caps, nodes, beg, end, favor = random_case(rnd)
        # lists longer than the slot range cannot be expressed in rows: the generator never makes them
        want = [tuple(m) for m in L.calc_partition_moves(names[:n_visit], beg, end, favor)]
        got = run_fast(names, n_visit, caps, nodes, beg, end, favor)
        assert got == want, (names[:n_visit], beg, end, favor)


def test_moves_available_oracle_matches_the_go_statements():
    """oracle_fast_moves_available against a direct Python reading of findAvailableMovesUnlocked
    (orchestrate.go:749-763) + LowestWeightPartitionMoveForNode (orchestrate.go:177-194) with partitions walked
    in ascending index (the one order the reference leaves to Go's map iteration)."""
    import ctypes
    import numpy as np
    FAST.oracle_fast_moves_available.argtypes = [ctypes.c_int32] * 2 + [ctypes.c_void_p] * 7
    weight = {"promote": 1, "demote": 2, "add": 3, "del": 4}
    kinds = ["add", "del", "promote", "demote"]          # enum blance_op_kind
    rng = np.random.default_rng(3)
    for trial in range(200):
        P, NN = int(rng.integers(0, 30)), int(rng.integers(1, 8))
        lens = rng.integers(0, 5, P)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        tot = int(off[-1])
        node = rng.integers(0, NN, max(tot, 1)).astype(np.int32)
        kind = rng.integers(0, 4, max(tot, 1)).astype(np.uint8)
        nxt = rng.integers(0, 6, max(P, 1)).astype(np.int32)
        available = {}
        for p in range(P):                                # for _, nextMoves := range o.mapPartitionToNextMoves
            if nxt[p] < lens[p]:                          # if nextMoves.Next < len(nextMoves.Moves)
                available.setdefault(int(node[off[p] + nxt[p]]), []).append(p)
        best = {}
        for n, moves in available.items():
            r = 0
            for i, p in enumerate(moves):                 # if MoveOpWeight[moves[r].Op] > MoveOpWeight[move.Op] { r = i }
                if weight[kinds[kind[off[moves[r]] + nxt[moves[r]]]]] > weight[kinds[kind[off[p] + nxt[p]]]]:
                    r = i
            best[n] = moves[r]
        r_off = np.zeros(NN + 1, np.int32); r_parts = np.zeros(max(P, 1), np.int32); r_best = np.zeros(NN, np.int32)
        assert FAST.oracle_fast_moves_available(P, NN, off.ctypes.data, node.ctypes.data, kind.ctypes.data, nxt.ctypes.data,
                                                r_off.ctypes.data, r_parts.ctypes.data, r_best.ctypes.data) == 0
        for n in range(NN):
            assert list(r_parts[r_off[n]:r_off[n + 1]]) == available.get(n, []), (trial, n)
            assert r_best[n] == best.get(n, -1), (trial, n)
