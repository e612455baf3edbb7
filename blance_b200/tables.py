"""Flat planner tables in numpy, laid out exactly as struct blance_plan_in /
blance_plan_out of include/blance_b200.h, for callers that already hold ids
instead of strings (bench.py, the large parity tests).  The arrays are plain host
memory; PlanTables.struct() returns the ctypes struct whose pointers alias them."""
import ctypes

import numpy as np

from . import abi as api      # ctypes structs + lazy library handle: importing tables loads no native code

NO_NODE = -1
SHAPE_ABSENT, SHAPE_NIL, SHAPE_LIST = 0, 1, 2

_I32 = ("state_priority", "state_constraints", "state_slot_off", "state_stickiness", "node_weight", "part_weight",
        "part_name_rank", "prev_rows", "cur_rows", "extra_tot_first", "extra_tot_rest", "rule_off")
_U8 = ("state_has_stickiness", "node_removed", "node_added", "node_has_weight", "part_in_prev", "part_in_assign",
       "part_has_weight", "prev_shape", "cur_shape")


class PlanTables:
    """Holds every array of a blance_plan_in.  Scalars are attributes; arrays are
    C-contiguous numpy arrays of the ABI's dtypes."""

    def __init__(self, n_nodes, n_states, n_parts, state_priority, state_constraints, n_node_ids=None):
        self.n_nodes = int(n_nodes)
        self.n_node_ids = int(n_node_ids if n_node_ids is not None else n_nodes)
        self.n_states = int(n_states)
        self.n_parts = int(n_parts)
        self.max_iters = 10
        self.booster_kind = 0
        self.add_is_nil = 0
        self.has_part_weights = 0
        self.has_node_weights = 0
        self.has_hier_rules = 0
        self.engine = 0
        self.n_rules = 0
        self.n_hier_bits = self.n_nodes
        S, P, N, NU = self.n_states, self.n_parts, self.n_nodes, self.n_node_ids
        self.state_priority = np.asarray(state_priority, np.int32)
        self.state_constraints = np.asarray(state_constraints, np.int32)
        caps = np.maximum(self.state_constraints, 0)
        self.state_slot_off = np.concatenate([[0], np.cumsum(caps)]).astype(np.int32)
        self.n_slots = int(self.state_slot_off[-1])
        self.top_state = int(np.argmin(self.state_priority)) if S else 0
        self.state_stickiness = np.zeros(S, np.int32)
        self.state_has_stickiness = np.zeros(S, np.uint8)
        self.node_removed = np.zeros(NU, np.uint8)
        self.node_added = np.zeros(NU, np.uint8)
        self.node_weight = np.zeros(N, np.int32)
        self.node_has_weight = np.zeros(N, np.uint8)
        self.part_in_prev = np.zeros(P, np.uint8)
        self.part_in_assign = np.ones(P, np.uint8)
        self.part_weight = np.ones(P, np.int32)
        self.part_has_weight = np.zeros(P, np.uint8)
        self.part_name_rank = np.arange(P, dtype=np.int32)
        self.prev_rows = np.full((P, self.n_slots), NO_NODE, np.int32)
        self.cur_rows = np.full((P, self.n_slots), NO_NODE, np.int32)
        self.prev_shape = np.zeros((P, S), np.uint8)
        self.cur_shape = np.zeros((P, S), np.uint8)
        self.extra_tot_first = np.zeros(N, np.int32)
        self.extra_tot_rest = np.zeros(N, np.int32)
        self.rule_off = np.zeros(S + 1, np.int32)
        self.ie_mask = np.zeros(0, np.uint32)

    @property
    def hier_words(self):
        return (self.n_hier_bits + 31) // 32

    def struct(self):
        s = api.PlanIn()
        for f in api._I32_FIELDS:
            setattr(s, f, int(getattr(self, f)))
        s.n_rules, s.n_hier_bits, s.engine = int(self.n_rules), int(self.n_hier_bits), int(self.engine)
        self._keep = []
        for f in api._PTR_FIELDS + ("rule_off", "ie_mask"):
            want = np.uint32 if f == "ie_mask" else (np.int32 if f in _I32 else np.uint8)
            a = np.ascontiguousarray(getattr(self, f), dtype=want)
            setattr(self, f, a)
            self._keep.append(a)
            setattr(s, f, a.ctypes.data if a.size else None)
        return s


class PlanResult:
    """Output buffers of a blance_plan_out."""

    def __init__(self, t):
        self.next_rows = np.full((t.n_parts, t.n_slots), NO_NODE, np.int32)
        self.next_shape = np.zeros((t.n_parts, t.n_states), np.uint8)
        self.warn = np.zeros((t.n_parts, t.n_states), np.uint8)
        # one spare element so zero-sized cases still have valid pointers
        self._pad = np.zeros(4, np.int32)
        self.out = api.PlanOut()
        self.out.next_rows = self.next_rows.ctypes.data if self.next_rows.size else self._pad.ctypes.data
        self.out.next_shape = self.next_shape.ctypes.data if self.next_shape.size else self._pad.ctypes.data
        self.out.warn = self.warn.ctypes.data if self.warn.size else self._pad.ctypes.data

    iters_run = property(lambda self: self.out.iters_run)
    converged = property(lambda self: self.out.converged)
    steps = property(lambda self: self.out.steps)
    sticky_steps = property(lambda self: self.out.sticky_steps)
    device_ms = property(lambda self: self.out.device_ms)
    kernel_ms = property(lambda self: self.out.kernel_ms)
    pass_ms = property(lambda self: self.out.pass_ms)


class Context:
    """A blance_ctx* (one per process/GPU)."""

    def __init__(self, device_id=-1, device_ids=None):
        """device_ids (a list) creates a multi-GPU context: batches are sharded over those devices."""
        self.lib = api.capi()
        self.ptr = ctypes.c_void_p()
        if device_ids is not None:
            arr = (ctypes.c_int * len(device_ids))(*device_ids)
            st = self.lib.blance_ctx_create_multi(ctypes.byref(self.ptr), arr, len(device_ids))
        else:
            st = self.lib.blance_ctx_create(ctypes.byref(self.ptr), device_id)
        if st != 0:
            raise api.BlanceError("blance_ctx_create failed (%d): %s" % (st, self.lib.blance_last_error(None).decode()))

    def device_count(self):
        return int(self.lib.blance_ctx_device_count(self.ptr))

    def _check(self, st, what):
        if st != 0:
            raise api.BlanceError("%s failed (%d): %s" % (what, st, self.lib.blance_last_error(self.ptr).decode()))

    def plan_next_map(self, tables, result=None):
        """blance_plan_next_map: host buffers in, host buffers out."""
        result = result or PlanResult(tables)
        s = tables.struct()
        self._check(self.lib.blance_plan_next_map(self.ptr, ctypes.byref(s), ctypes.byref(result.out)), "blance_plan_next_map")
        return result

    def plan_next_map_batch(self, tables_list, results=None):
        n = len(tables_list)
        results = results or [PlanResult(t) for t in tables_list]
        ins = (api.PlanIn * n)(*[t.struct() for t in tables_list])
        outs = (api.PlanOut * n)(*[r.out for r in results])
        self._check(self.lib.blance_plan_next_map_batch(self.ptr, n, ins, outs), "blance_plan_next_map_batch")
        for r, o in zip(results, outs):
            r.out = o
        return results

    def prepare_batch(self, tables_list, results=None):
        """Builds the blance_plan_in / blance_plan_out arrays of a batch once; run_batch() is then only the
        C call (what a compiled host would do per request)."""
        n = len(tables_list)
        results = results or [PlanResult(t) for t in tables_list]
        ins = (api.PlanIn * n)(*[t.struct() for t in tables_list])
        outs = (api.PlanOut * n)(*[r.out for r in results])
        return n, ins, outs, results, tables_list

    def run_batch(self, prepared):
        n, ins, outs, results, _ = prepared
        self._check(self.lib.blance_plan_next_map_batch(self.ptr, n, ins, outs), "blance_plan_next_map_batch")
        for r, o in zip(results, outs):
            r.out = o
        return results

    def upload(self, tables):
        plan = ctypes.c_void_p()
        s = tables.struct()
        self._check(self.lib.blance_plan_upload(self.ptr, ctypes.byref(s), ctypes.byref(plan)), "blance_plan_upload")
        return plan

    def run(self, plan):
        self._check(self.lib.blance_plan_run(self.ptr, plan), "blance_plan_run")

    def timing(self, plan):
        """(kernel_ms, pass_ms, pass_launches) of the last run()."""
        k, p, n = ctypes.c_float(), ctypes.c_float(), ctypes.c_int32()
        self._check(self.lib.blance_plan_timing(plan, ctypes.byref(k), ctypes.byref(p), ctypes.byref(n)), "blance_plan_timing")
        return k.value, p.value, n.value

    def fetch(self, plan, result):
        self._check(self.lib.blance_plan_fetch(self.ptr, plan, ctypes.byref(result.out)), "blance_plan_fetch")
        return result

    def free(self, plan):
        self.lib.blance_plan_free(self.ptr, plan)

    def calc_partition_moves(self, slot_off, beg_rows, end_rows, favor_min_nodes, n_visit_states=None):
        slot_off = np.ascontiguousarray(slot_off, np.int32)
        beg = np.ascontiguousarray(beg_rows, np.int32)
        end = np.ascontiguousarray(end_rows, np.int32)
        n_states = len(slot_off) - 1
        n_parts = beg.shape[0]
        max_ops = max(1, 2 * int(slot_off[-1]))
        op_node = np.zeros((n_parts, max_ops), np.int32)
        op_state = np.zeros((n_parts, max_ops), np.uint8)
        op_kind = np.zeros((n_parts, max_ops), np.uint8)
        op_count = np.zeros(n_parts, np.int32)
        st = self.lib.blance_calc_partition_moves(
            self.ptr, n_parts, n_states, n_states if n_visit_states is None else n_visit_states, slot_off.ctypes.data,
            beg.ctypes.data, end.ctypes.data, int(bool(favor_min_nodes)), max_ops, op_node.ctypes.data,
            op_state.ctypes.data, op_kind.ctypes.data, op_count.ctypes.data)
        self._check(st, "blance_calc_partition_moves")
        return op_node, op_state, op_kind, op_count

    def moves_create(self, slot_off, beg_rows, end_rows, favor_min_nodes, n_node_ids, n_visit_states=None):
        """blance_moves_create: CalcPartitionMoves of every partition, resident on the device in CSR form.
        Returns (handle, total_ops)."""
        slot_off = np.ascontiguousarray(slot_off, np.int32)
        beg = np.ascontiguousarray(beg_rows, np.int32)
        end = np.ascontiguousarray(end_rows, np.int32)
        n_states = len(slot_off) - 1
        h, tot = ctypes.c_void_p(), ctypes.c_int64()
        st = self.lib.blance_moves_create(self.ptr, beg.shape[0], n_states, n_states if n_visit_states is None else n_visit_states,
                                          slot_off.ctypes.data, beg.ctypes.data, end.ctypes.data, int(bool(favor_min_nodes)),
                                          int(n_node_ids), ctypes.byref(h), ctypes.byref(tot))
        self._check(st, "blance_moves_create")
        return (h, beg.shape[0], int(n_node_ids)), int(tot.value)

    def moves_fetch(self, handle, total_ops):
        h, n_parts, _ = handle
        off = np.zeros(n_parts + 1, np.int64)
        node = np.zeros(max(1, total_ops), np.int32)
        state = np.zeros(max(1, total_ops), np.uint8)
        kind = np.zeros(max(1, total_ops), np.uint8)
        self._check(self.lib.blance_moves_fetch(self.ptr, h, off.ctypes.data, node.ctypes.data, state.ctypes.data, kind.ctypes.data), "blance_moves_fetch")
        return off, node[:total_ops], state[:total_ops], kind[:total_ops]

    def moves_available(self, handle, next_idx):
        h, n_parts, n_node_ids = handle
        nxt = np.ascontiguousarray(next_idx, np.int32)
        node_off = np.zeros(n_node_ids + 1, np.int32)
        node_parts = np.zeros(max(1, n_parts), np.int32)
        best = np.zeros(max(1, n_node_ids), np.int32)
        self._check(self.lib.blance_moves_available(self.ptr, h, nxt.ctypes.data, node_off.ctypes.data, node_parts.ctypes.data, best.ctypes.data),
                    "blance_moves_available")
        return node_off, node_parts[:node_off[-1]], best[:n_node_ids]

    def moves_free(self, handle):
        self.lib.blance_moves_free(self.ptr, handle[0])

    def kernel_launches(self):
        return int(self.lib.blance_ctx_kernel_launches(self.ptr))

    def close(self):
        if self.ptr:
            self.lib.blance_ctx_destroy(self.ptr)
            self.ptr = ctypes.c_void_p()
