/* oracle/fast.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * "Fast" CPU oracle: the array-form restatement of blance's planner
 * (plan.go:23-331) and CalcPartitionMoves (moves.go:41-136) behind the SAME flat
 * tables as the product's C ABI (include/blance_b200.h), in plain C: int32 rows,
 * dense count tables, bit sets for the hierarchy, and an O(N) masked arg-min in
 * place of the reference's O(N log N) comparison sort (the node order
 * (score, position) of plan.go:617-628 is a strict total order, so "sort then take
 * the first k" equals k successive arg-mins).
 *
 * Parity status: PINNED — driven through the host interning layer it reproduces
 * every golden vector of the reference (tests/test_fast_oracle.py) and equals the
 * literal oracle (oracle/literal.cpp) on randomised instances.  It is the only
 * CPU path that can produce the expected PartitionMap of the 1M x 1024 workload.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs
 * may link or call this file.  Build: see oracle/Makefile (-ffp-contract=off: the
 * score of plan.go:634-689 is IEEE binary64, round-to-nearest, never fused).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "blance_b200.h"

#define NONE BLANCE_NO_NODE

typedef struct {
  const blance_plan_in* in;
  int32_t N, NU, S, PU, SL, HW;
  int32_t* rows;        /* working rows [PU][SL] */
  uint8_t* shape;       /* [PU][S] */
  int32_t* prev_rows;
  uint8_t* prev_shape;
  uint8_t* in_prev;
  uint8_t* valid;       /* [N] node still in nodesAll \ nodesToRemove */
  int64_t* counts;      /* [S][N] stateNodeCounts */
  int64_t* tot;         /* [N] nodePartitionCounts */
  int32_t* n2n;         /* [NU+1][N] nodeToNodeCounts, row NU = "" */
  uint8_t* warn;
  int32_t n_valid;
  int64_t P;            /* len(prevMap) */
  int64_t steps;
  /* scratch */
  double* score;
  uint8_t* cand;
  uint32_t* hm;
} fo_t;

static int32_t* row_of(int32_t* rows, int32_t SL, int32_t p) { return rows + (size_t)p * SL; }

static int list_len(const int32_t* row, int32_t lo, int32_t hi) {
  int n = 0;
  while (lo + n < hi && row[lo + n] != NONE) n++;
  return n;
}

static int list_has(const int32_t* row, int32_t lo, int32_t hi, int32_t node) {
  for (int32_t i = lo; i < hi && row[i] != NONE; i++)
    if (row[i] == node) return 1;
  return 0;
}

/* plan.go:634-689, evaluated for one candidate node */
static double node_score(const fo_t* f, int32_t s, int32_t n, int32_t top, double cur) {
  const blance_plan_in* in = f->in;
  double lower = 0.0, filled = 0.0;
  if (f->P > 0) {
    lower = (double)f->n2n[(size_t)top * f->N + n] / (double)f->P;
    filled = (0.001 * (double)f->tot[n]) / (double)f->P;
  }
  double r = (double)f->counts[(size_t)s * f->N + n];
  r = r + lower;
  r = r + filled;
  if (in->has_node_weights && in->node_has_weight[n]) {
    int32_t w = in->node_weight[n];
    if (w > 0) {
      r = r / (double)w;
    } else if (w < 0 && in->booster_kind == BLANCE_BOOSTER_CBGT_MAX) {
      double b = (double)(-(int64_t)w);         /* control_test.go:19-26 */
      if (b < cur) b = cur;
      r = r + b;
    }
  }
  r = r - cur;
  return r;
}

/* arg-min of (score, position) over nodes with mask[n] != 0 (and, when bits is
 * given, bit n of bits set); -1 if none.  plan.go:617-628. */
static int32_t argmin_masked(const fo_t* f, const uint8_t* mask, const uint32_t* bits) {
  int32_t best = -1;
  double bs = 0.0;
  for (int32_t n = 0; n < f->N; n++) {
    if (!mask[n]) continue;
    if (bits && !((bits[n >> 5] >> (n & 31)) & 1u)) continue;
    if (best < 0 || f->score[n] < bs) { best = n; bs = f->score[n]; }
  }
  return best;
}

/* removeNodesFromNodesByState(row, removeList, dec), plan.go:290-297,408-421.
 * Works on a snapshot of each list so that duplicate handling follows
 * StringsIntersectStrings (misc.go:40-51): one decrement per distinct node. */
static void remove_nodes(fo_t* f, int32_t p, const int32_t* rm, int n_rm, int64_t w) {
  const blance_plan_in* in = f->in;
  int32_t* row = row_of(f->rows, f->SL, p);
  int32_t snap[64];
  for (int32_t s = 0; s < f->S; s++) {
    uint8_t* sh = &f->shape[(size_t)p * f->S + s];
    if (*sh == BLANCE_SHAPE_ABSENT) continue;
    int32_t lo = in->state_slot_off[s], hi = in->state_slot_off[s + 1];
    int len = list_len(row, lo, hi);
    int32_t* list = len <= 64 ? snap : (int32_t*)malloc(sizeof(int32_t) * (size_t)len);
    memcpy(list, row + lo, sizeof(int32_t) * (size_t)len);
    int32_t out = lo;
    for (int i = 0; i < len; i++) {
      int32_t node = list[i];
      int hit = 0;
      for (int j = 0; j < n_rm; j++) hit |= (rm[j] == node);
      if (!hit) { row[out++] = node; continue; }
      int dup = 0;
      for (int j = 0; j < i; j++) dup |= (list[j] == node);
      if (!dup && node < f->N) {
        f->counts[(size_t)s * f->N + node] -= w;
        f->tot[node] -= w;
      }
    }
    for (; out < hi; out++) row[out] = NONE;
    *sh = BLANCE_SHAPE_LIST;   /* StringsRemoveStrings always returns a non-nil slice, misc.go:29 */
    if (list != snap) free(list);
  }
}

static void one_step(fo_t* f, int32_t p, int32_t s) {            /* plan.go:268-302 + findBestNodes */
  const blance_plan_in* in = f->in;
  const int32_t N = f->N;
  int32_t* row = row_of(f->rows, f->SL, p);
  const int32_t k = in->state_constraints[s];
  const int32_t lo = in->state_slot_off[s], hi = in->state_slot_off[s + 1];
  f->steps++;

  int64_t w_p = 1;
  double stick = 1.5;                                              /* plan.go:104-115 */
  if (in->has_part_weights) {
    if (in->part_has_weight[p]) {
      w_p = in->part_weight[p];
      stick = (double)in->part_weight[p];
    } else if (in->state_has_stickiness[s]) {
      stick = (double)in->state_stickiness[s];
    }
  }

  int32_t top = f->NU;                                             /* plan.go:134-138; NU = "" */
  {
    int32_t tlo = in->state_slot_off[in->top_state];
    if (tlo < in->state_slot_off[in->top_state + 1] && row[tlo] != NONE) top = row[tlo];
  }

  /* candidates = nodesNext minus nodes held in higher-priority states, plan.go:142-156 */
  int have_higher_key = 0;
  memcpy(f->cand, f->valid, (size_t)N);
  for (int32_t s2 = 0; s2 < f->S; s2++) {
    if (f->shape[(size_t)p * f->S + s2] == BLANCE_SHAPE_ABSENT) continue;
    if (in->state_priority[s2] >= in->state_priority[s]) continue;
    have_higher_key = 1;
    for (int32_t i = in->state_slot_off[s2]; i < in->state_slot_off[s2 + 1] && row[i] != NONE; i++)
      if (row[i] < N) f->cand[row[i]] = 0;
  }
  int n_cand = 0;
  for (int32_t n = 0; n < N; n++) {
    if (!f->cand[n]) continue;
    n_cand++;
    double cur = list_has(row, lo, hi, n) ? stick : 0.0;
    f->score[n] = node_score(f, s, n, top, cur);
  }

  int32_t chosen_buf[64];
  int32_t* chosen = k <= 64 ? chosen_buf : (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
  int n_chosen = 0;

  if (in->has_hier_rules) {                                        /* plan.go:174-226 */
    int n_rules = in->rule_off[s + 1] - in->rule_off[s];
    int max_picks = n_rules * k;
    int32_t* picks = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_picks > 0 ? max_picks : 1));
    int n_picks = 0;
    const size_t stride_a = (size_t)f->HW;
    const size_t stride_r = (size_t)(f->NU + 1) * f->HW;
    for (int32_t r = in->rule_off[s]; r < in->rule_off[s + 1]; r++) {
      int32_t h = top;
      if (h == f->NU && n_picks > 0) h = picks[0];
      for (int32_t i = 0; i < k; i++) {
        memset(f->hm, 0, sizeof(uint32_t) * (size_t)f->HW);
        for (int a = -1; a < n_picks; a++) {                       /* anchors = [h] ++ picks */
          int32_t anchor = a < 0 ? h : picks[a];
          const uint32_t* res = in->ie_mask + (size_t)r * stride_r + (size_t)anchor * stride_a;
          int empty = 1;
          for (int32_t wd = 0; wd < f->HW; wd++) empty &= (f->hm[wd] == 0);
          if (empty) memcpy(f->hm, res, sizeof(uint32_t) * (size_t)f->HW);   /* plan.go:746-749 */
          else for (int32_t wd = 0; wd < f->HW; wd++) f->hm[wd] &= res[wd];
        }
        int32_t best = argmin_masked(f, f->cand, f->hm);
        if (best >= 0) picks[n_picks++] = best;                    /* plan.go:214-216 */
        else if (n_cand > 0) picks[n_picks++] = argmin_masked(f, f->cand, NULL);   /* plan.go:217-220 */
      }
    }
    /* first k of dedupe(picks ++ ascending(cand)), plan.go:224-229 */
    for (int i = 0; i < n_picks && n_chosen < k; i++) {
      int dup = 0;
      for (int j = 0; j < n_chosen; j++) dup |= (chosen[j] == picks[i]);
      if (!dup) chosen[n_chosen++] = picks[i];
    }
    free(picks);
  }
  /* fill up from the flat (score, position) order, skipping what is chosen
   * (cand is rebuilt at the start of every step, so it can be consumed here) */
  for (int j = 0; j < n_chosen; j++) f->cand[chosen[j]] = 0;
  while (n_chosen < k) {
    int32_t best = argmin_masked(f, f->cand, NULL);
    if (best < 0) break;
    chosen[n_chosen++] = best;
    f->cand[best] = 0;
  }
  if (n_chosen < k) f->warn[(size_t)p * f->S + s] = 1;             /* plan.go:228-235 */

  for (int i = 0; i < n_chosen; i++) f->n2n[(size_t)top * N + chosen[i]] += 1;   /* plan.go:238-245 */

  /* nil vs empty result: candidateNodes stays nil only when nodesNext is empty, no
   * higher-priority key filtered it and the hierarchy block did not run (plan.go:142,
   * 149-150, 225) */
  uint8_t new_shape = BLANCE_SHAPE_LIST;
  if (n_chosen == 0 && f->n_valid == 0 && !have_higher_key && !in->has_hier_rules) new_shape = BLANCE_SHAPE_NIL;

  /* plan.go:290-301 */
  int old_len = list_len(row, lo, hi);
  int32_t old_buf[64];
  int32_t* old = old_len <= 64 ? old_buf : (int32_t*)malloc(sizeof(int32_t) * (size_t)old_len);
  memcpy(old, row + lo, sizeof(int32_t) * (size_t)old_len);
  if (f->shape[(size_t)p * f->S + s] == BLANCE_SHAPE_ABSENT) old_len = 0;
  remove_nodes(f, p, old, old_len, w_p);
  remove_nodes(f, p, chosen, n_chosen, w_p);
  for (int32_t i = lo; i < hi; i++) row[i] = (i - lo) < n_chosen ? chosen[i - lo] : NONE;
  f->shape[(size_t)p * f->S + s] = new_shape;
  for (int i = 0; i < n_chosen; i++) {
    f->counts[(size_t)s * N + chosen[i]] += w_p;
    f->tot[chosen[i]] += w_p;
  }
  if (old != old_buf) free(old);
  if (chosen != chosen_buf) free(chosen);
}

typedef struct { uint64_t key; int32_t p; } okey_t;

/* The pass loop of plan.go:268-302.  tools/spec_model.c (a design model of the product's speculative
 * pass kernel, also test infrastructure) replaces it to check that kernel's decision rules step by
 * step against one_step(). */
#ifndef FO_PASS
#define FO_PASS(f, order, lim, s) do { for (int32_t i_ = 0; i_ < (lim); i_++) one_step((f), (order)[i_].p, (s)); } while (0)
#endif
static int okey_cmp(const void* a, const void* b) {
  const okey_t* x = (const okey_t*)a; const okey_t* y = (const okey_t*)b;
  return x->key < y->key ? -1 : x->key > y->key ? 1 : 0;
}

#if defined(__GNUC__)
#define FO_EXPORT __attribute__((visibility("default")))
#else
#define FO_EXPORT
#endif

/* Same contract as blance_plan_next_map() minus the ctx; max_steps_per_pass < 0
 * means no cap (bench.py's bounded cpu sample uses a cap). */
FO_EXPORT int oracle_fast_plan_next_map_capped(const blance_plan_in* in, blance_plan_out* out,
                                               int64_t max_steps_per_pass) {
  fo_t f;
  memset(&f, 0, sizeof f);
  f.in = in;
  f.N = in->n_nodes; f.NU = in->n_node_ids; f.S = in->n_states; f.PU = in->n_parts; f.SL = in->n_slots;
  f.HW = (in->n_hier_bits + 31) / 32;
  const size_t rows_n = (size_t)f.PU * f.SL, shape_n = (size_t)f.PU * f.S;
  f.rows = (int32_t*)malloc(sizeof(int32_t) * (rows_n + 1));
  f.shape = (uint8_t*)malloc(shape_n + 1);
  f.prev_rows = (int32_t*)malloc(sizeof(int32_t) * (rows_n + 1));
  f.prev_shape = (uint8_t*)malloc(shape_n + 1);
  f.in_prev = (uint8_t*)malloc((size_t)f.PU + 1);
  f.valid = (uint8_t*)malloc((size_t)f.N + 1);
  f.counts = (int64_t*)malloc(sizeof(int64_t) * ((size_t)f.S * f.N + 1));
  f.tot = (int64_t*)malloc(sizeof(int64_t) * ((size_t)f.N + 1));
  f.n2n = (int32_t*)malloc(sizeof(int32_t) * ((size_t)(f.NU + 1) * f.N + 1));
  f.warn = out->warn;
  f.score = (double*)malloc(sizeof(double) * ((size_t)f.N + 1));
  f.cand = (uint8_t*)malloc((size_t)f.N + 1);
  f.hm = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)f.HW + 1));
  okey_t* order = (okey_t*)malloc(sizeof(okey_t) * ((size_t)f.PU + 1));

  memcpy(f.rows, in->cur_rows, sizeof(int32_t) * rows_n);
  memcpy(f.shape, in->cur_shape, shape_n);
  memcpy(f.prev_rows, in->prev_rows, sizeof(int32_t) * rows_n);
  memcpy(f.prev_shape, in->prev_shape, shape_n);
  memcpy(f.in_prev, in->part_in_prev, (size_t)f.PU);
  f.n_valid = 0;
  for (int32_t n = 0; n < f.N; n++) { f.valid[n] = !in->node_removed[n]; f.n_valid += f.valid[n]; }

  int rm_active = 0;
  for (int32_t n = 0; n < f.NU; n++) rm_active |= in->node_removed[n];
  int add_is_nil = in->add_is_nil, add_active = 1;
  const int32_t* extra = in->extra_tot_first;

  out->iters_run = 0; out->converged = 0; out->steps = 0;
  out->device_ms = out->kernel_ms = out->pass_ms = 0.f;
  out->sticky_steps = 0;

  for (int it = 0; it < in->max_iters; it++) {                      /* plan.go:32 */
    /* plan.go:83-88: working rows = partitionsToAssign rows minus removed nodes */
    if (rm_active) {
      for (int32_t p = 0; p < f.PU; p++) {
        if (!in->part_in_assign[p]) continue;
        int32_t* row = row_of(f.rows, f.SL, p);
        for (int32_t s = 0; s < f.S; s++) {
          int32_t lo = in->state_slot_off[s], hi = in->state_slot_off[s + 1], o = lo;
          for (int32_t i = lo; i < hi && row[i] != NONE; i++)
            if (!in->node_removed[row[i]]) row[o++] = row[i];
          for (; o < hi; o++) row[o] = NONE;
        }
      }
    }
    for (int32_t p = 0; p < f.PU; p++)
      if (in->part_in_assign[p])
        for (int32_t s = 0; s < f.S; s++)
          if (f.shape[(size_t)p * f.S + s] != BLANCE_SHAPE_ABSENT) f.shape[(size_t)p * f.S + s] = BLANCE_SHAPE_LIST;

    /* plan.go:94: counts from ALL of prevMap */
    memset(f.counts, 0, sizeof(int64_t) * (size_t)f.S * f.N);
    f.P = 0;
    for (int32_t p = 0; p < f.PU; p++) {
      if (!f.in_prev[p]) continue;
      f.P++;
      int64_t w = (in->has_part_weights && in->part_has_weight[p]) ? in->part_weight[p] : 1;
      const int32_t* row = f.prev_rows + (size_t)p * f.SL;
      for (int32_t s = 0; s < f.S; s++)
        for (int32_t i = in->state_slot_off[s]; i < in->state_slot_off[s + 1] && row[i] != NONE; i++)
          if (row[i] < f.N) f.counts[(size_t)s * f.N + row[i]] += w;
    }
    for (int32_t n = 0; n < f.N; n++) {
      int64_t t = extra ? extra[n] : 0;
      for (int32_t s = 0; s < f.S; s++) t += f.counts[(size_t)s * f.N + n];
      f.tot[n] = t;
    }
    memset(f.warn, 0, shape_n);                                     /* plan.go:70 */

    for (int32_t s = 0; s < f.S; s++) {                             /* plan.go:307-324 */
      if (in->state_constraints[s] <= 0) continue;
      /* partition order, plan.go:255-263,519-562 */
      int32_t n_order = 0;
      for (int32_t p = 0; p < f.PU; p++) {
        if (!in->part_in_assign[p]) continue;
        const int32_t* row = row_of(f.rows, f.SL, p);
        uint64_t bucket = 2;
        int b0 = 0;
        if (rm_active && f.in_prev[p]) {
          const int32_t* prow = f.prev_rows + (size_t)p * f.SL;
          for (int32_t i = in->state_slot_off[s]; i < in->state_slot_off[s + 1] && prow[i] != NONE; i++)
            b0 |= in->node_removed[prow[i]];
        }
        if (b0) bucket = 0;
        else if (!add_is_nil) {
          int hit = 0;
          if (add_active)
            for (int32_t i = 0; i < f.SL; i++)
              if (row[i] != NONE) hit |= in->node_added[row[i]];
          if (!hit) bucket = 1;
        }
        int64_t w = (in->has_part_weights && in->part_has_weight[p]) ? in->part_weight[p] : 1;
        uint64_t wkey = (uint64_t)(999999999LL - w);                /* plan.go:539 */
        order[n_order].key = (bucket << 62) | ((wkey & 0xFFFFFFFFull) << 30) | (uint64_t)in->part_name_rank[p];
        order[n_order].p = p;
        n_order++;
      }
      qsort(order, (size_t)n_order, sizeof(okey_t), okey_cmp);
      memset(f.n2n, 0, sizeof(int32_t) * (size_t)(f.NU + 1) * f.N); /* plan.go:266 */
      int32_t lim = n_order;
      if (max_steps_per_pass >= 0 && max_steps_per_pass < lim) lim = (int32_t)max_steps_per_pass;
      FO_PASS(&f, order, lim, s);
    }
    out->iters_run = it + 1;

    /* plan.go:36-42 */
    int match = 1;
    for (int32_t p = 0; p < f.PU && match; p++) {
      if (!in->part_in_assign[p]) continue;
      if (!f.in_prev[p] || (f.in_prev[p] & 2)) { match = 0; break; }   /* bit 1: prevMap entry with non-model keys */
      if (memcmp(f.shape + (size_t)p * f.S, f.prev_shape + (size_t)p * f.S, (size_t)f.S)) { match = 0; break; }
      if (memcmp(f.rows + (size_t)p * f.SL, f.prev_rows + (size_t)p * f.SL, sizeof(int32_t) * (size_t)f.SL)) match = 0;
    }
    if (match) { out->converged = 1; break; }
    /* plan.go:49-55 */
    for (int32_t p = 0; p < f.PU; p++) {
      if (!in->part_in_assign[p]) continue;
      memcpy(f.prev_rows + (size_t)p * f.SL, f.rows + (size_t)p * f.SL, sizeof(int32_t) * (size_t)f.SL);
      memcpy(f.prev_shape + (size_t)p * f.S, f.shape + (size_t)p * f.S, (size_t)f.S);
      f.in_prev[p] = 1;
    }
    rm_active = 0; add_is_nil = 0; add_active = 0;
    extra = in->extra_tot_rest;
  }

  memcpy(out->next_rows, f.rows, sizeof(int32_t) * rows_n);
  memcpy(out->next_shape, f.shape, shape_n);
  out->steps = f.steps;

  free(f.rows); free(f.shape); free(f.prev_rows); free(f.prev_shape); free(f.in_prev); free(f.valid);
  free(f.counts); free(f.tot); free(f.n2n); free(f.score); free(f.cand); free(f.hm); free(order);
  return BLANCE_OK;
}

FO_EXPORT int oracle_fast_plan_next_map(const blance_plan_in* in, blance_plan_out* out) {
  return oracle_fast_plan_next_map_capped(in, out, -1);
}

/* moves.go:41-136 per partition */
FO_EXPORT int oracle_fast_calc_partition_moves(int32_t n_parts, int32_t n_states, int32_t n_visit_states,
                                               const int32_t* slot_off, const int32_t* beg_rows,
                                               const int32_t* end_rows, int32_t favor_min_nodes, int32_t max_ops,
                                               int32_t* op_node, uint8_t* op_state, uint8_t* op_kind,
                                               int32_t* op_count) {
  const int32_t SL = slot_off[n_states];
  for (int32_t p = 0; p < n_parts; p++) {
    const int32_t* beg = beg_rows + (size_t)p * SL;
    const int32_t* end = end_rows + (size_t)p * SL;
    int32_t* on = op_node + (size_t)p * max_ops;
    uint8_t* os = op_state + (size_t)p * max_ops;
    uint8_t* ok = op_kind + (size_t)p * max_ops;
    int cnt = 0;
#define IN_ANY(row, node, res) do { res = 0; for (int32_t _i = 0; _i < SL; _i++) res |= (row[_i] == (node)); } while (0)
#define EMIT(node, st, kind) do { int _seen = 0; for (int _j = 0; _j < cnt; _j++) _seen |= (on[_j] == (node)); \
      if (!_seen && cnt < max_ops) { on[cnt] = (node); os[cnt] = (uint8_t)(st); ok[cnt] = (uint8_t)(kind); cnt++; } } while (0)
    for (int step = 0; step < n_visit_states; step++) {
      int si = favor_min_nodes ? n_visit_states - 1 - step : step;
      int32_t lo = slot_off[si], hi = slot_off[si + 1];
      for (int phase = 0; phase < 4; phase++) {
        /* favorMinNodes=false: promote, demote, add, del; true: del, demote, promote, add */
        int what = favor_min_nodes ? (int[]){3, 1, 0, 2}[phase] : phase;
        if (what == 0 || what == 1) {          /* promote: found in a LOWER state of beg; demote: in a HIGHER one */
          int jlo = what == 0 ? si + 1 : 0, jhi = what == 0 ? n_visit_states : si;
          for (int32_t i = lo; i < hi && end[i] != NONE; i++)
            for (int j = jlo; j < jhi; j++)
              for (int32_t b = slot_off[j]; b < slot_off[j + 1] && beg[b] != NONE; b++)
                if (beg[b] == end[i]) EMIT(end[i], si, what == 0 ? BLANCE_OP_PROMOTE : BLANCE_OP_DEMOTE);
        } else if (what == 2) {                /* add: end[s] \ beg[s], also absent from every beg state */
          for (int32_t i = lo; i < hi && end[i] != NONE; i++) {
            int inb; IN_ANY(beg, end[i], inb);
            if (!list_has(beg, lo, hi, end[i]) && !inb) EMIT(end[i], si, BLANCE_OP_ADD);
          }
        } else {                               /* del: beg[s] \ end[s], also absent from every end state */
          for (int32_t i = lo; i < hi && beg[i] != NONE; i++) {
            int ine; IN_ANY(end, beg[i], ine);
            if (!list_has(end, lo, hi, beg[i]) && !ine) EMIT(beg[i], BLANCE_OP_STATE_NONE, BLANCE_OP_DEL);
          }
        }
      }
    }
    op_count[p] = cnt;
#undef IN_ANY
#undef EMIT
  }
  return BLANCE_OK;
}

/* One round of findAvailableMovesUnlocked (orchestrate.go:749-763) over CSR move lists, followed by
 * LowestWeightPartitionMoveForNode (orchestrate.go:177-186, MoveOpWeight orchestrate.go:189-194) per node.
 * The reference walks a Go map (random order); this restatement - like the product - walks partitions in
 * ascending index, so per-node lists are ascending and weight ties go to the lowest partition index. */
FO_EXPORT int oracle_fast_moves_available(int32_t n_parts, int32_t n_node_ids, const int64_t* op_off,
                                          const int32_t* op_node, const uint8_t* op_kind, const int32_t* next,
                                          int32_t* node_off, int32_t* node_parts, int32_t* best_part) {
  static const int weight[4] = {3 /* add */, 4 /* del */, 1 /* promote */, 2 /* demote */};
  for (int32_t n = 0; n <= n_node_ids; n++) node_off[n] = 0;
  for (int32_t p = 0; p < n_parts; p++) {
    int64_t len = op_off[p + 1] - op_off[p];
    if (next[p] < 0 || next[p] >= len) continue;                      /* nextMoves.Next < len(nextMoves.Moves) */
    int32_t node = op_node[op_off[p] + next[p]];
    if (node >= 0 && node < n_node_ids) node_off[node + 1]++;
  }
  for (int32_t n = 0; n < n_node_ids; n++) node_off[n + 1] += node_off[n];
  int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n_node_ids + 1));
  memcpy(fill, node_off, sizeof(int32_t) * ((size_t)n_node_ids + 1));
  for (int32_t p = 0; p < n_parts; p++) {
    int64_t len = op_off[p + 1] - op_off[p];
    if (next[p] < 0 || next[p] >= len) continue;
    int32_t node = op_node[op_off[p] + next[p]];
    if (node >= 0 && node < n_node_ids) node_parts[fill[node]++] = p;
  }
  free(fill);
  for (int32_t n = 0; n < n_node_ids; n++) {
    int32_t r = -1, rw = 0;
    for (int32_t i = node_off[n]; i < node_off[n + 1]; i++) {          /* r = 0; if weight[moves[r]] > weight[move] r = i */
      int32_t p = node_parts[i];
      int w = weight[op_kind[op_off[p] + next[p]] & 3];
      if (r < 0 || rw > w) { r = p; rw = w; }
    }
    best_part[n] = r;
  }
  return BLANCE_OK;
}
