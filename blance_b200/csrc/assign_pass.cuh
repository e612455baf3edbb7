// blance_b200/csrc/assign_pass.cuh — the hot kernel: one state pass of blance's
// greedy planner (assignStateToPartitions + findBestNodes, plan.go:98-303).
//
// The pass is a loop-carried chain: every step reads the node counts the previous
// step wrote (plan.go:286-301), so partitions are walked strictly in the
// reference's order by ONE CTA per plan instance while the N candidate nodes of a
// step are scored in parallel, one node (or NPT nodes) per thread:
//
//   * per-node state lives in REGISTERS of the owning thread: the count of the
//     state being assigned, the all-state total, the cached "filled" term and the
//     node weight; only the owner ever updates them, so no atomics are needed;
//   * the step's partition record (row, weight, shape bits) is prefetched by warp 0
//     three steps ahead into a shared-memory ring, so the chain never waits on a
//     dependent order[] -> rows[] global load;
//   * the score is IEEE binary64 with the reference's exact operation order
//     (plan.go:634-689) using __dadd_rn/__dmul_rn/__ddiv_rn (never fused);
//   * "sort candidates, take the first k" (plan.go:171-172, 228-229) becomes k
//     CTA-wide arg-mins of the 96-bit key (ordered score bits, node position),
//     each done with redux.sync min steps inside the warp and one
//     shared-memory exchange + barrier across warps (plan.go:617-628 is a strict
//     total order, so arg-min == first of the sort);
//   * hierarchy rules (plan.go:174-226) are bit-set chains evaluated redundantly
//     by every warp (lane j owns word j), so a pick costs one masked arg-min.
#pragma once

#include <cuda_runtime.h>

#include "blance_b200.h"
#include "device_types.cuh"

namespace blance_dev {

struct Best { uint32_t hi, lo, pos; };

// order-preserving map double -> uint64 (total order == numeric order, -0 == +0 canonicalised by the caller)
__device__ __forceinline__ unsigned long long score_key(double r) {
  long long b = __double_as_longlong(r);
  unsigned long long u = (unsigned long long)b;
  return (b < 0) ? ~u : (u | 0x8000000000000000ull);
}

// lexicographic min of (hi, lo, pos) over the warp; non-participants pass pos = 0xFFFFFFFF
__device__ __forceinline__ Best warp_argmin(Best v) {
  const unsigned full = 0xFFFFFFFFu;
  uint32_t mhi = __reduce_min_sync(full, v.hi);
  uint32_t lo2 = (v.hi == mhi) ? v.lo : 0xFFFFFFFFu;
  uint32_t mlo = __reduce_min_sync(full, lo2);
  uint32_t p2 = (v.hi == mhi && v.lo == mlo) ? v.pos : 0xFFFFFFFFu;
  uint32_t mpos = __reduce_min_sync(full, p2);
  return Best{mhi, mlo, mpos};
}

// Shared memory of the pass kernel.
struct PassSmem {
  uint4 xchg[2][32];                       // per-warp partial arg-min, double buffered
  int32_t ring[BL_RING][BL_SLP_MAX + 4];   // step records: row[SLP], meta, weight, flags, partition
};

// CTA-wide arg-min.  All threads must call; returns the same Best in every thread.
__device__ __forceinline__ Best cta_argmin(Best mine, PassSmem& sm, int& xbuf, int nwarps, int warp, int lane) {
  Best w = warp_argmin(mine);
  if (nwarps == 1) return w;
  if (lane == 0) sm.xchg[xbuf][warp] = make_uint4(w.hi, w.lo, w.pos, 0u);
  __syncthreads();
  uint4 e = (lane < nwarps) ? sm.xchg[xbuf][lane] : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u);
  xbuf ^= 1;
  return warp_argmin(Best{e.x, e.y, e.z});
}

template <int NPT>
__global__ void __launch_bounds__(1024, 1) k_assign_pass(DPool pool, int s) {
  DInst& D = pool.insts[blockIdx.x];
  if (!D.active || s >= D.S) return;
  const int k = D.state_constraints[s];
  if (k <= 0) return;

  __shared__ PassSmem sm;
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int N = D.N, NU = D.NU, S = D.S, SLP = D.SLP, HW = D.HW;
  const int n_assign = D.n_assign;
  const int lo_s = D.state_slot_off[s], hi_s = D.state_slot_off[s + 1];
  const int prio_s = D.state_priority[s];
  const int top_slot = D.state_slot_off[D.top_state];
  const bool top_has_slot = D.state_slot_off[D.top_state + 1] > top_slot;
  const int Pn = D.P;
  const double Pd = (double)Pn;
  const bool has_pw = D.has_part_weights != 0;
  const bool hier = D.has_hier_rules != 0;
  const int booster = D.booster;
  const double state_stick = D.state_has_stickiness[s] ? (double)D.state_stickiness[s] : 1.5;

  // slot -> state map and the set of higher-priority states, as small bit tables
  uint32_t higher_states = 0;   // bit s2: priority[s2] < priority[s]
  for (int s2 = 0; s2 < S; ++s2)
    if (D.state_priority[s2] < prio_s) higher_states |= 1u << s2;

  int32_t* rows = pool.rows + D.rows_off;
  uint32_t* pmeta = pool.pmeta + D.part_off;
  const uint8_t* pflags = pool.pflags + D.part_off;
  const int32_t* pweight = pool.pweight + D.part_off;
  const int32_t* order = pool.order + D.part_off;
  int32_t* counts = pool.counts + D.counts_off;
  int32_t* n2n = pool.n2n + D.n2n_off;
  const int32_t* extra = (D.use_rest ? pool.extra_rest : pool.extra_first) + D.node_off;
  const uint32_t* ie_mask = pool.ie_mask + D.mask_off;

  // ---- per-node state in registers -------------------------------------------------
  int32_t c_s[NPT], tot[NPT];
  double ff[NPT], wd[NPT];
  uint32_t valid_bits = 0, wdiv_bits = 0, wboost_bits = 0;
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const int n = tid + j * T;
    c_s[j] = 0; tot[j] = 0; ff[j] = 0.0; wd[j] = 1.0;
    if (n < N) {
      int t = extra[n];
      for (int s2 = 0; s2 < S; ++s2) t += counts[s2 * N + n];
      tot[j] = t;
      c_s[j] = counts[s * N + n];
      if (!pool.node_removed[D.nodeid_off + n]) valid_bits |= 1u << j;
      if (D.has_node_weights && pool.node_has_weight[D.node_off + n]) {
        const int w = pool.node_weight[D.node_off + n];
        wd[j] = (double)w;
        if (w > 0) wdiv_bits |= 1u << j;
        else if (w < 0 && booster == BLANCE_BOOSTER_CBGT_MAX) wboost_bits |= 1u << j;
      }
      if (Pn > 0) ff[j] = __ddiv_rn(__dmul_rn(0.001, (double)t), Pd);     // plan.go:650
    }
  }

  // ---- step-record ring, filled by warp 0 -------------------------------------------
  // record layout in ring[slot]: [0..SLP) row, [SLP] meta, [SLP+1] weight, [SLP+2] flags, [SLP+3] partition
  const int REC = SLP + 4;
  int32_t p_ahead = -1;      // warp 0: partition index of step i+2 (uniform)
  int32_t rec_reg = 0;       // warp 0, lane t < REC: field t of step i+1's record, loaded one step ago
  auto load_field = [&](int32_t p, int t) -> int32_t {
    if (t < SLP) return rows[(size_t)p * SLP + t];
    if (t == SLP) return (int32_t)pmeta[p];
    if (t == SLP + 1) return pweight[p];
    if (t == SLP + 2) return (int32_t)pflags[p];
    return p;
  };
  if (warp == 0) {
    // prologue: record 0 straight into the ring, record 1 into registers, p of step 2
    if (n_assign > 0) {
      const int32_t p0 = order[0];
      for (int t = lane; t < REC; t += 32) sm.ring[0][t] = load_field(p0, t);
    }
    if (n_assign > 1) {
      const int32_t p1 = order[1];
      if (lane < REC) rec_reg = load_field(p1, lane);   // REC <= 36 needs two trips; handled below
    }
    if (n_assign > 2) p_ahead = order[2];
  }
  // REC can exceed 32 (SLP up to 32): lanes cover fields lane and lane+32.
  int32_t rec_reg_hi = 0;
  if (warp == 0 && n_assign > 1 && lane + 32 < REC) rec_reg_hi = load_field(order[1], lane + 32);
  __syncthreads();

  int xbuf = 0;
  long long steps = 0;

  for (int i = 0; i < n_assign; ++i) {
    // ---- warp 0: advance the prefetch pipeline --------------------------------------
    if (warp == 0) {
      if (i + 1 < n_assign) {          // publish record i+1 (read by everyone at step i+1)
        int32_t* dst = sm.ring[(i + 1) % BL_RING];
        if (lane < REC) dst[lane] = rec_reg;
        if (lane + 32 < REC) dst[lane + 32] = rec_reg_hi;
      }
      if (i + 2 < n_assign) {          // start loading record i+2
        if (lane < REC) rec_reg = load_field(p_ahead, lane);
        if (lane + 32 < REC) rec_reg_hi = load_field(p_ahead, lane + 32);
      }
      if (i + 3 < n_assign) p_ahead = order[i + 3];
    }

    const int32_t* rec = sm.ring[i % BL_RING];
    const uint32_t meta = (uint32_t)rec[SLP];
    const int32_t w_raw = rec[SLP + 1];
    const uint32_t flags = (uint32_t)rec[SLP + 2];
    const int32_t p = rec[SLP + 3];

    int32_t w_p = 1;
    double stick = 1.5;                                        // plan.go:104-115
    if (has_pw) {
      if (flags & PF_HAS_WEIGHT) { w_p = w_raw; stick = (double)w_raw; }
      else stick = state_stick;
    }
    int32_t top = NU;                                          // plan.go:134-138 (NU stands for "")
    if (top_has_slot) { const int32_t t0 = rec[top_slot]; if (t0 != BLANCE_NO_NODE) top = t0; }

    // ---- membership of my node(s) in the row --------------------------------------
    uint32_t memb[NPT];       // bit s2: my node is in the list of state s2
    bool have_higher_key = false;
#pragma unroll
    for (int j = 0; j < NPT; ++j) memb[j] = 0;
    for (int s2 = 0; s2 < S; ++s2) {
      if (meta_shape(meta, s2) == BLANCE_SHAPE_ABSENT) continue;
      if ((higher_states >> s2) & 1u) have_higher_key = true;
      for (int sl = D.state_slot_off[s2]; sl < D.state_slot_off[s2 + 1]; ++sl) {
        const int32_t x = rec[sl];
        if (x == BLANCE_NO_NODE) break;
#pragma unroll
        for (int j = 0; j < NPT; ++j)
          if (x == tid + j * T) memb[j] |= 1u << s2;
      }
    }

    // ---- score (plan.go:634-689) and key -------------------------------------------
    unsigned long long key[NPT];
    uint32_t cand_bits = 0;
    const int32_t* n2n_row = n2n + (size_t)top * N;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int n = tid + j * T;
      key[j] = ~0ull;
      const bool cand = ((valid_bits >> j) & 1u) && !(memb[j] & higher_states);   // plan.go:142-156
      if (!cand) continue;
      cand_bits |= 1u << j;
      const double cur = ((memb[j] >> s) & 1u) ? stick : 0.0;                      // plan.go:654-662
      double r = (double)c_s[j];
      if (Pn > 0) {
        const int32_t q = n2n_row[n];
        if (q != 0) r = __dadd_rn(r, __ddiv_rn((double)q, Pd));                    // plan.go:641-642,672
        r = __dadd_rn(r, ff[j]);                                                   // plan.go:673
      }
      if ((wdiv_bits >> j) & 1u) {
        r = __ddiv_rn(r, wd[j]);                                                   // plan.go:679
      } else if ((wboost_bits >> j) & 1u) {
        double b = -wd[j];                                                         // control_test.go:19-26
        if (b < cur) b = cur;
        r = __dadd_rn(r, b);
      }
      r = __dsub_rn(r, cur);                                                       // plan.go:686
      r = __dadd_rn(r, 0.0);                                                       // -0.0 -> +0.0 (equal under Go's <)
      key[j] = score_key(r);
    }

    // arg-min over nodes selected by `bits` (per-thread bit j); everyone calls
    auto argmin_bits = [&](uint32_t bits) -> uint32_t {
      Best mine{0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        if (!((bits >> j) & 1u)) continue;
        const uint32_t hi = (uint32_t)(key[j] >> 32), lo = (uint32_t)key[j], pos = (uint32_t)(tid + j * T);
        if (mine.pos == 0xFFFFFFFFu || hi < mine.hi || (hi == mine.hi && (lo < mine.lo || (lo == mine.lo && pos < mine.pos))))
          mine = Best{hi, lo, pos};
      }
      return cta_argmin(mine, sm, xbuf, nwarps, warp, lane).pos;
    };

    int32_t chosen[BL_K_MAX];
    int n_chosen = 0;
    uint32_t taken_bits = 0;

    if (hier) {                                                                    // plan.go:174-226
      int32_t picks[BL_PICK_MAX];
      int n_picks = 0;
      uint32_t flat0 = 0xFFFFFFFEu;   // lazily computed best of the flat order ("candidateNodes[0]")
      for (int r = D.rule_off[s]; r < D.rule_off[s + 1]; ++r) {
        int32_t h = top;
        if (h == NU && n_picks > 0) h = picks[0];                                  // plan.go:178-181
        for (int it = 0; it < k; ++it) {
          // running intersection over anchors [h] ++ picks with replace-on-empty (plan.go:743-751);
          // lane l holds words l, l+32, ... of the HW-word set
          uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
          for (int a = -1; a < n_picks; ++a) {
            const int32_t anchor = a < 0 ? h : picks[a];
            const uint32_t* res = ie_mask + ((size_t)r * (NU + 1) + anchor) * HW;
            const uint32_t r0 = lane < HW ? __ldg(res + lane) : 0u;
            const uint32_t r1 = lane + 32 < HW ? __ldg(res + lane + 32) : 0u;
            const uint32_t r2 = lane + 64 < HW ? __ldg(res + lane + 64) : 0u;
            const uint32_t r3 = lane + 96 < HW ? __ldg(res + lane + 96) : 0u;
            const bool empty = !__any_sync(0xFFFFFFFFu, (m0 | m1 | m2 | m3) != 0u);
            if (empty) { m0 = r0; m1 = r1; m2 = r2; m3 = r3; }
            else { m0 &= r0; m1 &= r1; m2 &= r2; m3 &= r3; }
          }
          uint32_t bits = 0;
#pragma unroll
          for (int j = 0; j < NPT; ++j) {
            const int wi = (tid + j * T) >> 5;            // word of my node: uniform across the warp
            const int src = wi & 31, grp = wi >> 5;
            const uint32_t word = __shfl_sync(0xFFFFFFFFu, grp == 0 ? m0 : grp == 1 ? m1 : grp == 2 ? m2 : m3, src);
            if ((word >> lane) & 1u) bits |= 1u << j;
          }
          const uint32_t best = argmin_bits(bits & cand_bits);
          if (best != 0xFFFFFFFFu) {
            if (n_picks < BL_PICK_MAX) picks[n_picks++] = (int32_t)best;            // plan.go:214-216
          } else {
            if (flat0 == 0xFFFFFFFEu) flat0 = argmin_bits(cand_bits);
            if (flat0 != 0xFFFFFFFFu && n_picks < BL_PICK_MAX) picks[n_picks++] = (int32_t)flat0;   // plan.go:217-220
          }
        }
      }
      for (int a = 0; a < n_picks && n_chosen < k; ++a) {                           // plan.go:224-229
        bool dup = false;
        for (int b = 0; b < n_chosen; ++b) dup |= (chosen[b] == picks[a]);
        if (!dup) chosen[n_chosen++] = picks[a];
      }
#pragma unroll
      for (int j = 0; j < NPT; ++j)
        for (int b = 0; b < n_chosen; ++b)
          if (chosen[b] == tid + j * T) taken_bits |= 1u << j;
    }
    while (n_chosen < k) {                                // the flat (score, position) order
      const uint32_t best = argmin_bits(cand_bits & ~taken_bits);
      if (best == 0xFFFFFFFFu) break;
      chosen[n_chosen++] = (int32_t)best;
#pragma unroll
      for (int j = 0; j < NPT; ++j)
        if ((int)best == tid + j * T) taken_bits |= 1u << j;
    }

    // ---- apply (plan.go:238-245, 290-301) on the owners' registers ---------------------
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int n = tid + j * T;
      const bool is_cur = (memb[j] >> s) & 1u, tk = (taken_bits >> j) & 1u;
      if (!(is_cur || tk) || n >= N) continue;
      int32_t t = tot[j];
      uint32_t dec = memb[j];                       // removed from every state that lists it, once per state
      if ((dec >> s) & 1u) { c_s[j] -= w_p; t -= w_p; dec &= ~(1u << s); }
      while (dec) {
        const int s2 = __ffs(dec) - 1;
        dec &= dec - 1;
        counts[s2 * N + n] -= w_p;                  // another state's count: only the owner touches it
        t -= w_p;
      }
      if (tk) { c_s[j] += w_p; t += w_p; n2n[(size_t)top * N + n] += 1; }
      if (t != tot[j]) {
        tot[j] = t;
        if (Pn > 0) ff[j] = __ddiv_rn(__dmul_rn(0.001, (double)t), Pd);
      }
    }

    // ---- thread 0 writes the partition's new row ---------------------------------------
    if (tid == 0) {
      uint32_t nmeta = meta;
      int32_t* row_out = rows + (size_t)p * SLP;
      for (int s2 = 0; s2 < S; ++s2) {
        if (s2 == s || meta_shape(meta, s2) == BLANCE_SHAPE_ABSENT) continue;
        nmeta = meta_set_shape(nmeta, s2, BLANCE_SHAPE_LIST);      // misc.go:29: non-nil after removal
        int o = D.state_slot_off[s2];
        bool changed = false;
        for (int sl = D.state_slot_off[s2]; sl < D.state_slot_off[s2 + 1]; ++sl) {
          const int32_t x = rec[sl];
          if (x == BLANCE_NO_NODE) break;
          bool rm = false;
          for (int q = lo_s; q < hi_s && rec[q] != BLANCE_NO_NODE; ++q) rm |= (rec[q] == x);
          for (int b = 0; b < n_chosen; ++b) rm |= (chosen[b] == x);
          if (rm) { changed = true; continue; }
          if (changed) row_out[o] = x;
          ++o;
        }
        if (changed)
          for (; o < D.state_slot_off[s2 + 1]; ++o) row_out[o] = BLANCE_NO_NODE;
      }
      for (int sl = lo_s; sl < hi_s; ++sl) row_out[sl] = (sl - lo_s) < n_chosen ? chosen[sl - lo_s] : BLANCE_NO_NODE;
      // nil result: candidateNodes stays nil only if nodesNext is empty, no higher-priority key
      // filtered it and the hierarchy block did not run (plan.go:142,149-150,225)
      const bool nil = (n_chosen == 0) && D.n_valid == 0 && !have_higher_key && !hier;
      nmeta = meta_set_shape(nmeta, s, nil ? BLANCE_SHAPE_NIL : BLANCE_SHAPE_LIST);
      if (n_chosen < k) nmeta |= 1u << (16 + s);                    // plan.go:228-235
      pmeta[p] = nmeta;
    }
    ++steps;
    // every thread reads ring[i % RING] only during step i; it is rewritten at step i+3 at the
    // earliest, with at least one barrier (every step runs >= 1 arg-min) in between when
    // nwarps > 1.  A single-warp CTA is ordered by program order + __syncwarp.
    if (nwarps == 1) __syncwarp();
  }

  // ---- write the per-node counts of this state back ---------------------------------
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const int n = tid + j * T;
    if (n < N) counts[s * N + n] = c_s[j];
  }
  if (tid == 0) D.steps += steps;
}

}  // namespace blance_dev
