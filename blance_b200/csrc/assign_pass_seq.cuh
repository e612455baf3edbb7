// blance_b200/csrc/assign_pass_seq.cuh — the assign pass for rebalances: sequencer
// warps decide the "sticky" steps alone and wake the whole CTA only when they must.
//
// Same chain as assign_pass.cuh (assignStateToPartitions + findBestNodes,
// plan.go:98-303), same results.  The observation: in a rebalance most steps re-elect the
// partition's current nodes.  Every node score is bounded below by its "base" key — the
// same formula (plan.go:634-689) with nodeToNodeCounts = 0 and no stickiness; all its
// operations are monotone — and base keys only change when a count changes.  So:
//
//   * the CTA keeps the smallest base keys of the live nodes in shared memory (glist: as
//     many as a row can block plus two, at most BL_GLIST) and a shared-memory mirror of the
//     per-node score inputs;
//   * the SEQUENCER warps (1 or SEQ_W_MAX per CTA) walk the steps in windows: one (step,
//     current node) item per lane.  For a step whose partition holds exactly k clean current
//     nodes the lanes compute the k exact scores from the mirror, find the smallest cached
//     base key among the nodes the partition could still take, and if the worst current node
//     beats it, the reference's sort would put exactly the current nodes first: the step is
//     decided with no arg-min round (counts do not change, so the cache stays valid; only
//     nodeToNodeCounts is bumped).  A window is committed up to its first step that is not
//     sticky; two steps of a window that share (top, node) are found exactly through a small
//     hash set in shared memory and the later one starts the next window;
//   * any other step is handed to the compute warps through two named barriers: they run
//     the full evaluation (the code of the lock-step kernel), update their registers, the
//     mirror and the base keys, and report the outcome; a count change drops the cache, which
//     is rebuilt (arg-min rounds over the base keys) after up to BL_CALM_MIN quiet steps.
//
// The host picks this kernel per pass when the state has no hierarchy rules, k <= BL_FAST_K
// and at least a quarter of the rows are eligible (k_pick_mode); otherwise the lock-step
// kernel runs.
#pragma once

#include "pass_common.cuh"

namespace blance_dev {

#define BL_GLIST 8         // cached smallest base keys
#define BL_FAST_K 4        // the sticky decision handles constraints up to this
#define BL_CALM_MIN 3      // rebuild the cache only after this many steps without a count change
#define SEQ_W_MAX 4        // sequencer warps per CTA (1 for wide batches, SEQ_W_MAX otherwise: c_abi.cu)
#define SEQ_RSTRIDE 16     // ring record stride in words (rows have SLP <= 8 here: k_pick_mode)
#define SEQ_DTAB_LOG 11    // (top, node) table of a window: 2^11 slots for at most 256 items

enum : int32_t { SEQ_CMD_EXIT = -1, SEQ_CMD_REBUILD = -2 };
enum : int { BAR_GO = 1, BAR_DONE = 2, BAR_ROUND = 3, BAR_W0 = 4, BAR_W1 = 5, BAR_W2 = 6, BAR_W3 = 7 };
enum : uint32_t { NF_VALID = 1, NF_BOOST = 2 };

struct SeqSmem {
  uint4 xchg[2][32];
  double qtab[BL_QTAB];
  alignas(16) int32_t slot_bit[BL_SLP_MAX];
  uint4 glist[BL_GLIST];
  unsigned long long dtab[1 << SEQ_DTAB_LOG];   // open-addressing set of the window's (top, node) pairs
  uint32_t dmin[1 << SEQ_DTAB_LOG];             // smallest item id per slot
  int32_t cmd;
  int32_t res_n, res_same;
  int32_t res_chosen[BL_K_MAX];
  int32_t g_len, g_complete;
  int32_t win_i, win_glen, win_gcomplete, win_gen;
  int32_t win_acc[SEQ_W_MAX];
};

// dynamic shared memory: the per-node mirror (cd, ff, wd, wy doubles + a flag byte), then the record ring
// (4 windows of 32*W steps, SEQ_RSTRIDE words each)
__host__ __device__ inline size_t seq_ring_offset(int N) { return ((size_t)N * 33 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t seq_dyn_smem_bytes(int N, int W) { return seq_ring_offset(N) + (size_t)(4 * 32 * W) * SEQ_RSTRIDE * 4; }

// CTA-wide arg-min over the compute warps only (named barrier BAR_ROUND, TC threads)
__device__ __forceinline__ Best seq_argmin(Best mine, uint32_t xchg, int& xbuf, int cw, int TC, int warp, int lane) {
  const Best w = warp_argmin(mine);
  const uint32_t base = xchg + (uint32_t)xbuf * 512u;
  if (lane == 0) sts128(base + (uint32_t)warp * 16u, w.hi, w.lo, w.pos, 0u);
  bar_sync(BAR_ROUND, TC);
  int4 e = make_int4(-1, -1, -1, 0);
  if (lane < cw) e = lds128(base + (uint32_t)lane * 16u);
  xbuf ^= 1;
  return warp_argmin(Best{(uint32_t)e.x, (uint32_t)e.y, (uint32_t)e.z});
}

// exact key from explicit inputs (plan.go:634-689); flags: NF_BOOST => wd holds the (negative) weight
__device__ __forceinline__ unsigned long long key_from(double cd, double ff, double wd, double wy, bool boost, bool has_nw,
                                                       int32_t q, double cur, uint32_t qtab_a, double Pd, double Py) {
  double qv = lds64f(qtab_a + ((uint32_t)q < BL_QTAB ? (uint32_t)q : 0u) * 8u);            // plan.go:641-642
  if ((uint32_t)q >= BL_QTAB) qv = q_over_p_slow(q, Pd, Py);
  const double base = __dadd_rn(__dadd_rn(cd, qv), ff);                                    // plan.go:672-673
  double r = base;
  if (has_nw && !boost) r = div_exact(r, wd, wy);                                          // plan.go:679 (identity for weight 1 / none)
  if (boost) {                                                                             // plan.go:680-681, control_test.go:19-26
    double b = -wd;
    if (b < cur) b = cur;
    r = __dadd_rn(base, b);
  }
  r = __dsub_rn(r, cur);                                                                   // plan.go:686
  // (r is never -0.0: counts convert to +0.0, x - x rounds to +0.0, and +0.0 / w = +0.0)
  return score_key(r);
}

// blockDim.x = TC + 32*W (TC compute threads, a power of two; W sequencer warps) ; dynamic smem = seq_dyn_smem_bytes(N, W)
// K = the state's constraints (1..BL_FAST_K): a template parameter so that the per-step lane groups,
// the key exchange and the divisions by k are compile-time.
template <int NPT, int K, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_assign_pass_seq(DPool pool, int s, int TC) {
  DInst& D = pool.insts[blockIdx.x];
  if (!D.active || s >= D.S || D.pass_mode != 1) return;
  constexpr int k = K;
  if (D.state_constraints[s] != K) return;       // this instantiation serves the instances whose state has exactly K

  __shared__ SeqSmem sm;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NT = blockDim.x, cw = TC >> 5;
  const int W = (NT - TC) >> 5;                  // sequencer warps: cw (the leader) .. cw + W - 1
  const int NL = TC + 32;                        // compute warps + leader: BAR_GO / BAR_DONE
  const int logTC = 31 - __clz(TC);
  const bool is_seq = warp >= cw;
  const int N = D.N, S = D.S, SL = D.SL, SLP = D.SLP;
  const int n_assign = D.n_assign;
  const int lo_s = D.state_slot_off[s];
  const int Pn = D.P;
  const double Pd = Pn > 0 ? (double)Pn : 1.0;
  const double Py = __ddiv_rn(1.0, Pd);
  const bool has_nw = D.has_node_weights != 0;

  uint32_t higher_states = 0;              // bit s2: priority[s2] < priority[s]  (plan.go:146-152)
  for (int s2 = 0; s2 < S; ++s2)
    if (D.state_priority[s2] < D.state_priority[s]) higher_states |= 1u << s2;

  const int REC = SLP + BL_REC_HDR;
  const int32_t* stream = pool.stream + D.stream_off;
  int32_t* ostream = pool.ostream + D.stream_off;
  int32_t* counts = pool.counts + D.counts_off;
  int32_t* n2n = pool.n2n + D.n2n_off;
  const int32_t* extra = (D.use_rest ? pool.extra_rest : pool.extra_first) + D.node_off;

  // mirror of the per-node score inputs: 32 bytes per node {cd, ff, wd, wy}, then one flag byte per node
  double* nd = reinterpret_cast<double*>(dyn_smem);
  uint8_t* nd_flag = reinterpret_cast<uint8_t*>(nd + 4 * (size_t)N);
  const uint32_t nd_a = (uint32_t)__cvta_generic_to_shared(dyn_smem);
  const uint32_t ndf_a = nd_a + 32u * (uint32_t)N;

  const uint32_t sm_a = (uint32_t)__cvta_generic_to_shared(&sm);
  const uint32_t ring_a = nd_a + (uint32_t)seq_ring_offset(N);
  const uint32_t rmask = (uint32_t)(4 * 32 * W) - 1u;     // ring records - 1
  const uint32_t xchg_a = sm_a + (uint32_t)offsetof(SeqSmem, xchg);
  const uint32_t qtab_a = sm_a + (uint32_t)offsetof(SeqSmem, qtab);
  const uint32_t sbit_a = sm_a + (uint32_t)offsetof(SeqSmem, slot_bit);
  const uint32_t glist_a = sm_a + (uint32_t)offsetof(SeqSmem, glist);

  // ---- pass constants ---------------------------------------------------------------------
  for (int i = tid; i < SLP; i += NT) {
    int st = 0;
    while (st + 1 < S && i >= D.state_slot_off[st + 1]) ++st;
    sm.slot_bit[i] = (i < SL) ? (1 << st) : 0;
  }
  for (int i = tid; i < BL_QTAB; i += NT) sm.qtab[i] = Pn > 0 ? __ddiv_rn((double)i, Pd) : 0.0;
  for (int i = tid; i < (1 << SEQ_DTAB_LOG); i += NT) { sm.dtab[i] = ~0ull; sm.dmin[i] = 0xFFFFFFFFu; }
  if (tid == 0) { sm.cmd = 0; sm.g_len = 0; sm.g_complete = 0; sm.res_n = 0; sm.res_same = 0; sm.win_i = -1; sm.win_gen = -1; }

  // ---- per-node state: registers of the owner + shared mirror ----------------------------------
  double cd[NPT], ff[NPT], wd[NPT], wy[NPT];
  int32_t tot[NPT];
  uint32_t valid_bits = 0, boost_bits = 0;
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const int n = tid + (j << logTC);
    cd[j] = 0.0; ff[j] = 0.0; wd[j] = 1.0; wy[j] = 1.0; tot[j] = 0;
    if (!is_seq && n < N) {
      int t = extra[n];
      for (int s2 = 0; s2 < S; ++s2) t += counts[s2 * N + n];
      tot[j] = t;
      cd[j] = (double)counts[s * N + n];
      if (!pool.node_removed[D.nodeid_off + n]) valid_bits |= 1u << j;
      if (has_nw && pool.node_has_weight[D.node_off + n]) {
        const int w = pool.node_weight[D.node_off + n];
        if (w > 1) { wd[j] = (double)w; wy[j] = __ddiv_rn(1.0, wd[j]); }          // plan.go:678-679
        else if (w < 0 && D.booster == BLANCE_BOOSTER_CBGT_MAX) { boost_bits |= 1u << j; wd[j] = (double)w; }
      }                                                                            // w == 0 / no booster: untouched
      if (Pn > 0) ff[j] = div_exact(__dmul_rn(0.001, (double)t), Pd, Py);          // plan.go:650
      nd[4 * n] = cd[j]; nd[4 * n + 1] = ff[j]; nd[4 * n + 2] = wd[j]; nd[4 * n + 3] = wy[j];
      nd_flag[n] = (uint8_t)((((valid_bits >> j) & 1u) ? NF_VALID : 0u) | (((boost_bits >> j) & 1u) ? NF_BOOST : 0u));
    }
  }
  __syncthreads();   // constants, qtab (needed by the base keys) and the mirror are in place

  // A row blocks at most (its k current nodes + the nodes it holds in higher-priority states) entries of
  // the cached list, so that many + 2 entries are enough to always find an unblocked one.
  int glist_want = k + 2;
  for (int s2 = 0; s2 < S; ++s2)
    if ((higher_states >> s2) & 1u) glist_want += D.state_slot_off[s2 + 1] - D.state_slot_off[s2];
  if (glist_want > BL_GLIST) glist_want = BL_GLIST;
  unsigned long long Lk[NPT];               // base keys: n2n = 0, not current
#pragma unroll
  for (int j = 0; j < NPT; ++j)
    Lk[j] = key_from(cd[j], ff[j], wd[j], wy[j], (boost_bits >> j) & 1u, has_nw, 0, 0.0, qtab_a, Pd, Py);
  int xbuf = 0;

  if (!is_seq) {
    // =========================== compute warps: serve commands ====================================
    for (;;) {
      bar_sync(BAR_GO, NL);
      const int32_t cmd = *(volatile int32_t*)&sm.cmd;
      if (cmd == SEQ_CMD_EXIT) break;
      if (cmd == SEQ_CMD_REBUILD) {
        // BL_GLIST smallest (base key, position) over the live nodes, by repeated arg-min
        uint32_t listed = 0;
        int len = 0, complete = 0;
        for (int r = 0; r < glist_want; ++r) {
          unsigned long long bk = ~0ull;
          uint32_t bpos = 0xFFFFFFFFu;
#pragma unroll
          for (int j = 0; j < NPT; ++j)
            if (((valid_bits & ~listed) >> j) & 1u)
              if (bpos == 0xFFFFFFFFu || Lk[j] < bk) { bk = Lk[j]; bpos = (uint32_t)(tid + (j << logTC)); }
          const Best b = seq_argmin(Best{(uint32_t)(bk >> 32), (uint32_t)bk, bpos}, xchg_a, xbuf, cw, TC, warp, lane);
          if (b.pos == 0xFFFFFFFFu) { complete = 1; break; }
          if (tid == 0) sts128(glist_a + (uint32_t)r * 16u, b.hi, b.lo, b.pos, 0u);
          if ((b.pos & (uint32_t)(TC - 1)) == (uint32_t)tid) listed |= 1u << (b.pos >> logTC);
          ++len;
        }
        if (tid == 0) { sm.g_len = len; sm.g_complete = complete; }
        bar_sync(BAR_DONE, NL);
        continue;
      }
      // ---- full evaluation of step `cmd` (the lock-step kernel's step) -------------------------------
      const int i = cmd;
      const uint32_t reca = ring_a + ((uint32_t)i & rmask) * (SEQ_RSTRIDE * 4u);
      const int4 hdr = lds128(reca + (uint32_t)SLP * 4u);               // meta, w_p, top, partition
      const int32_t w_p = hdr.y, top = hdr.z;
      const double stick = lds64f(reca + (uint32_t)SLP * 4u + 16u);
      const int n_cur = lds32(reca + (uint32_t)(SLP + 6) * 4u);
      const bool row_clean = lds32(reca + (uint32_t)(SLP + 7) * 4u) != 0;
      int32_t qn[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const int n = tid + (j << logTC);
        qn[j] = (Pn > 0 && n < N) ? __ldcg(n2n + (size_t)top * N + n) : 0;
      }
      constexpr int MW = (NPT + 3) / 4;
      uint32_t mw[MW];
#pragma unroll
      for (int w = 0; w < MW; ++w) mw[w] = 0;
      for (int c = 0; c < (SLP >> 2); ++c) {
        const int4 v = lds128(reca + (uint32_t)c * 16u), bb = lds128(sbit_a + (uint32_t)c * 16u);
        const int32_t xs[4] = {v.x, v.y, v.z, v.w};
        const int32_t bs[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int32_t x = xs[u];
          const int jj = x >> logTC;
          const uint32_t val = ((x & (TC - 1)) == tid) ? ((uint32_t)bs[u] << ((jj & 3) * 8)) : 0u;
#pragma unroll
          for (int w = 0; w < MW; ++w) mw[w] |= ((jj >> 2) == w) ? val : 0u;
        }
      }
      uint32_t memb[NPT];
      unsigned long long key[NPT];
      uint32_t cand_bits = 0;
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        memb[j] = (mw[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
        const bool cand = ((valid_bits >> j) & 1u) && !(memb[j] & higher_states);   // plan.go:142-156
        const double cur = ((memb[j] >> s) & 1u) ? stick : 0.0;                    // plan.go:654-662
        const unsigned long long kk = key_from(cd[j], ff[j], wd[j], wy[j], (boost_bits >> j) & 1u, has_nw, qn[j], cur, qtab_a, Pd, Py);
        key[j] = cand ? kk : ~0ull;
        cand_bits |= (cand ? 1u : 0u) << j;
      }
      int n_chosen = 0;
      uint32_t taken_bits = 0;
      bool same = row_clean;
      while (n_chosen < k) {                                // the flat (score, position) order
        unsigned long long bk = ~0ull;
        uint32_t bpos = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < NPT; ++j)
          if ((((cand_bits & ~taken_bits) >> j) & 1u) && (bpos == 0xFFFFFFFFu || key[j] < bk)) { bk = key[j]; bpos = (uint32_t)(tid + (j << logTC)); }
        const uint32_t best = seq_argmin(Best{(uint32_t)(bk >> 32), (uint32_t)bk, bpos}, xchg_a, xbuf, cw, TC, warp, lane).pos;
        if (best == 0xFFFFFFFFu) break;
        if (tid == 0) sm.res_chosen[n_chosen] = (int32_t)best;
        ++n_chosen;
        if ((best & (uint32_t)(TC - 1)) == (uint32_t)tid) taken_bits |= 1u << (best >> logTC);
        bool hit = false;
        for (int q = 0; q < n_cur; ++q) hit = hit || ((uint32_t)lds32(reca + (uint32_t)(lo_s + q) * 4u) == best);
        same = same && hit;
      }
      same = same && (n_chosen == n_cur);
      if (tid == 0) { sm.res_n = n_chosen; sm.res_same = same ? 1 : 0; }
      // ---- apply (plan.go:238-245, 290-301) on the owners' registers and the mirror --------------------
      uint32_t touched = taken_bits;
#pragma unroll
      for (int w = 0; w < MW; ++w) touched |= mw[w];
      if (touched)
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const bool is_cur = (memb[j] >> s) & 1u, tk = (taken_bits >> j) & 1u;
        const int n = tid + (j << logTC);
        if ((is_cur || tk) && n < N) {
          int32_t t = tot[j];
          uint32_t dec = memb[j];
          const double wpd = (double)w_p;
          if ((dec >> s) & 1u) { cd[j] = __dsub_rn(cd[j], wpd); t -= w_p; dec &= ~(1u << s); }
          while (dec) {
            const int s2 = __ffs(dec) - 1;
            dec &= dec - 1;
            red_add(&counts[s2 * N + n], -w_p);
            t -= w_p;
          }
          if (tk) {
            cd[j] = __dadd_rn(cd[j], wpd);
            t += w_p;
            red_add(&n2n[(size_t)top * N + n], 1);
          }
          if (t != tot[j]) {
            tot[j] = t;
            if (Pn > 0) ff[j] = div_exact(__dmul_rn(0.001, (double)t), Pd, Py);
          }
          nd[4 * n] = cd[j]; nd[4 * n + 1] = ff[j];
          Lk[j] = key_from(cd[j], ff[j], wd[j], wy[j], (boost_bits >> j) & 1u, has_nw, 0, 0.0, qtab_a, Pd, Py);
        }
      }
      bar_sync(BAR_DONE, NL);
    }
    // ---- write the per-node counts of this state back ------------------------------------------------
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int n = tid + (j << logTC);
      if (n < N) counts[s * N + n] = __double2int_rn(cd[j]);
    }
    return;
  }

  // ================================ sequencer warps ====================================================
  // A sticky step changes no count, so consecutive sticky steps depend on each other only through
  // nodeToNodeCounts[top][node] (+1 per earlier step with the same top that kept the same node).  The
  // sequencer warps therefore evaluate a WINDOW of consecutive steps at once - one (step, current node)
  // item per lane, U items per lane, W warps - assuming every earlier step of the window is sticky too;
  // the window is committed up to the first step that is not sticky, which goes to the compute warps, and
  // the next window starts right after it.  SIMD across STEPS instead of across nodes.  Two steps of a
  // window that share (top, node) would see each other's increment: a small hash set in shared memory finds
  // them exactly, and the later one is treated as not sticky (it becomes the head of the next window).
  // The LEADER (first sequencer warp) owns the step counter, the record ring, the cache policy and the
  // hand-offs to the compute warps; the other sequencer warps only take part in windows.
  constexpr int WS = 32 / K;                    // steps per sub-window (one item per lane)
  constexpr int U = K >= 2 ? 2 : 1;             // sub-windows per warp: every lane carries U items (ILP)
  constexpr int WT = WS * U;                    // steps per warp per window (<= 32)
  const int sw = warp - cw;                     // 0 = leader
  const int NS = 32 * W;                        // sequencer threads
  const int AH = 32 * W;                        // >= W * WT: look-ahead unit of the ring
  const int wstep = lane / k, wq = lane - wstep * k, gb = wstep * k;
  const bool wlane = wstep < WS;
  auto wbar = [&](int id) { if (W > 1) bar_sync(id, NS); else __syncwarp(); };

  // leader: records stream into the ring with cp.async; `loaded` = first record not yet requested
  int loaded = 0;
  // 16-byte chunks: a record is CH = REC/4 chunks, so one warp instruction moves 32/CH records
  const int CH = REC >> 2, RPB = 32 / CH;       // chunks per record, records per warp pass
  const int my_rec = lane / CH, my_chunk = lane - my_rec * CH;
  auto request_records = [&](int upto) {       // request records [loaded, min(upto, n_assign)) as one group
    const int hi = upto < n_assign ? upto : n_assign;
    if (my_rec < RPB)
      for (int step = loaded + my_rec; step < hi; step += RPB) {
        const uint32_t dst = ring_a + ((uint32_t)step & rmask) * (SEQ_RSTRIDE * 4u) + (uint32_t)my_chunk * 16u;
        const int32_t* src = stream + (size_t)step * REC + my_chunk * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (hi > loaded) loaded = hi;
  };
  if (sw == 0) {
    request_records(2 * AH);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  const uint32_t blk_mask = higher_states | (1u << s);
  uint32_t slot_blocked = 0;                  // bit sl: a node in slot sl cannot be taken from the cached list
  for (int sl = 0; sl < SL && sl < 8; ++sl)
    if ((uint32_t)lds32(sbit_a + (uint32_t)sl * 4u) & blk_mask) slot_blocked |= 1u << sl;

  // the cached list is the same for every lane and changes only at a rebuild: its nodes stay in registers
  uint32_t gpos[BL_GLIST];
#pragma unroll
  for (int g = 0; g < BL_GLIST; ++g) gpos[g] = 0xFFFFFFFFu;
  int my_gen = -1;

  // leader state
  int g_len = 0, g_complete = 0, calm = BL_CALM_MIN, gen = 0;
  // How many quiet steps to wait before rebuilding the cache: none while rebuilds pay off (the
  // cache served at least 8 sticky steps before it was dropped), up to BL_CALM_MIN otherwise.
  int need_calm = 0, served = 0;
  long long n_fast = 0;
  int i = 0;
  bool force_full = false;                      // the last window ended at a step that is not sticky
#ifdef BLANCE_PASS_TIMING
  long long t_win = 0, n_win = 0, t_slow = 0, n_slow = 0, n_slow_same = 0, t_reb = 0, n_reb = 0, n_cut = 0, t0 = clock64();
  long long wp[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tl = 0;
#define WP(ix) do { const long long n_ = clock64(); wp[ix] += n_ - tl; tl = n_; } while (0)
#else
#define WP(ix) do { } while (0)
#endif

  for (;;) {
    if (sw == 0) {
      // ---- leader: steps that need the compute warps, until a window is due ------------------------------------
      bool window_due = false;
      while (i < n_assign) {
#ifdef BLANCE_PASS_TIMING
        t0 = clock64();
#endif
        // keep the ring >= 2*AH records ahead.  Invariant: loaded >= i + AH (a window is at most AH steps), so
        // the group requested here never holds a record of the current window and may stay in flight; once
        // nothing is left to request, everything must have landed.
        if (loaded < n_assign && loaded < i + 2 * AH) {
          request_records(i + 3 * AH);
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const uint32_t reca = ring_a + ((uint32_t)i & rmask) * (SEQ_RSTRIDE * 4u);
        const bool first_eligible = lds32(reca + (uint32_t)(SLP + 7) * 4u) != 0 && lds32(reca + (uint32_t)(SLP + 6) * 4u) == k;
        // step i itself decides whether the cache has to be rebuilt first
        if (first_eligible && g_len == 0 && calm >= need_calm) {
          served = 0;
          if (lane == 0) *(volatile int32_t*)&sm.cmd = SEQ_CMD_REBUILD;
          bar_sync(BAR_GO, NL);
          bar_sync(BAR_DONE, NL);
          g_len = *(volatile int32_t*)&sm.g_len;
          g_complete = *(volatile int32_t*)&sm.g_complete;
          ++gen;
#ifdef BLANCE_PASS_TIMING
          { long long t1 = clock64(); t_reb += t1 - t0; t0 = t1; ++n_reb; }
#endif
        }
        if (first_eligible && g_len > 0 && !force_full) { window_due = true; break; }
        force_full = false;
        // step i is not sticky (or there is no cache): full evaluation by the compute warps
        if (lane == 0) *(volatile int32_t*)&sm.cmd = i;
        bar_sync(BAR_GO, NL);
        bar_sync(BAR_DONE, NL);
        {
          const int n_chosen = *(volatile int32_t*)&sm.res_n;
          int32_t* orec = ostream + (size_t)i * REC;
          if (lane < k) orec[lane] = lane < n_chosen ? *(volatile int32_t*)&sm.res_chosen[lane] : BLANCE_NO_NODE;
          if (lane == 0) orec[k] = n_chosen;
          if (*(volatile int32_t*)&sm.res_same) { if (calm < (1 << 30)) ++calm; }
          else {
            if (g_len > 0) need_calm = served >= 8 ? 0 : (need_calm < BL_CALM_MIN ? need_calm + 1 : BL_CALM_MIN);
            calm = 0; g_len = 0;
          }
        }
        ++i;
#ifdef BLANCE_PASS_TIMING
        { long long t1 = clock64(); t_slow += t1 - t0; ++n_slow; n_slow_same += *(volatile int32_t*)&sm.res_same ? 1 : 0; }
#endif
      }
      if (lane == 0) {
        *(volatile int32_t*)&sm.win_i = window_due ? i : -1;
        *(volatile int32_t*)&sm.win_glen = g_len;
        *(volatile int32_t*)&sm.win_gcomplete = g_complete;
        *(volatile int32_t*)&sm.win_gen = gen;
      }
    }
#ifdef BLANCE_PASS_TIMING
    tl = t0;
#endif
    WP(0);
    wbar(BAR_W0);                               // the window command (and the leader's ring records) are visible
    WP(1);
    const int wi = *(volatile int32_t*)&sm.win_i;
    if (wi < 0) break;
    const int wg_len = *(volatile int32_t*)&sm.win_glen, wg_complete = *(volatile int32_t*)&sm.win_gcomplete;
    {
      const int wgen = *(volatile int32_t*)&sm.win_gen;
      if (wgen != my_gen) {
        my_gen = wgen;
#pragma unroll
        for (int g = 0; g < BL_GLIST; ++g) gpos[g] = (uint32_t)lds32(glist_a + (uint32_t)g * 16u + 8u);
      }
    }

    // ---- my U steps of the window (warp sw holds steps wi + sw*WT ...), my current node ------------------------
    int jst[U];
    bool eligible[U], winner[U];
    int32_t top[U], rowv[U][8], c[U], q[U];
    int4 ma[U], mb[U];
    uint32_t fl[U], slot[U];
    double stick[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      jst[u] = wi + sw * WT + u * WS + wstep;
      const bool live = wlane && jst[u] < n_assign;
      const uint32_t reca = ring_a + ((uint32_t)(live ? jst[u] : wi) & rmask) * (SEQ_RSTRIDE * 4u);
      const int4 r0 = lds128(reca);                                            // slots 0..3
      const int4 r1 = SLP > 4 ? lds128(reca + 16u) : make_int4(-1, -1, -1, -1);   // slots 4..7
      top[u] = lds32(reca + (uint32_t)(SLP + 2) * 4u);
      stick[u] = lds64f(reca + (uint32_t)SLP * 4u + 16u);
      const int n_cur = lds32(reca + (uint32_t)(SLP + 6) * 4u);
      const bool row_clean = lds32(reca + (uint32_t)(SLP + 7) * 4u) != 0;
      rowv[u][0] = r0.x; rowv[u][1] = r0.y; rowv[u][2] = r0.z; rowv[u][3] = r0.w;
      rowv[u][4] = r1.x; rowv[u][5] = r1.y; rowv[u][6] = r1.z; rowv[u][7] = r1.w;
      eligible[u] = live && row_clean && n_cur == k;
      c[u] = lds32(reca + (uint32_t)(lo_s + wq) * 4u);                         // my current node
      if (!eligible[u]) c[u] = -1;
      const int cc = c[u] < 0 ? 0 : c[u];       // row_clean: 0 <= c < N
      // n2n first (L2 latency, overlapped with everything up to the key), then the mirror
      q[u] = (eligible[u] && Pn > 0) ? __ldcg(n2n + (size_t)top[u] * N + cc) : 0;
      ma[u] = lds128(nd_a + (uint32_t)cc * 32u);
      mb[u] = lds128(nd_a + (uint32_t)cc * 32u + 16u);
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(fl[u]) : "r"(ndf_a + (uint32_t)cc));
      // enter (top, node) into the window's set; every member of a slot reports its item id
      winner[u] = false;
      slot[u] = 0;
      if (eligible[u]) {
        const unsigned long long mk = ((unsigned long long)(uint32_t)top[u] << 32) | (uint32_t)c[u];
        uint32_t h = (((uint32_t)top[u] * 0x9E3779B1u) ^ ((uint32_t)c[u] * 0x85EBCA77u)) >> (32 - SEQ_DTAB_LOG);
        for (;;) {
          const unsigned long long old = atomicCAS(&sm.dtab[h], ~0ull, mk);
          if (old == ~0ull) { winner[u] = true; break; }
          if (old == mk) break;
          h = (h + 1u) & ((1u << SEQ_DTAB_LOG) - 1u);
        }
        slot[u] = h;
        atomicMin(&sm.dmin[h], (uint32_t)((sw * U + u) * 32 + lane));
      }
    }
    // smallest cached base key among the nodes my row does not block (independent of q)
    bool g_found[U];
    unsigned long long gk[U];
    uint32_t gp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      g_found[u] = false; gk[u] = 0; gp[u] = 0;
      int gi = -1;
      int32_t bn[8];                            // the nodes this row blocks
#pragma unroll
      for (int sl = 0; sl < 8; ++sl) bn[sl] = ((slot_blocked >> sl) & 1u) ? rowv[u][sl] : -2;
#pragma unroll
      for (int g = BL_GLIST - 1; g >= 0; --g)
        if (g < wg_len) {
          bool blocked = false;
#pragma unroll
          for (int sl = 0; sl < 8; ++sl) blocked = blocked || (bn[sl] == (int32_t)gpos[g]);
          if (!blocked) gi = g;
        }
      if (gi >= 0) {
        const int4 e = lds128(glist_a + (uint32_t)gi * 16u);
        g_found[u] = true;
        gk[u] = ((unsigned long long)(uint32_t)e.x << 32) | (uint32_t)e.y;
        gp[u] = (uint32_t)e.z;
      }
    }
    WP(2);
    wbar(BAR_W1);                               // the set is complete
    WP(3);
    bool accept[U], held[U];                    // held: sticky, but an earlier step of the window holds one of my pairs
    int rank[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // an earlier step of this window holds my (top, node): my n2n count is one short, so I am not decided here
      const bool nonfirst = eligible[u] && *(volatile uint32_t*)&sm.dmin[slot[u]] != (uint32_t)((sw * U + u) * 32 + lane);
      const bool nf_step = ((__ballot_sync(0xFFFFFFFFu, nonfirst) >> gb) & ((1u << k) - 1u)) != 0;
      unsigned long long mykey = ~0ull;
      if (eligible[u])
        mykey = key_from(__hiloint2double(ma[u].y, ma[u].x), __hiloint2double(ma[u].w, ma[u].z), __hiloint2double(mb[u].y, mb[u].x),
                         __hiloint2double(mb[u].w, mb[u].z), (fl[u] & NF_BOOST) != 0, has_nw, q[u], stick[u], qtab_a, Pd, Py);
      const bool ok_self = eligible[u] && (fl[u] & NF_VALID) != 0;
      bool okl = true;                          // every current node of my step is a live candidate
      // the k (key, node) pairs of my step: worst key of the step, my rank inside it
      unsigned long long mxk = mykey;
      int32_t mxp = c[u];
      rank[u] = 0;
      if (K == 1) okl = ok_self;
      if (K == 2) {                             // the other lane of my pair
        const unsigned long long ok_ = __shfl_xor_sync(0xFFFFFFFFu, mykey, 1);
        const int32_t oc = __shfl_xor_sync(0xFFFFFFFFu, ok_self ? c[u] : -1, 1);
        okl = ok_self && oc >= 0;
        if (ok_ < mykey || (ok_ == mykey && oc < c[u])) rank[u] = 1;
        if (ok_ > mxk || (ok_ == mxk && oc > mxp)) { mxk = ok_; mxp = oc; }
      }
      if (K > 2)
#pragma unroll
      for (int t = 0; t < K; ++t) {
        const unsigned long long ok_ = __shfl_sync(0xFFFFFFFFu, mykey, gb + t);
        const int32_t oc = __shfl_sync(0xFFFFFFFFu, c[u], gb + t);
        const bool okt = __shfl_sync(0xFFFFFFFFu, (int)ok_self, gb + t) != 0;
        okl = okl && okt;
        if (t != wq) {
          if (ok_ < mykey || (ok_ == mykey && oc < c[u])) ++rank[u];
          if (ok_ > mxk || (ok_ == mxk && oc > mxp)) { mxk = ok_; mxp = oc; }
        }
      }
      bool sticky = false;
      if (okl) sticky = g_found[u] ? (mxk < gk[u] || (mxk == gk[u] && (uint32_t)mxp < gp[u]))
                                   : (wg_complete != 0);   // every other live node is ineligible for this partition
      accept[u] = sticky && !nf_step;
      held[u] = sticky && nf_step;
    }
    // leading sticky steps of my warp, then of the window
    {
      int a_w = 0, held_w = 0;                  // held_w: my first undecided step is merely held back
      bool open = true;                         // no undecided step so far
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t rej = __ballot_sync(0xFFFFFFFFu, wlane && !accept[u]);
        const uint32_t hb = __ballot_sync(0xFFFFFFFFu, held[u]);
        int a = (rej ? (__ffs(rej) - 1) : 32) / k;
        if (a > WS) a = WS;
        if (open) { a_w += a; if (a < WS) held_w = (int)((hb >> (a * k)) & 1u); }
        open = open && a == WS;
      }
      if (lane == 0) *(volatile int32_t*)&sm.win_acc[sw] = a_w | (held_w << 8);
    }
    WP(4);
    wbar(BAR_W2);
    WP(5);
    int n_acc = 0, next_held = 0;
    {
      bool open = true;
      for (int w2 = 0; w2 < W; ++w2) {
        const int v = *(volatile int32_t*)&sm.win_acc[w2], a = v & 0xFF;
        if (open) { n_acc += a; if (a < WT) next_held = v >> 8; }
        open = open && a == WT;
      }
    }
    // commit the leading run of sticky steps; empty my slots of the set
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (wlane && sw * WT + u * WS + wstep < n_acc) {
        red_add(&n2n[(size_t)top[u] * N + c[u]], 1);               // plan.go:238-245
        int32_t* orec = ostream + (size_t)jst[u] * REC;
        orec[rank[u]] = c[u];                                      // ordered by (score, position)
        if (wq == 0) orec[k] = k;
      }
      if (winner[u]) { sm.dtab[slot[u]] = ~0ull; sm.dmin[slot[u]] = 0xFFFFFFFFu; }
    }
    // a hand-off to the compute warps may follow (even after a whole sticky window, when the next step is not
    // eligible): their n2n loads must see the increments of every sequencer warp
    WP(6);
    wbar(BAR_W3);
    WP(7);
    if (sw == 0) {
      n_fast += n_acc;
      served += n_acc;
      calm = calm + n_acc < (1 << 30) ? calm + n_acc : (1 << 30);
      i += n_acc;
#ifdef BLANCE_PASS_TIMING
      { long long t1 = clock64(); t_win += t1 - t0; ++n_win; n_cut += (n_acc < W * WT && next_held) ? 1 : 0; }
#endif
      // the step after the run is not sticky and goes to the compute warps - unless it was only held back
      // by a repeated (top, node): then it heads the next window
      force_full = n_acc < W * WT && !next_held;
    }
  }
#ifdef BLANCE_PASS_TIMING
  if (sw == 0 && lane == 0 && blockIdx.x == 0)
    printf("   leader phases (cycles per window): pre %.0f | W0 %.0f | loads+set+scan %.0f | W1 %.0f | keys+accept %.0f | W2 %.0f | commit %.0f | W3 %.0f\n",
           (double)wp[0] / (n_win ? n_win : 1), (double)wp[1] / (n_win ? n_win : 1), (double)wp[2] / (n_win ? n_win : 1), (double)wp[3] / (n_win ? n_win : 1),
           (double)wp[4] / (n_win ? n_win : 1), (double)wp[5] / (n_win ? n_win : 1), (double)wp[6] / (n_win ? n_win : 1), (double)wp[7] / (n_win ? n_win : 1));
  if (sw == 0 && lane == 0 && blockIdx.x == 0)
    printf("seq pass s=%d steps %d W=%d: sticky %lld in %lld windows (%.0f cyc/window, %.1f steps/window); full %lld (%.0f cyc each, %lld kept the row); rebuilds %lld (%.0f cyc each); %lld windows ended by a repeated pair\n",
           s, n_assign, W, n_fast, n_win, n_win ? (double)t_win / n_win : 0.0, n_win ? (double)n_fast / n_win : 0.0, n_slow,
           n_slow ? (double)t_slow / n_slow : 0.0, n_slow_same, n_reb, n_reb ? (double)t_reb / n_reb : 0.0, n_cut);
#endif
  if (sw == 0) {
    if (lane == 0) *(volatile int32_t*)&sm.cmd = SEQ_CMD_EXIT;
    bar_sync(BAR_GO, NL);
    if (lane == 0) { D.steps += n_assign; D.fast_steps += n_fast; }
  }
}

}  // namespace blance_dev
