// blance_b200/csrc/assign_pass_spec.cuh — the speculative assign pass: SCOUT warps decide steps ahead of
// time from a snapshot of the node state, one LEADER warp validates and commits them in the reference's
// order and resolves the steps that move data on its own.
//
// Same chain as assign_pass.cuh / assign_pass_seq.cuh (assignStateToPartitions + findBestNodes,
// plan.go:98-303), same results.  What is different is who does the work:
//
//   * The per-node score inputs {count of the state being assigned, filled term, weight, 1/weight} live in a
//     shared-memory MIRROR that only the leader (or the whole team, on its behalf) updates.  Every update
//     belongs to an EPOCH: the leader bumps `epoch` after the update is visible and stamps the touched nodes
//     in lastchg[].
//   * The pass's step records stream into a shared-memory RING by TMA (cp.async.bulk + mbarrier), one bulk
//     copy per 32-step chunk, issued by the scout warp that owns the chunk (chunk c belongs to scout
//     c mod SW; every ring slot is only ever written by its owner).
//   * SCOUTS sweep their chunks of the ring over and over, nearest to the leader first.  For a step whose
//     partition holds exactly K clean current nodes a scout computes the K exact scores (plan.go:634-689)
//     with  nodeToNodeCounts[top][c] = qstat + A[top][c]:  qstat = the number of earlier such steps of the
//     pass with the same (top, c) pair (k_pair_rank, computed before the pass with a sort) and A = the
//     deviation from that hypothesis, a matrix that only non-sticky outcomes touch (and touching it stamps
//     the node).  It publishes {worst score T, its node, the (score, position) ranks of the K nodes, the
//     epoch it read BEFORE reading anything else} as one 16-byte word.
//   * The LEADER walks the steps 32 at a time: a published result is still exact iff none of the step's
//     current nodes was stamped after the result's epoch; the step is then STICKY - the reference's sort
//     puts exactly the current nodes first - iff T is below B0, the smallest base key (score with
//     nodeToNodeCounts = 0 and no stickiness, a lower bound of every other candidate's score; all operations
//     of plan.go:634-689 are monotone) over all live nodes.  Sticky steps change no count: the leader commits
//     the leading run of accepted steps with two fire-and-forget stores per step.
//   * The first step that is not accepted is RESOLVED by the leader alone: it keeps the smallest base keys
//     (up to 64, two per lane, unsorted) and a lower bound `ub` of every base key that is not listed.  The
//     exact scores of the listed nodes and of the partition's current nodes, K warp arg-mins, and the proof
//     that the K-th winner is below `ub` give the reference's first K nodes without looking at the other
//     nodes.  A mover updates the mirror, the matrices, the list and the epoch - about a microsecond,
//     against a CTA-wide N-way arg-min per pick in the other kernels.
//   * Rows that are not clean, a failed proof and a list that has run dry go to the TEAM: the scouts stop
//     sweeping and run the full evaluation / the list rebuild together (named barriers), exactly as the
//     lock-step kernel would.
//
// Exactness never depends on timing: a result is only used under the two tests above, anything else falls
// through to an exact evaluation.  tools/spec_model.c checks these rules step by step inside the CPU oracle.
#pragma once

#include "assign_pass_seq.cuh"

namespace blance_dev {

#define SP_LPL 2             // list entries per leader lane
#define SP_D 2               // ring chunks per scout warp
#define SP_NPTS 8            // nodes per scout thread in team operations (N <= 32 * SW * SP_NPTS)

enum : int { SPB_ALL = 8, SPB_GO = 9, SPB_DONE = 10, SPB_TEAM = 11 };
enum : int32_t { SP_OP_EXIT = 1, SP_OP_FULL = 3 };
enum : uint32_t { SPZ_NEVER = 0x80000000u };

struct SpecCtl {
  uint4 xchg[2][32];                 // team arg-min partials
  uint4 pubq[64];                    // leader -> publisher: {A offset or -1, delta, epoch, 1 = last entry of its epoch}
  alignas(16) int32_t slot_bit[BL_SLP_MAX];
  alignas(8) unsigned long long mbar[32 * SP_D];
  int32_t epoch, front, cmd_seq, cmd_op, cmd_arg, cmd_epoch;
  int32_t pub_head, pub_done, pub_quit, commit_done;
  int32_t abort_flag;                // a wait loop gave up (never in a correct run): every role leaves, the host reports an error
  int32_t res_n, res_same;
  int32_t res_chosen[BL_K_MAX];
};

__host__ __device__ inline size_t spec_dyn_smem_bytes(int N, int SW) {
  const size_t H = (size_t)32 * SW * SP_D;
  const size_t Np = ((size_t)N + 3) & ~(size_t)3;        // every array starts 16-byte aligned
  size_t b = Np * 32 + Np * 8 + Np * 4 + Np * 4 + ((Np + 15) & ~(size_t)15);   // mirror, base keys, totals, stamps, flags
  b += H * 64 + H * 16 + H * 16 + H;   // records (<= 16 words), qstat, results, accepted ranks
  return b;
}

__device__ __forceinline__ int32_t ld_relaxed_gpu(const int32_t* p) {
  int32_t v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int32_t lds32v(uint32_t a) {
  int32_t v;
  asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32v(uint32_t a, int32_t x) {
  asm volatile("st.volatile.shared.b32 [%0], %1;" :: "r"(a), "r"(x) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t a, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  return ok != 0;
}

// exact key (plan.go:634-689) with an explicit nodeToNodeCounts value q
__device__ __forceinline__ unsigned long long sp_key(double cd, double ff, double wd, double wy, bool boost, bool has_nw,
                                                     int32_t q, double cur, bool have_p, double Pd, double Py) {
  const double qv = have_p ? div_exact((double)q, Pd, Py) : 0.0;                          // plan.go:641-642
  const double base = __dadd_rn(__dadd_rn(cd, qv), ff);                                   // plan.go:672-673
  double r = base;
  if (has_nw && !boost) r = div_exact(r, wd, wy);                                         // plan.go:679
  if (boost) {                                                                            // plan.go:680-681
    double b = -wd;
    if (b < cur) b = cur;
    r = __dadd_rn(base, b);
  }
  r = __dsub_rn(r, cur);                                                                  // plan.go:686
  return score_key(r);
}

// warp arg-min of (hi, lo, pos): one redux when the high words already decide it
__device__ __forceinline__ Best warp_argmin_q(Best v) {
  const unsigned full = 0xFFFFFFFFu;
  const uint32_t mhi = __reduce_min_sync(full, v.hi);
  const uint32_t eq = __ballot_sync(full, v.hi == mhi);
  if ((eq & (eq - 1u)) == 0u) {
    const int src = __ffs(eq) - 1;
    return Best{mhi, __shfl_sync(full, v.lo, src), __shfl_sync(full, v.pos, src)};
  }
  const uint32_t lo2 = (v.hi == mhi) ? v.lo : 0xFFFFFFFFu;
  const uint32_t mlo = __reduce_min_sync(full, lo2);
  const uint32_t p2 = (lo2 == mlo && v.hi == mhi) ? v.pos : 0xFFFFFFFFu;
  return Best{mhi, mlo, __reduce_min_sync(full, p2)};
}

// base key: the score with nodeToNodeCounts = 0 and no stickiness (cd + 0/P == cd exactly)
__device__ __forceinline__ unsigned long long sp_base_key(double cd, double ff, double wd, double wy, bool boost, bool has_nw) {
  const double base = __dadd_rn(cd, ff);
  double r = base;
  if (has_nw && !boost) r = div_exact(r, wd, wy);
  if (boost) { double b = -wd; if (b < 0.0) b = 0.0; r = __dadd_rn(base, b); }
  return score_key(r);
}

__device__ __forceinline__ bool lex_lt(unsigned long long ka, uint32_t pa, unsigned long long kb, uint32_t pb) {
  return ka < kb || (ka == kb && pa < pb);
}

// K = the state's constraints (1..BL_FAST_K).  blockDim.x = 32 * NW warps; warp 0 is the leader, the warps
// of idle_mask exit at once (they keep the leader's scheduler free), the first of the others is the PUBLISHER
// (it applies the movers' updates of A and publishes the epochs, so that the leader never waits for a fence
// over global atomics), the rest are the SW scouts.  SW * SP_D must be a power of two (swd_shift = its log2).
template <int K>
__global__ void __launch_bounds__(448, 1) k_assign_pass_spec(DPool pool, int s, int SW, unsigned idle_mask, int swd_shift) {
  DInst& D = pool.insts[blockIdx.x];
  if (!D.active || s >= D.S || D.pass_mode != 2) return;
  if (D.state_constraints[s] != K) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((idle_mask >> warp) & 1u) return;
  const bool is_leader = warp == 0;
  const int arank = __popc(~idle_mask & ((1u << warp) - 1u));           // rank among the active warps: 0 leader, 1 publisher, 2 committer
  const bool is_pub = arank == 1, is_com = arank == 2;
  const int sidx = arank - 3;                                           // scout number
  const int NTA = 32 * (SW + 3), TS = 32 * SW, NTT = 32 * (SW + 1);     // all active threads / scouts / scouts + leader
  const int atid = arank * 32 + lane;                                   // rank among the active threads
  const int SWD = SW * SP_D, H = 32 * SWD;
  const uint32_t swd_mask = (uint32_t)SWD - 1u;

  __shared__ SpecCtl ctl;
  extern __shared__ __align__(16) unsigned char dyn_smem[];

  const int N = D.N, S = D.S, SL = D.SL, SLP = D.SLP;
  const int n_assign = D.n_assign;
  const int lo_s = D.state_slot_off[s];
  const int Pn = D.P;
  const bool have_p = Pn > 0;
  const double Pd = have_p ? (double)Pn : 1.0;
  const double Py = __ddiv_rn(1.0, Pd);
  const bool has_nw = D.has_node_weights != 0;
  uint32_t higher_states = 0;              // bit s2: priority[s2] < priority[s]  (plan.go:146-152)
  for (int s2 = 0; s2 < S; ++s2)
    if (D.state_priority[s2] < D.state_priority[s]) higher_states |= 1u << s2;

  const int REC = SLP + BL_REC_HDR;
  const uint32_t RECB = (uint32_t)REC * 4u;
  const int32_t* stream = pool.stream + D.stream_off;
  int32_t* ostream = pool.ostream + D.stream_off;
  const int32_t* qstat_g = pool.qstat + D.part_off * 4;
  int32_t* counts = pool.counts + D.counts_off;
  int32_t* G = pool.n2n + D.n2n_off;                 // nodeToNodeCounts (plan.go:266)
  int32_t* A = pool.n2n_dev + D.n2n_off;             // its deviation from the all-sticky hypothesis
  const int32_t* extra = (D.use_rest ? pool.extra_rest : pool.extra_first) + D.node_off;

  // ---- shared memory map -------------------------------------------------------------------------------
  const uint32_t base_a = (uint32_t)__cvta_generic_to_shared(dyn_smem);
  const uint32_t Np = ((uint32_t)N + 3u) & ~3u;                  // (padded: every array starts 16-byte aligned)
  const uint32_t nd_a = base_a;                                  // mirror: {cd, ff, wd, wy} per node
  const uint32_t bk_a = nd_a + 32u * Np;                         // base key of every node (score with n2n = 0, no stickiness)
  const uint32_t tot_a = bk_a + 8u * Np;                         // all-state totals
  const uint32_t chg_a = tot_a + 4u * Np;                        // lastchg
  const uint32_t flg_a = chg_a + 4u * Np;                        // NF_VALID | NF_BOOST
  const uint32_t rec_a = flg_a + ((Np + 15u) & ~15u);            // ring: records
  const uint32_t qs_a = rec_a + (uint32_t)H * RECB;              // ring: qstat (4 per step)
  const uint32_t dyn_a = qs_a + (uint32_t)H * 16u;               // ring: scout results
  const uint32_t acc_a = dyn_a + (uint32_t)H * 16u;              // ring: what the leader decided (0x80 | ranks: sticky; 1: resolved)
  const uint32_t ctl_a = (uint32_t)__cvta_generic_to_shared(&ctl);
  const uint32_t xchg_a = ctl_a + (uint32_t)offsetof(SpecCtl, xchg);
  const uint32_t sbit_a = ctl_a + (uint32_t)offsetof(SpecCtl, slot_bit);
  const uint32_t mbar_a = ctl_a + (uint32_t)offsetof(SpecCtl, mbar);
  const uint32_t epoch_a = ctl_a + (uint32_t)offsetof(SpecCtl, epoch);
  const uint32_t front_a = ctl_a + (uint32_t)offsetof(SpecCtl, front);
  const uint32_t seq_a = ctl_a + (uint32_t)offsetof(SpecCtl, cmd_seq);
  const uint32_t pubq_a = ctl_a + (uint32_t)offsetof(SpecCtl, pubq);
  const uint32_t pubh_a = ctl_a + (uint32_t)offsetof(SpecCtl, pub_head);
  const uint32_t pubd_a = ctl_a + (uint32_t)offsetof(SpecCtl, pub_done);
  const uint32_t pubx_a = ctl_a + (uint32_t)offsetof(SpecCtl, pub_quit);
  const uint32_t cdone_a = ctl_a + (uint32_t)offsetof(SpecCtl, commit_done);
  const uint32_t abort_a = ctl_a + (uint32_t)offsetof(SpecCtl, abort_flag);

  // ---- pass constants, mirror, ring ------------------------------------------------------------------
  for (int i = atid; i < SLP; i += NTA) {
    int st = 0;
    while (st + 1 < S && i >= D.state_slot_off[st + 1]) ++st;
    ctl.slot_bit[i] = (i < SL) ? (1 << st) : 0;
  }
  for (int n = atid; n < N; n += NTA) {
    int t = extra[n];
    for (int s2 = 0; s2 < S; ++s2) t += counts[s2 * N + n];
    const double cd = (double)counts[s * N + n];
    double wd = 1.0, wy = 1.0;
    uint32_t fl = pool.node_removed[D.nodeid_off + n] ? 0u : NF_VALID;
    if (has_nw && pool.node_has_weight[D.node_off + n]) {
      const int w = pool.node_weight[D.node_off + n];
      if (w > 1) { wd = (double)w; wy = __ddiv_rn(1.0, wd); }                          // plan.go:678-679
      else if (w < 0 && D.booster == BLANCE_BOOSTER_CBGT_MAX) { fl |= NF_BOOST; wd = (double)w; }
    }
    const double ff = have_p ? div_exact(__dmul_rn(0.001, (double)t), Pd, Py) : 0.0;   // plan.go:650
    double* nd = reinterpret_cast<double*>(dyn_smem) + 4 * (size_t)n;
    nd[0] = cd; nd[1] = ff; nd[2] = wd; nd[3] = wy;
    *reinterpret_cast<unsigned long long*>(dyn_smem + (bk_a - base_a) + 8u * n) = sp_base_key(cd, ff, wd, wy, (fl & NF_BOOST) != 0, has_nw);
    sts32(tot_a + 4u * n, t);
    sts32(chg_a + 4u * n, 0);
    dyn_smem[(flg_a - base_a) + n] = (unsigned char)fl;
  }
  for (int i = atid; i < H; i += NTA) sts128(dyn_a + 16u * i, 0u, 0u, 0x3FFu << 21, 0u);      // generation 1023: never written
  if (atid == 0) {
    for (int i = 0; i < SWD; ++i) mbar_init(mbar_a + 8u * i, 1);
    ctl.epoch = 0; ctl.front = 0; ctl.cmd_seq = 0; ctl.cmd_op = 0; ctl.cmd_arg = 0; ctl.cmd_epoch = 0;
    ctl.pub_head = 0; ctl.pub_done = 0; ctl.pub_quit = 0; ctl.commit_done = 0; ctl.abort_flag = 0;
    ctl.res_n = 0; ctl.res_same = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  bar_sync(SPB_ALL, NTA);

  uint32_t slot_blocked = 0;                 // bit sl: slot sl belongs to a higher-priority state (its node is not a candidate)
  uint32_t slot_state_s = 0;                 //         slot sl belongs to the state being assigned
  uint32_t sbit8[8];                         // state bit of slot sl
#pragma unroll
  for (int sl = 0; sl < 8; ++sl) sbit8[sl] = sl < SL ? (uint32_t)lds32(sbit_a + 4u * sl) : 0u;
  for (int sl = 0; sl < SL && sl < 8; ++sl) {
    const uint32_t b = (uint32_t)lds32(sbit_a + 4u * sl);
    if (b & higher_states) slot_blocked |= 1u << sl;
    if (b & (1u << s)) slot_state_s |= 1u << sl;
  }

  // node owned by (scout, lane, j) in team operations: consecutive ids are spread over the warps
  auto team_node = [&](int j) { return sidx + SW * (lane + 32 * j); };
  int xbuf = 0;
  auto team_argmin = [&](Best mine) {
    const Best w = warp_argmin(mine);
    const uint32_t b = xchg_a + (uint32_t)xbuf * 512u;
    if (lane == 0) sts128(b + (uint32_t)sidx * 16u, w.hi, w.lo, w.pos, 0u);
    bar_sync(SPB_TEAM, TS);
    int4 e = make_int4(-1, -1, -1, 0);
    if (lane < SW) e = lds128(b + (uint32_t)lane * 16u);
    xbuf ^= 1;
    return warp_argmin(Best{(uint32_t)e.x, (uint32_t)e.y, (uint32_t)e.z});
  };

  if (is_pub) {
    // =================================== publisher =====================================================
    // Applies the queued A updates (fire-and-forget atomics), fences them, then publishes the epoch of the last
    // complete group: a scout that reads epoch e afterwards also sees every update that belongs to epochs <= e.
    int tail = 0;
    for (;;) {
      const int head = lds32v(pubh_a);
      if (head == tail) {
        if (lds32v(pubx_a)) break;
        __nanosleep(40);
        continue;
      }
      const int cnt = (head - tail) < 32 ? (head - tail) : 32;
      int32_t ep = -1;
      if (lane < cnt) {
        const int4 q = lds128(pubq_a + 16u * (uint32_t)((tail + lane) & 63));
        if (q.x >= 0) atomicAdd(&A[q.x], q.y);
        if (q.w) ep = q.z;
      }
      __threadfence_block();
      ep = __reduce_max_sync(0xFFFFFFFFu, ep);
      tail += cnt;
      if (lane == 0) {
        if (ep >= 0) sts32v(epoch_a, ep);
        sts32v(pubd_a, tail);
      }
      __syncwarp();
    }
    return;
  }
  if (is_com) {
    // =================================== committer =====================================================
    // Trails the leader: for every step the leader accepted as sticky it bumps nodeToNodeCounts (plan.go:238-245)
    // and stores the outcome (the ranks of the current nodes); the leader waits for commit_done before it reads
    // nodeToNodeCounts in a resolve.  Keeps the scattered global atomics out of the leader's instruction stream.
    uint8_t* srank = pool.srank + D.part_off;
    int c = 0;
    for (;;) {
      const int f = lds32v(front_a);
      if (f == c) {
        if (lds32v(pubx_a)) break;
        __nanosleep(20);
        continue;
      }
      while (c < f) {
        const int j = c + lane;
        if (j < f) {
          const uint32_t cj = (uint32_t)j >> 5;
          const uint32_t slot = ((cj & swd_mask) << 5) | ((uint32_t)j & 31u);
          const uint32_t a = dyn_smem[(acc_a - base_a) + slot];
          if (a & 0x80u) {
            const uint32_t reca = rec_a + slot * RECB;
            if (have_p) {
              const int32_t top = lds32(reca + (uint32_t)(SLP + 2) * 4u);
#pragma unroll
              for (int q = 0; q < K; ++q) atomicAdd(&G[(size_t)top * N + lds32(reca + (uint32_t)(lo_s + q) * 4u)], 1);
            }
            srank[j] = (uint8_t)a;
          }
        }
        c = (c + 32 < f) ? c + 32 : f;
      }
      __syncwarp();
      if (lane == 0) sts32v(cdone_a, c);
    }
    return;
  }
  if (!is_leader) {
    // =================================== scouts ========================================================
    int chunk[SP_D];
    bool loaded[SP_D];
    int32_t seen_epoch[SP_D];                             // epoch at the last check that found slot d up to date (-1: none)
    uint32_t phase = 0;                                   // bit d: parity to wait for on slot d
    auto issue_load = [&](int d, int c) {
      if (lane == 0) {
        const int first = c * 32;
        const int nrec = (n_assign - first) < 32 ? (n_assign - first) : 32;
        const uint32_t slot0 = (uint32_t)(sidx + SW * d) * 32u;
        const uint32_t mb = mbar_a + 8u * (uint32_t)(sidx + SW * d);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(mb, (uint32_t)nrec * (RECB + 16u));
        bulk_g2s(rec_a + slot0 * RECB, stream + (size_t)first * REC, (uint32_t)nrec * RECB, mb);
        bulk_g2s(qs_a + slot0 * 16u, qstat_g + (size_t)first * 4, (uint32_t)nrec * 16u, mb);
      }
    };
#pragma unroll
    for (int d = 0; d < SP_D; ++d) {
      chunk[d] = sidx + SW * d;
      loaded[d] = false;
      seen_epoch[d] = -1;
      if (chunk[d] * 32 < n_assign) issue_load(d, chunk[d]);
    }
    int my_seq = 0;
    for (;;) {
      bool any_work = false, busy = false;
#pragma unroll
      for (int d = 0; d < SP_D; ++d) {
        // ---- a command of the leader? ----------------------------------------------------------------
        const int seq = lds32v(seq_a);
        if (seq != my_seq) {
          my_seq = seq;
          bar_sync(SPB_GO, NTT);
          const int op = *(volatile int32_t*)&ctl.cmd_op;
          if (op == SP_OP_EXIT) goto scouts_done;
          bool changed = false;
          const int32_t E1 = *(volatile int32_t*)&ctl.cmd_epoch;
          if (op == SP_OP_FULL) {
            // ---- full evaluation of step cmd_arg (the lock-step kernel's step) from the mirror --------------
            const int i = *(volatile int32_t*)&ctl.cmd_arg;
            const int ci = i >> 5;
            const uint32_t slot = (((uint32_t)ci & swd_mask) << 5) | (uint32_t)(i & 31);
            const uint32_t reca = rec_a + slot * RECB;
            const int4 hdr = lds128(reca + (uint32_t)SLP * 4u);               // meta, w_p, top, partition
            const int32_t w_p = hdr.y, top = hdr.z;
            const double stick = lds64f(reca + (uint32_t)SLP * 4u + 16u);
            const int n_cur = lds32(reca + (uint32_t)(SLP + 6) * 4u);
            const bool row_clean = lds32(reca + (uint32_t)(SLP + 7) * 4u) != 0;
            const bool elig = row_clean && n_cur == K;
            uint32_t memb[SP_NPTS];
            unsigned long long key[SP_NPTS];
            uint32_t cand_bits = 0, taken_bits = 0;
#pragma unroll
            for (int j = 0; j < SP_NPTS; ++j) {
              const int n = team_node(j);
              memb[j] = 0;
              key[j] = ~0ull;
              if (n < N) {
                for (int sl = 0; sl < SLP; ++sl)
                  if (lds32(reca + 4u * sl) == n) memb[j] |= (uint32_t)lds32(sbit_a + 4u * sl);
                const int4 ma = lds128(nd_a + (uint32_t)n * 32u), mb = lds128(nd_a + (uint32_t)n * 32u + 16u);
                const uint32_t fl = dyn_smem[(flg_a - base_a) + n];
                const int32_t q = have_p ? ld_relaxed_gpu(G + (size_t)top * N + n) : 0;
                const bool cand = (fl & NF_VALID) && !(memb[j] & higher_states);           // plan.go:142-156
                const double cur = ((memb[j] >> s) & 1u) ? stick : 0.0;                    // plan.go:654-662
                if (cand) {
                  key[j] = sp_key(__hiloint2double(ma.y, ma.x), __hiloint2double(ma.w, ma.z), __hiloint2double(mb.y, mb.x),
                                  __hiloint2double(mb.w, mb.z), (fl & NF_BOOST) != 0, has_nw, q, cur, have_p, Pd, Py);
                  cand_bits |= 1u << j;
                }
              }
            }
            int n_chosen = 0;
            bool same = row_clean;
            while (n_chosen < K) {                                // the flat (score, position) order
              unsigned long long bk = ~0ull;
              uint32_t bpos = 0xFFFFFFFFu;
#pragma unroll
              for (int j = 0; j < SP_NPTS; ++j)
                if ((((cand_bits & ~taken_bits) >> j) & 1u) && (bpos == 0xFFFFFFFFu || key[j] < bk)) { bk = key[j]; bpos = (uint32_t)team_node(j); }
              const uint32_t best = team_argmin(Best{(uint32_t)(bk >> 32), (uint32_t)bk, bpos}).pos;
              if (best == 0xFFFFFFFFu) break;
              if (sidx == 0 && lane == 0) ctl.res_chosen[n_chosen] = (int32_t)best;
              ++n_chosen;
#pragma unroll
              for (int j = 0; j < SP_NPTS; ++j)
                if ((uint32_t)team_node(j) == best) taken_bits |= 1u << j;
              bool hit = false;
              for (int q = 0; q < n_cur; ++q) hit = hit || ((uint32_t)lds32(reca + (uint32_t)(lo_s + q) * 4u) == best);
              same = same && hit;
            }
            same = same && (n_chosen == n_cur) && elig;      // (a kept short row still moves A: the hypothesis did not count it)
            if (sidx == 0 && lane == 0) { ctl.res_n = n_chosen; ctl.res_same = same ? 1 : 0; }
            changed = !same;
            // ---- apply (plan.go:238-245, 290-301) on the mirror; owners only --------------------------------
#pragma unroll
            for (int j = 0; j < SP_NPTS; ++j) {
              const int n = team_node(j);
              const bool is_cur = (memb[j] >> s) & 1u, tk = (taken_bits >> j) & 1u;
              if (n < N && (is_cur || tk)) {
                if (tk) atomicAdd(&G[(size_t)top * N + n], 1);
                if (changed) {
                  const int dA = (tk ? 1 : 0) - ((elig && is_cur) ? 1 : 0);
                  if (dA) atomicAdd(&A[(size_t)top * N + n], dA);
                  double* nd = reinterpret_cast<double*>(dyn_smem) + 4 * (size_t)n;
                  double cd = nd[0];
                  int32_t t0 = lds32(tot_a + 4u * n), t = t0;
                  uint32_t dec = memb[j];
                  const double wpd = (double)w_p;
                  if ((dec >> s) & 1u) { cd = __dsub_rn(cd, wpd); t -= w_p; dec &= ~(1u << s); }
                  while (dec) {
                    const int s2 = __ffs(dec) - 1;
                    dec &= dec - 1;
                    atomicSub(&counts[s2 * N + n], w_p);
                    t -= w_p;
                  }
                  if (tk) { cd = __dadd_rn(cd, wpd); t += w_p; }
                  nd[0] = cd;
                  if (t != t0) {
                    sts32(tot_a + 4u * n, t);
                    if (have_p) nd[1] = div_exact(__dmul_rn(0.001, (double)t), Pd, Py);
                  }
                  sts32(chg_a + 4u * n, E1);
                  *reinterpret_cast<unsigned long long*>(dyn_smem + (bk_a - base_a) + 8u * n) =
                      sp_base_key(cd, nd[1], nd[2], nd[3], (dyn_smem[(flg_a - base_a) + n] & NF_BOOST) != 0, has_nw);
                }
              }
            }
          }
          if (changed) __threadfence_block();         // my atomics on A are ordered before the epoch the leader queues
          bar_sync(SPB_DONE, NTT);
        }
        // ---- my chunk of ring slot d -------------------------------------------------------------------
        int c = chunk[d];
        if (c * 32 >= n_assign) continue;
        any_work = true;
        const int fr = lds32v(cdone_a);
        if (fr >= (c + 1) * 32) {                       // consumed (by the leader AND the committer): the slot takes its next chunk
          c += SWD;
          chunk[d] = c;
          loaded[d] = false;
          seen_epoch[d] = -1;
          if (c * 32 < n_assign) issue_load(d, c);
          continue;
        }
        const uint32_t mb = mbar_a + 8u * (uint32_t)(sidx + SW * d);
        if (!loaded[d]) {
          if (!mbar_test(mb, (phase >> d) & 1u)) continue;
          loaded[d] = true;
          phase ^= 1u << d;
        }
        // ---- evaluate the chunk if one of its results is missing or out of date -----------------------------
        // (nothing is stamped without an epoch bump, so a chunk found up to date stays so until the epoch moves)
        const int32_t e0 = lds32v(epoch_a);
        if (e0 == seen_epoch[d]) continue;
        const uint32_t slot = (uint32_t)(sidx + SW * d) * 32u + (uint32_t)lane;
        const int j = c * 32 + lane;
        const uint32_t gen = ((uint32_t)c >> swd_shift) & 0x3FFu;
        const uint32_t reca = rec_a + slot * RECB;
        const bool live = j < n_assign;
        const int n_cur = lds32(reca + (uint32_t)(SLP + 6) * 4u);
        const bool clean = lds32(reca + (uint32_t)(SLP + 7) * 4u) != 0;
        const bool elig = live && clean && n_cur == K;
        int32_t cn[K];
#pragma unroll
        for (int q = 0; q < K; ++q) { cn[q] = lds32(reca + (uint32_t)(lo_s + q) * 4u); if (!elig) cn[q] = 0; }
        {
          const int4 old = lds128(dyn_a + slot * 16u);
          bool valid = (((uint32_t)old.z >> 21) & 0x3FFu) == gen;
          if (valid && elig) {
#pragma unroll
            for (int q = 0; q < K; ++q) valid = valid && lds32v(chg_a + 4u * (uint32_t)cn[q]) <= old.w;
          }
          if (!__any_sync(0xFFFFFFFFu, live && !valid)) { seen_epoch[d] = e0; continue; }
        }
        busy = true;
        const int32_t e = lds32v(epoch_a);
        __threadfence_block();                          // everything below is read after the epoch
        uint32_t z = SPZ_NEVER | (gen << 21) | 0x1FFFu;
        unsigned long long T = ~0ull;
        if (elig) {
          const int32_t top = lds32(reca + (uint32_t)(SLP + 2) * 4u);
          const double stick = lds64f(reca + (uint32_t)SLP * 4u + 16u);
          const int4 qs = lds128(qs_a + slot * 16u);
          const int32_t qsv[4] = {qs.x, qs.y, qs.z, qs.w};
          unsigned long long key[K];
          bool ok = true;
#pragma unroll
          for (int q = 0; q < K; ++q) {
            const int32_t a = have_p ? ld_relaxed_gpu(A + (size_t)top * N + cn[q]) : 0;
            const int4 ma = lds128(nd_a + (uint32_t)cn[q] * 32u), mb2 = lds128(nd_a + (uint32_t)cn[q] * 32u + 16u);
            const uint32_t fl = dyn_smem[(flg_a - base_a) + cn[q]];
            ok = ok && (fl & NF_VALID);
            key[q] = sp_key(__hiloint2double(ma.y, ma.x), __hiloint2double(ma.w, ma.z), __hiloint2double(mb2.y, mb2.x),
                            __hiloint2double(mb2.w, mb2.z), (fl & NF_BOOST) != 0, has_nw, qsv[q] + a, stick, have_p, Pd, Py);
          }
          if (ok) {
            int worst = 0;
            uint32_t ranks = 0;
#pragma unroll
            for (int q = 0; q < K; ++q) {
              int rank = 0;
#pragma unroll
              for (int t = 0; t < K; ++t)
                if (t != q && lex_lt(key[t], (uint32_t)cn[t], key[q], (uint32_t)cn[q])) ++rank;
              ranks |= (uint32_t)rank << (2 * q);
              if (rank == K - 1) worst = q;
            }
            T = key[0];
            uint32_t tp = (uint32_t)cn[0];
#pragma unroll
            for (int q = 1; q < K; ++q) if (worst == q) { T = key[q]; tp = (uint32_t)cn[q]; }
            z = (gen << 21) | (ranks << 13) | tp;
          }
        }
        __threadfence_block();                          // the record (TMA) and my reads are ordered before the result
        if (live) sts128(dyn_a + slot * 16u, (uint32_t)(T >> 32), (uint32_t)T, z, (uint32_t)e);
      }
      if (!any_work) __nanosleep(500);
      else if (!busy) __nanosleep(100);
    }
  scouts_done:
    // ---- write the per-node counts of this state back ------------------------------------------------------
    for (int n = sidx * 32 + lane; n < N; n += TS) counts[s * N + n] = __double2int_rn(reinterpret_cast<double*>(dyn_smem)[4 * (size_t)n]);
    return;
  }

  // ======================================== leader =============================================================
  unsigned long long Lk[SP_LPL];
  int32_t Ln[SP_LPL];
#pragma unroll
  for (int u = 0; u < SP_LPL; ++u) { Lk[u] = ~0ull; Ln[u] = -1; }
  unsigned long long ubk = ~0ull, B0k = ~0ull, lb1k = ~0ull;      // lb1: lower bound of the second-column entries
  uint32_t ubp = 0xFFFFFFFFu, B0p = 0xFFFFFFFFu, lb1p = 0xFFFFFFFFu;
  long long n_round2 = 0;
  int32_t E = 0;
  int seq = 0, pub_head = 0, movers_since_rebuild = 0;
  bool b0_clamped = false;
  long long n_fast = 0, n_res = 0, n_mov = 0, n_team = 0, n_reb = 0, n_wait = 0, n_stale = 0, n_cwait = 0;
  long long why[4] = {0, 0, 0, 0};
  long long cyc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc = clock64();
  const long long t_begin = tc;
#ifdef BLANCE_SPEC_TIMING
#define SP_T(ix) do { const long long n_ = clock64(); cyc[ix] += n_ - tc; tc = n_; } while (0)
#else
#define SP_T(ix) do { } while (0)
#endif

  // Watchdog of the leader's wait loops: a correct run never gets near the limit; if one does, say where and stop
  // the pass (the plan then fails with BLANCE_ERR_CUDA instead of hanging the GPU).
  long long spins = 0;
  auto stuck = [&](int where, int a, int b) {
    if (++spins < (1ll << 26)) return false;
    if (lane == 0) {
      printf("[blance] speculative pass stuck (wait %d) at step %d of %d: %d %d | front %d commit_done %d epoch %d pub %d/%d\n", where, a, n_assign, b, 0,
             *(volatile int32_t*)&ctl.front, *(volatile int32_t*)&ctl.commit_done, *(volatile int32_t*)&ctl.epoch,
             *(volatile int32_t*)&ctl.pub_done, *(volatile int32_t*)&ctl.pub_head);
      sts32v(abort_a, 1);
      D.spec_abort = 1;
    }
    return true;
  };
  bool aborted = false;
  auto team_cmd = [&](int op, int arg) {
    if (lane == 0) {
      *(volatile int32_t*)&ctl.cmd_op = op;
      *(volatile int32_t*)&ctl.cmd_arg = arg;
      *(volatile int32_t*)&ctl.cmd_epoch = E + 1;
      __threadfence_block();
      sts32v(seq_a, ++seq);
    } else ++seq;
    __syncwarp();
    bar_sync(SPB_GO, NTT);
    if (op != SP_OP_EXIT) bar_sync(SPB_DONE, NTT);
  };
  // The list.  Lane l owns the nodes n = l (mod 32): its two entries are (after a rebuild) the two smallest base
  // keys of its class and (ublk, ublp) bounds every unlisted node of the class from below; ub = the smallest of
  // the 32 lane bounds.  Updates stay inside the owner lane; a rebuild is a scan of the base keys in shared
  // memory by the leader alone.  B0 = min(smallest listed key, ub) is a lower bound of every live base key.
  unsigned long long ublk = ~0ull;
  uint32_t ublp = 0xFFFFFFFFu;
  auto recompute_ub = [&]() {
    const Best b = warp_argmin_q(Best{(uint32_t)(ublk >> 32), (uint32_t)ublk, ublp});
    ubk = ((unsigned long long)b.hi << 32) | b.lo;
    ubp = b.pos;
  };
  auto recompute_b0 = [&]() {
    unsigned long long bk = ~0ull;
    uint32_t bp = 0xFFFFFFFFu;
#pragma unroll
    for (int u = 0; u < SP_LPL; ++u)
      if (Ln[u] >= 0 && lex_lt(Lk[u], (uint32_t)Ln[u], bk, bp)) { bk = Lk[u]; bp = (uint32_t)Ln[u]; }
    const Best b = warp_argmin_q(Best{(uint32_t)(bk >> 32), (uint32_t)bk, bp});
    B0k = ((unsigned long long)b.hi << 32) | b.lo;
    B0p = b.pos;
    b0_clamped = !lex_lt(B0k, B0p, ubk, ubp);
    if (b0_clamped) { B0k = ubk; B0p = ubp; }
  };
  auto rebuild_list = [&]() {               // the two smallest base keys of my class, and the third as my bound
    unsigned long long k0 = ~0ull, k1 = ~0ull, k2 = ~0ull;
    uint32_t n0 = 0xFFFFFFFFu, n1 = 0xFFFFFFFFu, n2 = 0xFFFFFFFFu;
    for (int n = lane; n < N; n += 32) {
      if (!(dyn_smem[(flg_a - base_a) + n] & NF_VALID)) continue;
      const unsigned long long k = *reinterpret_cast<const volatile unsigned long long*>(dyn_smem + (bk_a - base_a) + 8u * n);
      if (lex_lt(k, (uint32_t)n, k2, n2)) {
        if (lex_lt(k, (uint32_t)n, k1, n1)) {
          k2 = k1; n2 = n1;
          if (lex_lt(k, (uint32_t)n, k0, n0)) { k1 = k0; n1 = n0; k0 = k; n0 = (uint32_t)n; }
          else { k1 = k; n1 = (uint32_t)n; }
        } else { k2 = k; n2 = (uint32_t)n; }
      }
    }
    Lk[0] = k0; Ln[0] = (int32_t)n0; Lk[1] = k1; Ln[1] = (int32_t)n1;      // (n = 0xFFFFFFFF reads as -1: no entry)
    ublk = k2; ublp = n2;
    recompute_ub();
    {
      const Best b = warp_argmin_q(Best{(uint32_t)(k1 >> 32), (uint32_t)k1, n1});
      lb1k = ((unsigned long long)b.hi << 32) | b.lo;
      lb1p = b.pos;
    }
    recompute_b0();
    movers_since_rebuild = 0;
    ++n_reb;
  };
  // a node whose base key changed: its owner lane updates / lists / bounds it.  tk is its new key (tvalid: live)
  auto list_touch = [&](int32_t tx, unsigned long long tk, bool tvalid) {
    bool ub_moved = false;
    if ((tx & 31) == lane) {
      if (Ln[0] == tx) { if (tvalid) Lk[0] = tk; else { Ln[0] = -1; Lk[0] = ~0ull; } }
      else if (Ln[1] == tx) { if (tvalid) Lk[1] = tk; else { Ln[1] = -1; Lk[1] = ~0ull; } }
      else if (tvalid && lex_lt(tk, (uint32_t)tx, ublk, ublp)) {          // below my bound: it has to be listed
        if (Ln[0] < 0) { Ln[0] = tx; Lk[0] = tk; }
        else if (Ln[1] < 0) { Ln[1] = tx; Lk[1] = tk; }
        else {
          const int big = lex_lt(Lk[0], (uint32_t)Ln[0], Lk[1], (uint32_t)Ln[1]) ? 1 : 0;
          unsigned long long ek = tk;
          uint32_t en = (uint32_t)tx;                                     // the one that stays out: the largest of the three
          if (lex_lt(tk, (uint32_t)tx, Lk[big], (uint32_t)Ln[big])) { ek = Lk[big]; en = (uint32_t)Ln[big]; Lk[big] = tk; Ln[big] = tx; }
          if (lex_lt(ek, en, ublk, ublp)) { ublk = ek; ublp = en; ub_moved = true; }
        }
      }
      if (Ln[0] >= 0 && Ln[1] >= 0 && lex_lt(Lk[1], (uint32_t)Ln[1], Lk[0], (uint32_t)Ln[0])) {   // smaller entry first
        const unsigned long long k = Lk[0]; Lk[0] = Lk[1]; Lk[1] = k;
        const int32_t n = Ln[0]; Ln[0] = Ln[1]; Ln[1] = n;
      } else if (Ln[0] < 0 && Ln[1] >= 0) { Lk[0] = Lk[1]; Ln[0] = Ln[1]; Lk[1] = ~0ull; Ln[1] = -1; }
    }
    // lb1 bounds the second-column entries from below: it moves when the node sits there now with a smaller key
    if (__any_sync(0xFFFFFFFFu, (tx & 31) == lane && Ln[1] >= 0 && lex_lt(Lk[1], (uint32_t)Ln[1], lb1k, lb1p))) {
      const int o = tx & 31;
      lb1k = __shfl_sync(0xFFFFFFFFu, Lk[1], o);
      lb1p = (uint32_t)__shfl_sync(0xFFFFFFFFu, Ln[1], o);
    }
    if (__any_sync(0xFFFFFFFFu, ub_moved)) recompute_ub();
  };
  // Hands epoch E + 1 to the publisher: the A updates of the lanes with `has`, then the marker that lets it
  // publish.  The mirror / lastchg stores above are ordinary shared-memory stores of this warp, issued before
  // the store of pub_head, so whoever sees the epoch also sees them.
  auto publish = [&](bool has, int32_t idx, int32_t dA) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, has);
    const int n = __popc(m) + 1;
    while (pub_head + n - lds32v(pubd_a) > 64) { __nanosleep(20); if (stuck(1, pub_head, n)) { aborted = true; break; } }
    if (has) sts128(pubq_a + 16u * (uint32_t)((pub_head + __popc(m & ((1u << lane) - 1u))) & 63), (uint32_t)idx, (uint32_t)dA, (uint32_t)(E + 1), 0u);
    if (lane == 0) sts128(pubq_a + 16u * (uint32_t)((pub_head + n - 1) & 63), 0xFFFFFFFFu, 0u, (uint32_t)(E + 1), 1u);
    pub_head += n;
    __syncwarp();
    if (lane == 0) sts32v(pubh_a, pub_head);
    ++E;
  };

  rebuild_list();
  SP_T(6);

  int i = 0;
  // The leader looks at the steps in aligned windows of 64 (two per lane).  A window is loaded once - results,
  // current nodes, and the stamp test of every result - and stays in registers: after a mover the rest of the
  // window is judged again from the registers (a result whose current node the mover touched is dropped, the
  // others only meet the new B0), so a mover costs a ballot, not a reload.
  int w = -64;                                   // base of the window in registers
  int4 r[2];
  int32_t cn[2][K];
  uint32_t slotv[2];
  bool have[2], never[2], fresh[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) { r[u] = make_int4(0, 0, 0, 0); slotv[u] = 0; have[u] = never[u] = fresh[u] = false; for (int q = 0; q < K; ++q) cn[u][q] = 0; }
  while (i < n_assign) {
    if (i >= w + 64) {
      w = i & ~63;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = w + 32 * u + lane;
        const uint32_t cj = (uint32_t)j >> 5;
        const uint32_t slot = ((cj & swd_mask) << 5) | ((uint32_t)j & 31u);
        slotv[u] = slot;
        r[u] = lds128(dyn_a + slot * 16u);
#pragma unroll
        for (int q = 0; q < K; ++q) cn[u][q] = lds32(rec_a + slot * RECB + (uint32_t)(lo_s + q) * 4u);
        have[u] = j < n_assign && (((uint32_t)r[u].z >> 21) & 0x3FFu) == ((cj >> swd_shift) & 0x3FFu);
        never[u] = ((uint32_t)r[u].z & SPZ_NEVER) != 0;
        fresh[u] = have[u] && !never[u];
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (fresh[u]) {
#pragma unroll
          for (int q = 0; q < K; ++q) fresh[u] = fresh[u] && lds32(chg_a + 4u * (uint32_t)cn[u][q]) <= r[u].w;
        }
    }
    // ---- accept the leading run of results that are exact and sticky ------------------------------------------------
    const int done = i - w;                        // window positions below `done` are behind the leader
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const unsigned long long T = ((unsigned long long)(uint32_t)r[u].x << 32) | (uint32_t)r[u].y;
      ok[u] = (32 * u + lane < done) || (fresh[u] && lex_lt(T, (uint32_t)r[u].z & 0x1FFFu, B0k, B0p));
    }
    const uint32_t okm0 = __ballot_sync(0xFFFFFFFFu, ok[0]), okm1 = __ballot_sync(0xFFFFFFFFu, ok[1]);
    const int f = okm0 != 0xFFFFFFFFu ? (__ffs(~okm0) - 1) : (okm1 != 0xFFFFFFFFu ? 32 + (__ffs(~okm1) - 1) : 64);
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (32 * u + lane >= done && 32 * u + lane < f)                     // accepted: the committer does the rest
        dyn_smem[(acc_a - base_a) + slotv[u]] = (unsigned char)(0x80u | (((uint32_t)r[u].z >> 13) & 0xFFu));
    n_fast += f - done;
    i = w + f;
    __syncwarp();
    if (f > done && lane == 0) sts32v(front_a, i);
    if (aborted) break;
    if (f == 64 || i >= n_assign) continue;
    SP_T(0);
    // ---- step i was not accepted -----------------------------------------------------------------------------------
    const uint32_t slot_i = ((((uint32_t)i >> 5) & swd_mask) << 5) | ((uint32_t)i & 31u);
    {
      const bool fr = f < 32 ? fresh[0] : fresh[1];
      if (!__shfl_sync(0xFFFFFFFFu, (int)fr, f & 31)) {
        // no usable result in the registers: its scout may have written one since the window was loaded
        const int4 rr = lds128(dyn_a + slot_i * 16u);
        const bool hv = (((uint32_t)rr.z >> 21) & 0x3FFu) == ((((uint32_t)i >> 5) >> swd_shift) & 0x3FFu);
        if (!hv) {                                                        // its scout has not got there yet
          ++n_wait; __nanosleep(40); SP_T(1);
          if (stuck(2, i, (int)rr.z)) { aborted = true; break; }
          continue;
        }
        spins = 0;
        bool frs = !((uint32_t)rr.z & SPZ_NEVER);
        if (frs) {
#pragma unroll
          for (int q = 0; q < K; ++q) frs = frs && lds32(chg_a + 4u * (uint32_t)lds32(rec_a + slot_i * RECB + (uint32_t)(lo_s + q) * 4u)) <= rr.w;
        }
        const unsigned long long T = ((unsigned long long)(uint32_t)rr.x << 32) | (uint32_t)rr.y;
        if (frs && lex_lt(T, (uint32_t)rr.z & 0x1FFFu, B0k, B0p)) {     // it is sticky after all
          if (lane == 0) {
            dyn_smem[(acc_a - base_a) + slot_i] = (unsigned char)(0x80u | (((uint32_t)rr.z >> 13) & 0xFFu));
            sts32v(front_a, i + 1);
          }
          ++i;
          ++n_fast;
          w = -128;                                                       // its neighbours may be new as well: reload the window
          continue;
        }
        if (!((uint32_t)rr.z & SPZ_NEVER) && !frs) ++n_stale;
      }
    }
    ++n_res;
    {
      const uint32_t rb = rec_a + slot_i * RECB;
      if (lane == 0) dyn_smem[(acc_a - base_a) + slot_i] = 1;            // the leader writes this step's outcome itself
      const int4 hdr = lds128(rb + (uint32_t)SLP * 4u);               // meta, w_p, top, partition
      const int32_t w_p = hdr.y, top = hdr.z;
      const double stick = lds64f(rb + (uint32_t)SLP * 4u + 16u);
      const int n_cur = lds32(rb + (uint32_t)(SLP + 6) * 4u);
      const bool row_clean = lds32(rb + (uint32_t)(SLP + 7) * 4u) != 0;
      // the row, in every lane (two broadcast loads); slots beyond SLP read as empty
      int32_t rowv[8];
      {
        const int4 r0 = lds128(rb);
        const int4 r1 = SLP > 4 ? lds128(rb + 16u) : make_int4(BLANCE_NO_NODE, BLANCE_NO_NODE, BLANCE_NO_NODE, BLANCE_NO_NODE);
        rowv[0] = r0.x; rowv[1] = r0.y; rowv[2] = r0.z; rowv[3] = r0.w;
        rowv[4] = r1.x; rowv[5] = r1.y; rowv[6] = r1.z; rowv[7] = r1.w;
      }
      const int32_t mycur = lane < n_cur ? lds32(rb + (uint32_t)(lo_s + (lane & 7)) * 4u) : -1;   // lane q: q-th current node
      const int32_t* Gt = G + (size_t)top * N;
      // nodeToNodeCounts[top] must hold every commit before step i.  The committer trails by a few hundred cycles;
      // only an accepted step with the same top among the ones it has not reached yet makes the leader wait.
      {
        int cd = lds32v(cdone_a);
        if (cd < i) {
          bool clash = i - cd > 128;
          for (int j0 = cd; j0 < i && !clash; j0 += 32) {
            const int j = j0 + lane;
            bool mine = false;
            if (j < i) {
              const uint32_t sj = ((((uint32_t)j >> 5) & swd_mask) << 5) | ((uint32_t)j & 31u);
              mine = (dyn_smem[(acc_a - base_a) + sj] & 0x80u) && lds32(rec_a + sj * RECB + (uint32_t)(SLP + 2) * 4u) == top;
            }
            clash = __any_sync(0xFFFFFFFFu, mine);
          }
          if (clash) { ++n_cwait; while (lds32v(cdone_a) < i) { if (stuck(3, i, cd)) { aborted = true; break; } } spins = 0; }
        }
      }
      bool retried = false;
    resolve_again:
      bool resolved = false;
      int n_ch = 0;
      int32_t chosen[K];
#pragma unroll
      for (int t = 0; t < K; ++t) chosen[t] = BLANCE_NO_NODE;
      bool same = false;
      int reason = 0;
      if (row_clean && n_cur <= K) {
        // Candidates: on lanes < n_cur a current node (with the stickiness), and my listed nodes unless a
        // higher-priority state of the row holds them.  (A listed node that is also current needs no test: its
        // listed key lacks the stickiness, so the current twin is picked first and takes the listed one with it -
        // unless the stickiness is negative, then the state's own slots block too.)  The second list entry of
        // every lane is only looked at when the K-th winner of the first round is not below lb1, a lower bound
        // of all second entries.
        const uint32_t blk = stick < 0.0 ? (slot_blocked | slot_state_s) : slot_blocked;
        unsigned long long ck[SP_LPL + 1];
        int32_t cnode[SP_LPL + 1];
        bool cur_ok = true;
        auto eval = [&](int u) {                                        // exact key of candidate u (plan.go:634-689)
          ck[u] = ~0ull;
          if (u < SP_LPL) {
            cnode[u] = Ln[u];
            if (blk)
#pragma unroll
              for (int sl = 0; sl < 8; ++sl)
                if (((blk >> sl) & 1u) && rowv[sl] == Ln[u]) cnode[u] = -1;
          } else cnode[u] = mycur;
          if (cnode[u] >= 0) {
            const int32_t g = have_p ? ld_relaxed_gpu(Gt + cnode[u]) : 0;
            const int4 ma = lds128(nd_a + (uint32_t)cnode[u] * 32u), mb = lds128(nd_a + (uint32_t)cnode[u] * 32u + 16u);
            const uint32_t fl = dyn_smem[(flg_a - base_a) + cnode[u]];
            if (u == SP_LPL && !(fl & NF_VALID)) cur_ok = false;
            ck[u] = sp_key(__hiloint2double(ma.y, ma.x), __hiloint2double(ma.w, ma.z), __hiloint2double(mb.y, mb.x),
                           __hiloint2double(mb.w, mb.z), (fl & NF_BOOST) != 0, has_nw, g, u == SP_LPL ? stick : 0.0, have_p, Pd, Py);
          }
        };
        eval(SP_LPL);
        eval(0);
#pragma unroll
        for (int u = 1; u < SP_LPL; ++u) { cnode[u] = -1; ck[u] = ~0ull; }
        SP_T(2);
        if (__all_sync(0xFFFFFFFFu, cur_ok)) {
          const int32_t cnodeC = cnode[SP_LPL];
          for (int round = 0; round < 2; ++round) {
            unsigned long long lastk = 0;
            uint32_t lastp = 0;
            bool hit_all = true;
            n_ch = 0;
            uint32_t alive = 0;                                         // bit u: my candidate u is still in play
#pragma unroll
            for (int u = 0; u <= SP_LPL; ++u) if (cnode[u] >= 0) alive |= 1u << u;
            for (int t = 0; t < K; ++t) {
              unsigned long long bk = ~0ull;
              uint32_t bp = 0xFFFFFFFFu;
#pragma unroll
              for (int u = 0; u <= SP_LPL; ++u)
                if (((alive >> u) & 1u) && lex_lt(ck[u], (uint32_t)cnode[u], bk, bp)) { bk = ck[u]; bp = (uint32_t)cnode[u]; }
              const Best b = warp_argmin_q(Best{(uint32_t)(bk >> 32), (uint32_t)bk, bp});
              if (b.pos == 0xFFFFFFFFu) break;
              chosen[t] = (int32_t)b.pos;
              ++n_ch;
              lastk = ((unsigned long long)b.hi << 32) | b.lo;
              lastp = b.pos;
              hit_all = hit_all && __any_sync(0xFFFFFFFFu, cnodeC == (int32_t)b.pos);
#pragma unroll
              for (int u = 0; u <= SP_LPL; ++u) if (cnode[u] == (int32_t)b.pos) alive &= ~(1u << u);
            }
            // could an entry of the second column be among the first K?
            if (round == 0 && SP_LPL > 1 && !(n_ch == K && lex_lt(lastk, lastp, lb1k, lb1p))) {
#pragma unroll
              for (int u = 1; u < SP_LPL; ++u) eval(u);
              ++n_round2;
              continue;
            }
            const bool complete = ubp == 0xFFFFFFFFu;                    // every live node is listed
            if (n_ch == K) { resolved = complete || lex_lt(lastk, lastp, ubk, ubp); reason = 3; }
            else { resolved = complete; reason = 2; }
            same = resolved && hit_all && n_ch == n_cur && n_cur == K;
            break;
          }
        } else reason = 1;
        SP_T(3);
      }
      if (!resolved && reason == 3 && movers_since_rebuild > 0 && !retried) {
        // the K-th winner is not provably below every unlisted node: a fresh list usually settles it
        retried = true;
        rebuild_list();
        SP_T(6);
        goto resolve_again;
      }
      if (!resolved) {
        // ---- the team evaluates the step (and rebuilds the list if a count changed) ----------------------------------
        ++n_team;
        ++why[reason];
        team_cmd(SP_OP_FULL, i);
        n_ch = *(volatile int32_t*)&ctl.res_n;
        same = *(volatile int32_t*)&ctl.res_same != 0;
#pragma unroll
        for (int t = 0; t < K; ++t) chosen[t] = t < n_ch ? *(volatile int32_t*)&ctl.res_chosen[t] : BLANCE_NO_NODE;
        int32_t* orec = ostream + (size_t)i * REC;
        if (lane == 0) {
#pragma unroll
          for (int t = 0; t < K; ++t) orec[t] = chosen[t];
          orec[K] = n_ch;
        }
        if (!same) { rebuild_list(); publish(false, 0, 0); ++n_mov; w = -64 - 64; }   // (stamps changed: reload the window)
        SP_T(6);
      } else {
        if (lane <= K) {                                               // outcome record: chosen[0..K), n_chosen
          int32_t v = n_ch;
#pragma unroll
          for (int t = 0; t < K; ++t) if (lane == t) v = chosen[t];
          ostream[(size_t)i * REC + lane] = v;
          if (lane < n_ch && have_p) atomicAdd(&G[(size_t)top * N + v], 1);                  // plan.go:238-245
        }
        if (!same) {
          // ---- a mover: lanes 0..n_cur-1 take the old nodes, lanes 8..8+n_ch-1 the new ones -------------------------
          ++n_mov;
          ++movers_since_rebuild;
          const int32_t E1 = E + 1;
          const bool elig = n_cur == K;
          // A[top][x] changes by (x is chosen) - (the hypothesis counted x: eligible row and x current); a node whose
          // count or A entry changes is stamped, so results computed from the old values are rejected
          int delta = 0, dA = 0;
          int32_t newx = chosen[0];
#pragma unroll
          for (int t = 1; t < K; ++t) if (lane - 8 == t) newx = chosen[t];
          const bool is_old = lane < n_cur, is_new = lane >= 8 && lane < 8 + n_ch;
          const int32_t x = is_old ? mycur : (is_new ? newx : -1);
          bool again = false;                   // old node that is chosen again / new node that was current
          uint32_t memb = 0;                    // states (other than s) whose list holds my new node
          if (is_old) {
#pragma unroll
            for (int t = 0; t < K; ++t) again = again || (t < n_ch && chosen[t] == x);
          }
          if (is_new) {
#pragma unroll
            for (int sl = 0; sl < 8; ++sl)
              if (rowv[sl] == x) {
                if ((slot_state_s >> sl) & 1u) again = true;
                else memb |= sbit8[sl];
              }
          }
          bool act = false;
          if (is_old && !again) { act = true; delta = -w_p; dA = elig ? -1 : 0; }
          if (is_new && !again) { act = true; delta = w_p; dA = 1; }
          if (is_new && again && !elig) { act = true; dA = 1; memb = 0; }      // kept node of a short row: only A moves
          act = act && x >= 0;
          unsigned long long nk = ~0ull;
          bool tvalid = false;
          if (act) {
            const int4 ma = lds128(nd_a + (uint32_t)x * 32u), mb = lds128(nd_a + (uint32_t)x * 32u + 16u);
            const uint32_t fl = dyn_smem[(flg_a - base_a) + x];
            double cd = __hiloint2double(ma.y, ma.x), ff = __hiloint2double(ma.w, ma.z);
            cd = __dadd_rn(cd, (double)delta);
            int32_t t0 = lds32(tot_a + 4u * (uint32_t)x), t = t0 + delta;
            while (memb) {
              const int s2 = __ffs(memb) - 1;
              memb &= memb - 1;
              atomicSub(&counts[s2 * N + x], w_p);
              t -= w_p;
            }
            if (t != t0) {
              sts32(tot_a + 4u * (uint32_t)x, t);
              if (have_p) ff = div_exact(__dmul_rn(0.001, (double)t), Pd, Py);
            }
            double* nd = reinterpret_cast<double*>(dyn_smem) + 4 * (size_t)x;
            nd[0] = cd; nd[1] = ff;
            sts32(chg_a + 4u * (uint32_t)x, E1);
            nk = sp_base_key(cd, ff, __hiloint2double(mb.y, mb.x), __hiloint2double(mb.w, mb.z), (fl & NF_BOOST) != 0, has_nw);
            *reinterpret_cast<unsigned long long*>(dyn_smem + (bk_a - base_a) + 8u * (uint32_t)x) = nk;
            tvalid = (fl & NF_VALID) != 0;
          }
          SP_T(4);
          // list: every touched node goes to its owner lane (new key in place, listed if it fell below the lane's
          // bound); window results computed from its old state are dropped
          const uint32_t actm = __ballot_sync(0xFFFFFFFFu, act);
          for (uint32_t m = actm; m; m &= m - 1u) {
            const int src = __ffs(m) - 1;
            const int32_t tx = __shfl_sync(0xFFFFFFFFu, x, src);
            const unsigned long long tk = __shfl_sync(0xFFFFFFFFu, nk, src);
            const bool tv = __shfl_sync(0xFFFFFFFFu, (int)tvalid, src) != 0;
            list_touch(tx, tk, tv);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
              for (int q = 0; q < K; ++q) if (cn[u][q] == tx) fresh[u] = false;
          }
          publish(act && dA != 0 && have_p, (int32_t)((size_t)top * N + (x < 0 ? 0 : x)), dA);
          recompute_b0();
          // a list whose smallest key is no longer below ub has lost its grip on the minimum: scan the base keys again
          if (b0_clamped && ubp != 0xFFFFFFFFu && movers_since_rebuild >= 4) { SP_T(5); rebuild_list(); SP_T(6); }
          else SP_T(5);
        }
      }
      ++i;
      if (lane == 0) sts32v(front_a, i);
    }
  }
  SP_T(0);
  team_cmd(SP_OP_EXIT, 0);
  if (lane == 0) sts32v(pubx_a, 1);
  if (lane == 0) {
    D.steps += n_assign;
    D.fast_steps += n_fast;
    D.spec_resolved += n_res; D.spec_movers += n_mov; D.spec_team += n_team; D.spec_rebuilds += n_reb;
    D.spec_waits += n_wait; D.spec_stale += n_stale;
    cyc[7] = clock64() - t_begin;
    for (int x = 0; x < 8; ++x) D.spec_cyc[x] += cyc[x];
    for (int x = 0; x < 4; ++x) D.spec_why[x] += why[x];
    D.spec_cwait += n_cwait;
    D.spec_round2 += n_round2;
  }
#undef SP_T
}

}  // namespace blance_dev
