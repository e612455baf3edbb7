// blance_b200/csrc/host_api.cpp — see host_api.hpp.  Interning (strings -> flat
// int32 tables), the calls into the CUDA library, and the way back to maps.
#include "host_api.hpp"

#include <algorithm>
#include <cstdio>
#include <mutex>
#include <unordered_set>

namespace blance {

namespace {

const Strs kNoStrs;
inline const Strs& deref(const OptStrs& s) { return s ? *s : kNoStrs; }

[[noreturn]] void invalid(const std::string& msg) { throw BlanceError(BLANCE_ERR_INVALID_ARG, "blance: " + msg); }

// fmt.Sprintf("%10d", v), plan.go:527,539
std::string pad10(long long v) {
  char buf[32];
  std::snprintf(buf, sizeof buf, "%10lld", v);
  return buf;
}

// strconv.Atoi (plan.go:525): optional sign, decimal digits, must fit int64.
bool go_atoi(const std::string& s, long long* out) {
  size_t i = 0;
  if (s.empty()) return false;
  bool neg = false;
  if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
  if (i >= s.size()) return false;
  unsigned long long acc = 0;
  const unsigned long long lim = neg ? (1ULL << 63) : (1ULL << 63) - 1;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    unsigned d = unsigned(s[i] - '0');
    if (acc > (lim - d) / 10) return false;
    acc = acc * 10 + d;
  }
  *out = neg ? -(long long)acc : (long long)acc;
  return true;
}

// The name part of the partition sort key: the padded form first, the raw name as
// the final tie-break (plan.go:519-528, 512).
struct NameKey { std::string padded, raw; };
bool name_key_less(const NameKey& a, const NameKey& b) {
  if (a.padded != b.padded) return a.padded < b.padded;
  return a.raw < b.raw;
}

struct Interner {
  std::unordered_map<std::string, int32_t> ids;
  Strs names;
  int32_t get(const std::string& s) {
    auto it = ids.find(s);
    if (it != ids.end()) return it->second;
    int32_t id = int32_t(names.size());
    ids.emplace(s, id);
    names.push_back(s);
    return id;
  }
  int32_t find(const std::string& s) const {
    auto it = ids.find(s);
    return it == ids.end() ? -1 : it->second;
  }
};

// sortStateNames (plan.go:437-474).  Go starts from random map order and its
// comparator is inconsistent when name order disagrees with priority order (the
// reference is then non-deterministic); starting from ascending names and running
// Go's small-slice insertion sort gives (priority, name) order whenever the
// reference is deterministic.
Strs sort_state_names(const PartitionModel& model) {
  Strs s;
  for (const auto& kv : model) s.push_back(kv.first);
  std::sort(s.begin(), s.end());
  auto less = [&](const std::string& i, const std::string& j) {
    return model.at(i).Priority < model.at(j).Priority || i < j;
  };
  for (size_t i = 1; i < s.size(); ++i)
    for (size_t j = i; j > 0 && less(s[j], s[j - 1]); --j) std::swap(s[j], s[j - 1]);
  return s;
}

// --- hierarchy as strings (plan.go:703-774), evaluated once per (rule, anchor) ---
struct Hierarchy {
  const std::unordered_map<std::string, std::string>* parents;
  std::unordered_map<std::string, Strs> children;
  std::unordered_map<std::string, Strs> leaves_cache;

  explicit Hierarchy(const std::unordered_map<std::string, std::string>* p) : parents(p) {
    Strs nodes;                                          // plan.go:705-716
    if (p) for (const auto& kv : *p) nodes.push_back(kv.first);
    std::sort(nodes.begin(), nodes.end());
    for (const auto& c : nodes) children[p->at(c)].push_back(c);
  }
  std::string ancestor(std::string node, int level) const {   // plan.go:755-762
    while (level > 0) {
      if (!parents) { node.clear(); }
      else { auto it = parents->find(node); node = it == parents->end() ? std::string() : it->second; }
      --level;
    }
    return node;
  }
  const Strs& leaves(const std::string& node, int depth = 0) {   // plan.go:764-774
    auto hit = leaves_cache.find(node);
    if (hit != leaves_cache.end()) return hit->second;
    if (depth > 4096) invalid("NodeHierarchy contains a cycle (the reference recurses forever)");
    Strs rv;
    auto it = children.find(node);
    if (it == children.end() || it->second.empty()) rv.push_back(node);
    else for (const auto& c : it->second) { const Strs& sub = leaves(c, depth + 1); rv.insert(rv.end(), sub.begin(), sub.end()); }
    return leaves_cache.emplace(node, std::move(rv)).first->second;
  }
};

int32_t checked_i32(long long v, const char* what) {
  if (v < INT32_MIN || v > INT32_MAX) invalid(std::string(what) + " does not fit int32");
  return int32_t(v);
}

}  // namespace

// ------------------------------------------------------------------------------------

std::unique_ptr<InternedPlan> InternPlan(const PartitionMap& prevMap, const PartitionMap& partitionsToAssign,
                                         const Strs& nodesAll, const OptStrs& nodesToRemove,
                                         const OptStrs& nodesToAdd, const PartitionModel& model,
                                         const PlanNextMapOptions& options) {
  auto ip = std::make_unique<InternedPlan>();
  blance_plan_in& in = ip->in;

  // ---- nodes
  Interner nodes;
  for (const auto& n : nodesAll) {
    if (nodes.find(n) >= 0) invalid("nodesAll contains '" + n + "' twice");
    nodes.get(n);
  }
  const int32_t N = int32_t(nodesAll.size());

  // ---- states
  ip->state_names = sort_state_names(model);
  const int32_t S = int32_t(ip->state_names.size());
  std::unordered_map<std::string, int32_t> state_id;
  for (int32_t s = 0; s < S; ++s) state_id[ip->state_names[size_t(s)]] = s;
  ip->state_priority.resize(size_t(S));
  ip->state_constraints.resize(size_t(S));
  ip->state_stickiness.assign(size_t(S), 0);
  ip->state_has_stickiness.assign(size_t(S), 0);
  int32_t top_state = -1;
  for (int32_t s = 0; s < S; ++s) {
    const auto& name = ip->state_names[size_t(s)];
    const auto& ms = model.at(name);
    ip->state_priority[size_t(s)] = ms.Priority;
    int k = ms.Constraints;                                        // plan.go:308-319
    if (options.ModelStateConstraints) {
      auto it = options.ModelStateConstraints->find(name);
      if (it != options.ModelStateConstraints->end()) k = it->second;
    }
    ip->state_constraints[size_t(s)] = k;
    if (options.StateStickiness) {
      auto it = options.StateStickiness->find(name);
      if (it != options.StateStickiness->end()) { ip->state_stickiness[size_t(s)] = it->second; ip->state_has_stickiness[size_t(s)] = 1; }
    }
  }
  {
    // plan.go:126-132 walks the model in map order; ties on the minimum priority are
    // resolved here as "first in ascending name order" (the reference is random).
    Strs by_name = ip->state_names;
    std::sort(by_name.begin(), by_name.end());
    for (const auto& n : by_name)
      if (top_state < 0 || model.at(n).Priority < ip->state_priority[size_t(top_state)]) top_state = state_id[n];
  }

  // ---- partitions, indexed in the name order of the partition sort key
  std::vector<NameKey> keys;
  {
    std::unordered_set<std::string> seen;
    auto add = [&](const PartitionMap& m) {
      for (const auto& kv : m) {
        if (!kv.second.Name.empty() && kv.second.Name != kv.first)
          invalid("Partition.Name '" + kv.second.Name + "' differs from its map key '" + kv.first + "'");
        if (!seen.insert(kv.first).second) continue;
        NameKey k;
        k.raw = kv.first;
        k.padded = kv.first;
        long long v;
        if (go_atoi(kv.first, &v) && v >= 0) k.padded = pad10(v);
        keys.push_back(std::move(k));
      }
    };
    add(prevMap);
    add(partitionsToAssign);
  }
  std::sort(keys.begin(), keys.end(), name_key_less);
  const int32_t PU = int32_t(keys.size());
  ip->part_names.resize(size_t(PU));
  std::unordered_map<std::string, int32_t> part_id;
  part_id.reserve(size_t(PU) * 2);
  for (int32_t p = 0; p < PU; ++p) { ip->part_names[size_t(p)] = std::move(keys[size_t(p)].raw); part_id[ip->part_names[size_t(p)]] = p; }

  // ---- slot layout: a state's range holds max(constraints, longest input list)
  std::vector<int32_t> cap(size_t(S), 0);
  for (int32_t s = 0; s < S; ++s) cap[size_t(s)] = std::max(0, ip->state_constraints[size_t(s)]);
  auto scan_caps = [&](const PartitionMap& m, bool must_be_model) {
    for (const auto& kv : m)
      for (const auto& sn : kv.second.NodesByState) {
        auto it = state_id.find(sn.first);
        if (it == state_id.end()) {
          if (must_be_model)
            invalid("partition '" + kv.first + "' has state '" + sn.first + "' that is not in the model (the reference panics, plan.go:148)");
          continue;
        }
        cap[size_t(it->second)] = std::max(cap[size_t(it->second)], int32_t(deref(sn.second).size()));
      }
  };
  scan_caps(prevMap, false);
  scan_caps(partitionsToAssign, true);
  ip->state_slot_off.assign(size_t(S) + 1, 0);
  for (int32_t s = 0; s < S; ++s) ip->state_slot_off[size_t(s) + 1] = ip->state_slot_off[size_t(s)] + cap[size_t(s)];
  const int32_t SL = ip->state_slot_off[size_t(S)];

  // ---- rows
  ip->part_in_prev.assign(size_t(PU), 0);
  ip->part_in_assign.assign(size_t(PU), 0);
  ip->prev_rows.assign(size_t(PU) * size_t(SL), BLANCE_NO_NODE);
  ip->cur_rows.assign(size_t(PU) * size_t(SL), BLANCE_NO_NODE);
  ip->prev_shape.assign(size_t(PU) * size_t(S), BLANCE_SHAPE_ABSENT);
  ip->cur_shape.assign(size_t(PU) * size_t(S), BLANCE_SHAPE_ABSENT);
  ip->part_weight.assign(size_t(PU), 1);
  ip->part_has_weight.assign(size_t(PU), 0);
  ip->part_name_rank.resize(size_t(PU));
  for (int32_t p = 0; p < PU; ++p) ip->part_name_rank[size_t(p)] = p;
  if (options.PartitionWeights)
    for (const auto& kv : *options.PartitionWeights) {
      auto it = part_id.find(kv.first);
      if (it == part_id.end()) continue;
      ip->part_weight[size_t(it->second)] = kv.second;
      ip->part_has_weight[size_t(it->second)] = 1;
    }

  struct Extra { int32_t part; int32_t node; };
  std::vector<Extra> extras;   // prevMap entries under non-model states (only feed tot)
  auto fill = [&](const PartitionMap& m, std::vector<int32_t>& rows, std::vector<uint8_t>& shape,
                  std::vector<uint8_t>& present, bool is_prev) {
    for (const auto& kv : m) {
      const int32_t p = part_id.at(kv.first);
      present[size_t(p)] = 1;
      for (const auto& sn : kv.second.NodesByState) {
        auto it = state_id.find(sn.first);
        if (it == state_id.end()) {
          if (is_prev)
            for (const auto& n : deref(sn.second)) extras.push_back({p, nodes.get(n)});
          continue;
        }
        const int32_t s = it->second;
        shape[size_t(p) * size_t(S) + size_t(s)] = sn.second ? BLANCE_SHAPE_LIST : BLANCE_SHAPE_NIL;
        int32_t slot = ip->state_slot_off[size_t(s)];
        for (const auto& n : deref(sn.second)) rows[size_t(p) * size_t(SL) + size_t(slot++)] = nodes.get(n);
      }
    }
  };
  fill(prevMap, ip->prev_rows, ip->prev_shape, ip->part_in_prev, true);
  fill(partitionsToAssign, ip->cur_rows, ip->cur_shape, ip->part_in_assign, false);

  // ---- node flags (after every name that can occur has been interned)
  for (const auto& n : deref(nodesToRemove)) nodes.get(n);
  for (const auto& n : deref(nodesToAdd)) nodes.get(n);
  const int32_t NU = int32_t(nodes.names.size());
  ip->node_removed.assign(size_t(NU), 0);
  ip->node_added.assign(size_t(NU), 0);
  for (const auto& n : deref(nodesToRemove)) ip->node_removed[size_t(nodes.find(n))] = 1;
  for (const auto& n : deref(nodesToAdd)) ip->node_added[size_t(nodes.find(n))] = 1;
  ip->node_weight.assign(size_t(N), 0);
  ip->node_has_weight.assign(size_t(N), 0);
  if (options.NodeWeights)
    for (const auto& kv : *options.NodeWeights) {
      int32_t id = nodes.find(kv.first);
      if (id < 0 || id >= N) continue;
      ip->node_weight[size_t(id)] = kv.second;
      ip->node_has_weight[size_t(id)] = 1;
    }

  // plan.go:544-545 dereferences prevMap[name] whenever nodesToRemove is non-empty
  if (!deref(nodesToRemove).empty())
    for (int32_t p = 0; p < PU; ++p)
      if (ip->part_in_assign[size_t(p)] && !ip->part_in_prev[size_t(p)])
        invalid("partition '" + ip->part_names[size_t(p)] + "' is being assigned with nodesToRemove set but is missing from prevMap (the reference panics, plan.go:544)");

  // ---- counts under non-model states
  ip->extra_tot_first.assign(size_t(N), 0);
  ip->extra_tot_rest.assign(size_t(N), 0);
  for (const auto& e : extras) {
    if (e.node >= N) continue;
    long long w = (options.PartitionWeights && ip->part_has_weight[size_t(e.part)]) ? ip->part_weight[size_t(e.part)] : 1;
    ip->extra_tot_first[size_t(e.node)] = checked_i32((long long)ip->extra_tot_first[size_t(e.node)] + w, "count");
    if (!ip->part_in_assign[size_t(e.part)])
      ip->extra_tot_rest[size_t(e.node)] = checked_i32((long long)ip->extra_tot_rest[size_t(e.node)] + w, "count");
  }

  // int32 is wide enough for every count the device keeps: sum |w_p| * slots
  {
    long long bound = 0;
    for (int32_t p = 0; p < PU; ++p) {
      long long w = (options.PartitionWeights && ip->part_has_weight[size_t(p)]) ? ip->part_weight[size_t(p)] : 1;
      bound += (w < 0 ? -w : w) * std::max<long long>(1, SL);
    }
    if (bound > INT32_MAX) invalid("sum of partition weights x slots exceeds int32 (the device keeps int32 counts)");
  }

  // ---- hierarchy bit sets
  ip->rule_off.assign(size_t(S) + 1, 0);
  int32_t n_rules = 0, n_hier_bits = N;
  if (options.HierarchyRules) {
    std::vector<HierarchyRule> rules;
    for (int32_t s = 0; s < S; ++s) {
      auto it = options.HierarchyRules->find(ip->state_names[size_t(s)]);
      if (it != options.HierarchyRules->end())
        for (const auto& r : it->second) rules.push_back(r);
      ip->rule_off[size_t(s) + 1] = int32_t(rules.size());
    }
    n_rules = int32_t(rules.size());
    if (n_rules > 0) {
      Hierarchy h(options.NodeHierarchy ? &*options.NodeHierarchy : nullptr);
      // pass 1: the lists, and the leaf names outside nodesAll
      Interner extra_bits;
      std::vector<std::vector<int32_t>> lists(size_t(n_rules) * size_t(NU + 1));
      for (int32_t r = 0; r < n_rules; ++r)
        for (int32_t a = 0; a <= NU; ++a) {
          const std::string anchor = a < NU ? nodes.names[size_t(a)] : std::string();
          const Strs& inc = h.leaves(h.ancestor(anchor, rules[size_t(r)].IncludeLevel));
          const Strs& exc = h.leaves(h.ancestor(anchor, rules[size_t(r)].ExcludeLevel));
          std::unordered_set<std::string> ex(exc.begin(), exc.end());
          auto& out = lists[size_t(r) * size_t(NU + 1) + size_t(a)];
          for (const auto& leaf : inc) {
            if (ex.count(leaf)) continue;                          // plan.go:733
            int32_t id = nodes.find(leaf);
            if (id >= 0 && id < N) out.push_back(id);
            else out.push_back(N + extra_bits.get(leaf));
          }
        }
      n_hier_bits = N + int32_t(extra_bits.names.size());
      const size_t HW = size_t((n_hier_bits + 31) / 32);
      ip->ie_mask.assign(size_t(n_rules) * size_t(NU + 1) * HW, 0u);
      for (size_t i = 0; i < lists.size(); ++i)
        for (int32_t b : lists[i]) ip->ie_mask[i * HW + size_t(b >> 5)] |= 1u << (b & 31);
    }
  }

  ip->node_names = nodes.names;

  in.n_nodes = N; in.n_node_ids = NU; in.n_states = S; in.n_parts = PU; in.n_slots = SL;
  in.max_iters = options.MaxIterationsPerPlan;
  in.top_state = top_state < 0 ? 0 : top_state;
  in.booster_kind = options.NodeScoreBooster;
  in.add_is_nil = nodesToAdd ? 0 : 1;
  in.has_part_weights = options.PartitionWeights ? 1 : 0;
  in.has_node_weights = options.NodeWeights ? 1 : 0;
  in.has_hier_rules = options.HierarchyRules ? 1 : 0;
  in.state_priority = ip->state_priority.data();
  in.state_constraints = ip->state_constraints.data();
  in.state_slot_off = ip->state_slot_off.data();
  in.state_stickiness = ip->state_stickiness.data();
  in.state_has_stickiness = ip->state_has_stickiness.data();
  in.node_removed = ip->node_removed.data();
  in.node_added = ip->node_added.data();
  in.node_weight = ip->node_weight.data();
  in.node_has_weight = ip->node_has_weight.data();
  in.part_in_prev = ip->part_in_prev.data();
  in.part_in_assign = ip->part_in_assign.data();
  in.part_weight = ip->part_weight.data();
  in.part_has_weight = ip->part_has_weight.data();
  in.part_name_rank = ip->part_name_rank.data();
  in.prev_rows = ip->prev_rows.data();
  in.prev_shape = ip->prev_shape.data();
  in.cur_rows = ip->cur_rows.data();
  in.cur_shape = ip->cur_shape.data();
  in.extra_tot_first = ip->extra_tot_first.data();
  in.extra_tot_rest = ip->extra_tot_rest.data();
  in.n_rules = n_rules;
  in.n_hier_bits = n_hier_bits;
  in.rule_off = ip->rule_off.data();
  in.ie_mask = ip->ie_mask.empty() ? nullptr : ip->ie_mask.data();
  in.engine = options.Engine;
  return ip;
}

PlanOutBuffers::PlanOutBuffers(const InternedPlan& ip) {
  next_rows.assign(size_t(ip.in.n_parts) * size_t(ip.in.n_slots) + 1, BLANCE_NO_NODE);
  next_shape.assign(size_t(ip.in.n_parts) * size_t(ip.in.n_states) + 1, 0);
  warn.assign(size_t(ip.in.n_parts) * size_t(ip.in.n_states) + 1, 0);
  out.next_rows = next_rows.data();
  out.next_shape = next_shape.data();
  out.warn = warn.data();
}

PartitionMap UninternPlan(const InternedPlan& ip, const PlanOutBuffers& ob, Warnings* warnings) {
  const blance_plan_in& in = ip.in;
  PartitionMap next;
  next.reserve(size_t(in.n_parts));
  for (int32_t p = 0; p < in.n_parts; ++p) {
    if (!ip.part_in_assign[size_t(p)]) continue;                    // plan.go:326-330
    Partition part;
    part.Name = ip.part_names[size_t(p)];
    const int32_t* row = ob.next_rows.data() + size_t(p) * size_t(in.n_slots);
    for (int32_t s = 0; s < in.n_states; ++s) {
      const uint8_t sh = ob.next_shape[size_t(p) * size_t(in.n_states) + size_t(s)];
      if (sh == BLANCE_SHAPE_ABSENT) continue;
      if (sh == BLANCE_SHAPE_NIL) { part.NodesByState[ip.state_names[size_t(s)]] = std::nullopt; continue; }
      Strs list;
      for (int32_t i = ip.state_slot_off[size_t(s)]; i < ip.state_slot_off[size_t(s) + 1] && row[i] != BLANCE_NO_NODE; ++i)
        list.push_back(ip.node_names[size_t(row[i])]);
      part.NodesByState[ip.state_names[size_t(s)]] = std::move(list);
    }
    if (warnings)
      for (int32_t s = 0; s < in.n_states; ++s)
        if (ob.warn[size_t(p) * size_t(in.n_states) + size_t(s)]) {
          char buf[32];                                              // plan.go:231-234
          std::snprintf(buf, sizeof buf, "%d", ip.state_constraints[size_t(s)]);
          (*warnings)[part.Name].push_back(std::string("could not meet constraints: ") + buf +
                                           ", stateName: " + ip.state_names[size_t(s)] +
                                           ", partitionName: " + part.Name);
        }
    next.emplace(part.Name, std::move(part));
  }
  return next;
}

blance_ctx* DefaultContext() {
  static std::mutex mu;
  static blance_ctx* ctx = nullptr;
  std::lock_guard<std::mutex> g(mu);
  if (!ctx) {
    int st = blance_ctx_create(&ctx, -1);
    if (st != BLANCE_OK) {
      ctx = nullptr;
      throw BlanceError(st, std::string("blance_ctx_create failed: ") + blance_last_error(nullptr));
    }
  }
  return ctx;
}

PartitionMap PlanNextMapEx(PartitionMap& prevMap, PartitionMap& partitionsToAssign, const Strs& nodesAll,
                           const OptStrs& nodesToRemove, const OptStrs& nodesToAdd,
                           const PartitionModel& model, const PlanNextMapOptions& options,
                           Warnings* warnings, PlanStats* stats) {
  if (warnings) warnings->clear();
  auto ip = InternPlan(prevMap, partitionsToAssign, nodesAll, nodesToRemove, nodesToAdd, model, options);
  PlanOutBuffers ob(*ip);
  blance_ctx* ctx = DefaultContext();
  int st = blance_plan_next_map(ctx, &ip->in, &ob.out);
  if (st != BLANCE_OK) throw BlanceError(st, std::string("blance_plan_next_map failed: ") + blance_last_error(ctx));
  if (stats) {
    stats->iters_run = ob.out.iters_run; stats->converged = ob.out.converged; stats->steps = ob.out.steps;
    stats->device_ms = ob.out.device_ms; stats->kernel_ms = ob.out.kernel_ms; stats->pass_ms = ob.out.pass_ms;
  }
  if (ob.out.iters_run <= 0) return PartitionMap{};                  // MaxIterationsPerPlan <= 0: plan.go:32,57
  PartitionMap next = UninternPlan(*ip, ob, warnings);
  // plan.go:49-52: after any non-matching iteration the caller's maps hold the new
  // partitions; when the loop ends their content equals the returned map.
  if (ob.out.iters_run >= 2 || !ob.out.converged) {
    const bool same = &prevMap == &partitionsToAssign;
    for (const auto& kv : next) {
      prevMap[kv.first] = kv.second;
      if (!same) partitionsToAssign[kv.first] = kv.second;
    }
  }
  return next;
}

// ------------------------------------------------------------------------------------
// CalcPartitionMoves

namespace {

struct MovesTables {
  Interner nodes;
  Strs state_names;      // visit states first, then the other keys that occur
  int32_t n_visit = 0;
  std::vector<int32_t> slot_off, beg_rows, end_rows;
  Strs part_names;
};

void intern_moves(const Strs& states, const std::vector<const NodesByState*>& begs,
                  const std::vector<const NodesByState*>& ends, MovesTables* t) {
  std::unordered_map<std::string, int32_t> sid;
  for (const auto& s : states) {
    if (sid.count(s)) invalid("CalcPartitionMoves: state '" + s + "' listed twice");
    sid[s] = int32_t(t->state_names.size());
    t->state_names.push_back(s);
  }
  t->n_visit = int32_t(states.size());
  static const NodesByState kEmptyNbs;
  auto scan = [&](const NodesByState* nbs, std::vector<int32_t>& cap) {
    if (!nbs) return;
    for (const auto& kv : *nbs) {
      auto it = sid.find(kv.first);
      if (it == sid.end()) { it = sid.emplace(kv.first, int32_t(t->state_names.size())).first; t->state_names.push_back(kv.first); cap.push_back(0); }
      cap[size_t(it->second)] = std::max(cap[size_t(it->second)], int32_t(deref(kv.second).size()));
    }
  };
  std::vector<int32_t> cap(t->state_names.size(), 0);
  for (auto* b : begs) scan(b, cap);
  for (auto* e : ends) scan(e, cap);
  const size_t S = t->state_names.size();
  t->slot_off.assign(S + 1, 0);
  for (size_t s = 0; s < S; ++s) t->slot_off[s + 1] = t->slot_off[s] + cap[s];
  const size_t SL = size_t(t->slot_off[S]), P = begs.size();
  t->beg_rows.assign(P * SL + 1, BLANCE_NO_NODE);
  t->end_rows.assign(P * SL + 1, BLANCE_NO_NODE);
  auto fill = [&](const NodesByState* nbs, int32_t* row) {
    if (!nbs) return;
    for (const auto& kv : *nbs) {
      int32_t slot = t->slot_off[size_t(sid.at(kv.first))];
      for (const auto& n : deref(kv.second)) row[slot++] = t->nodes.get(n);
    }
  };
  for (size_t p = 0; p < P; ++p) { fill(begs[p], t->beg_rows.data() + p * SL); fill(ends[p], t->end_rows.data() + p * SL); }
}

std::vector<std::vector<NodeStateOp>> run_moves(MovesTables& t, size_t P, bool favorMinNodes) {
  const size_t S = t.state_names.size(), SL = size_t(t.slot_off[S]);
  const int32_t max_ops = int32_t(std::max<size_t>(1, 2 * SL));
  std::vector<int32_t> op_node(P * size_t(max_ops) + 1), op_count(P + 1);
  std::vector<uint8_t> op_state(P * size_t(max_ops) + 1), op_kind(P * size_t(max_ops) + 1);
  blance_ctx* ctx = DefaultContext();
  int st = blance_calc_partition_moves(ctx, int32_t(P), int32_t(S), t.n_visit, t.slot_off.data(), t.beg_rows.data(),
                                       t.end_rows.data(), favorMinNodes ? 1 : 0, max_ops, op_node.data(),
                                       op_state.data(), op_kind.data(), op_count.data());
  if (st != BLANCE_OK) throw BlanceError(st, std::string("blance_calc_partition_moves failed: ") + blance_last_error(ctx));
  static const char* kOps[] = {"add", "del", "promote", "demote"};
  std::vector<std::vector<NodeStateOp>> out(P);
  for (size_t p = 0; p < P; ++p)
    for (int32_t i = 0; i < op_count[p]; ++i) {
      const size_t o = p * size_t(max_ops) + size_t(i);
      NodeStateOp op;
      op.Node = t.nodes.names[size_t(op_node[o])];
      op.State = op_state[o] == BLANCE_OP_STATE_NONE ? std::string() : t.state_names[op_state[o]];
      op.Op = kOps[op_kind[o]];
      out[p].push_back(std::move(op));
    }
  return out;
}

}  // namespace

std::vector<NodeStateOp> CalcPartitionMoves(const Strs& states, const NodesByState& beg, const NodesByState& end,
                                            bool favorMinNodes) {
  MovesTables t;
  intern_moves(states, {&beg}, {&end}, &t);
  return run_moves(t, 1, favorMinNodes)[0];
}

std::unordered_map<std::string, std::vector<NodeStateOp>> CalcPartitionMovesMap(
    const Strs& states, const PartitionMap& beg, const PartitionMap& end, bool favorMinNodes) {
  Strs names;
  std::unordered_set<std::string> seen;
  for (const auto& kv : beg) if (seen.insert(kv.first).second) names.push_back(kv.first);
  for (const auto& kv : end) if (seen.insert(kv.first).second) names.push_back(kv.first);
  std::sort(names.begin(), names.end());
  std::vector<const NodesByState*> begs, ends;
  for (const auto& n : names) {
    auto b = beg.find(n); auto e = end.find(n);
    begs.push_back(b == beg.end() ? nullptr : &b->second.NodesByState);
    ends.push_back(e == end.end() ? nullptr : &e->second.NodesByState);
  }
  MovesTables t;
  intern_moves(states, begs, ends, &t);
  auto ops = run_moves(t, names.size(), favorMinNodes);
  std::unordered_map<std::string, std::vector<NodeStateOp>> out;
  for (size_t p = 0; p < names.size(); ++p) out.emplace(names[p], std::move(ops[p]));
  return out;
}

}  // namespace blance
