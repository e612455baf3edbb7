// oracle/literal.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See literal.hpp.
//
// Restates plan.go / moves.go / misc.go of couchbase/blance with the same
// containers (hash maps of strings, slices) and the same comparator-driven
// sorts.  Wherever the Go code iterates a map in (random) map order the result
// is order independent except for the two cases SURVEY.md section 9 lists (D1:
// top-priority state on equal priorities, D2: inconsistent state comparator);
// for those this file fixes "ascending state name" as the starting order.
#include "literal.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <unordered_set>

namespace blance_oracle {

namespace {

const Strs kEmpty;

inline const Strs& deref(const OptStrs& s) { return s ? *s : kEmpty; }

// Go's fmt.Sprintf("%10d", v)
std::string pad10(long long v) {
  char buf[32];
  std::snprintf(buf, sizeof buf, "%10lld", v);
  return buf;
}

// strconv.Atoi: optional sign, then decimal digits only, no overflow of int64.
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
    if (acc > (lim - d) / 10) return false;   // would overflow int64
    acc = acc * 10 + d;
  }
  *out = neg ? -(long long)acc : (long long)acc;
  return true;
}

// Keys of a NodesByState in a fixed (ascending) order: every loop over the map
// below is order independent, this only makes runs reproducible.
Strs sorted_keys(const NodesByState& nbs) {
  Strs keys;
  keys.reserve(nbs.size());
  for (const auto& kv : nbs) keys.push_back(kv.first);
  std::sort(keys.begin(), keys.end());
  return keys;
}

NodesByState copy_nodes_by_state(const NodesByState& nbs) {   // plan.go:345-351
  NodesByState rv;
  for (const auto& kv : nbs) {
    // append([]string(nil), nodes...) yields nil for an empty source
    if (deref(kv.second).empty()) rv[kv.first] = std::nullopt;
    else rv[kv.first] = deref(kv.second);
  }
  return rv;
}

using StateNodeCounts = std::unordered_map<std::string, IntMap>;

void adjust_state_node_counts(StateNodeCounts& snc, const std::string& state,
                              const Strs& nodes, int64_t amt) {   // plan.go:353-363
  for (const auto& node : nodes) {
    IntMap& s = snc[state];
    s[node] = s[node] + amt;
  }
}

template <class Cb>
NodesByState remove_nodes_cb(const NodesByState& nbs, const OptStrs& remove, Cb cb, bool have_cb) {
  NodesByState rv;                                                // plan.go:408-421
  for (const auto& state : sorted_keys(nbs)) {
    const Strs& nodes = deref(nbs.at(state));
    if (have_cb) cb(state, strings_intersect_strings(nodes, remove));
    rv[state] = strings_remove_strings(nodes, remove);
  }
  return rv;
}

bool nodes_by_state_deep_equal(const NodesByState& a, const NodesByState& b) {
  if (a.size() != b.size()) return false;                          // reflect.DeepEqual on maps
  for (const auto& kv : a) {
    auto it = b.find(kv.first);
    if (it == b.end()) return false;
    if (kv.second.has_value() != it->second.has_value()) return false;   // nil vs non-nil slice
    if (kv.second && *kv.second != *it->second) return false;
  }
  return true;
}

// --- partitionSorter, plan.go:481-562 -------------------------------------
struct PartitionSorter {
  std::string state_name;
  const PartitionMap* prev_map = nullptr;
  const OptStrs* nodes_to_remove = nullptr;
  const OptStrs* nodes_to_add = nullptr;
  const std::optional<IntMap>* partition_weights = nullptr;

  std::vector<std::string> score(const Partition& p) const {       // plan.go:519-562
    const std::string& name = p.name;
    std::string name_str = name;
    long long n;
    if (go_atoi(name, &n) && n >= 0) name_str = pad10(n);
    long long weight = 1;
    if (partition_weights && partition_weights->has_value()) {
      auto it = (*partition_weights)->find(name);
      if (it != (*partition_weights)->end()) weight = it->second;
    }
    std::string weight_str = pad10(999999999LL - weight);
    if (prev_map && nodes_to_remove && nodes_to_remove->has_value() && !(*nodes_to_remove)->empty()) {
      auto it = prev_map->find(name);
      if (it == prev_map->end() || !it->second) {
        // the reference dereferences a nil *Partition here (plan.go:544-545)
        std::fprintf(stderr, "blance oracle: partition %s missing from prevMap while nodesToRemove is non-empty "
                             "(the reference panics)\n", name.c_str());
        std::abort();
      }
      auto lp = it->second->nodes_by_state.find(state_name);
      if (lp != it->second->nodes_by_state.end() && lp->second.has_value() &&
          !strings_intersect_strings(*lp->second, *nodes_to_remove).empty())
        return {"0", weight_str, name_str};
    }
    if (nodes_to_add && nodes_to_add->has_value()) {
      Strs flat = flatten_nodes_by_state(p.nodes_by_state, {});
      if (strings_intersect_strings(flat, *nodes_to_add).empty())
        return {"1", weight_str, name_str};
    }
    return {"2", weight_str, name_str};
  }

  bool less(const Partition& a, const Partition& b) const {         // plan.go:495-513
    auto ei = score(a), ej = score(b);
    for (size_t x = 0; x < ei.size() && x < ej.size(); ++x) {
      if (ei[x] < ej[x]) return true;
      if (ei[x] > ej[x]) return false;
    }
    if (ei.size() != ej.size()) return ei.size() < ej.size();
    return a.name < b.name;
  }
};

// --- nodeSorter, plan.go:598-689 ------------------------------------------
struct NodeSorter {
  const std::string* state_name;
  const Partition* partition;
  int64_t num_partitions;
  const std::string* top_priority_node;
  const StateNodeCounts* state_node_counts;
  const std::unordered_map<std::string, IntMap>* node_to_node_counts;
  const IntMap* node_partition_counts;
  const IntMap* node_positions;
  const std::optional<IntMap>* node_weights;
  double stickiness;
  Booster booster;

  double score(const std::string& node) const {                     // plan.go:634-689
    double lower_priority_balance_factor = 0.0;
    if (node_to_node_counts && num_partitions > 0) {
      auto m = node_to_node_counts->find(*top_priority_node);
      if (m != node_to_node_counts->end()) {
        auto c = m->second.find(node);
        int64_t v = c == m->second.end() ? 0 : c->second;
        lower_priority_balance_factor = double(v) / double(num_partitions);
      }
    }
    double filled_factor = 0.0;
    if (node_partition_counts && num_partitions > 0) {
      auto c = node_partition_counts->find(node);
      if (c != node_partition_counts->end())
        filled_factor = (0.001 * double(c->second)) / double(num_partitions);
    }
    double current_factor = 0.0;
    if (partition) {
      auto it = partition->nodes_by_state.find(*state_name);
      if (it != partition->nodes_by_state.end())
        for (const auto& sn : deref(it->second))
          if (sn == node) current_factor = stickiness;
    }
    double r = 0.0;
    if (state_node_counts) {
      auto nc = state_node_counts->find(*state_name);
      if (nc != state_node_counts->end()) {
        auto c = nc->second.find(node);
        r = double(c == nc->second.end() ? 0 : c->second);
      }
    }
    r = r + lower_priority_balance_factor;
    r = r + filled_factor;
    if (node_weights->has_value()) {
      auto w = (*node_weights)->find(node);
      if (w != (*node_weights)->end()) {
        if (w->second > 0) {
          r = r / double(w->second);
        } else if (w->second < 0 && booster != Booster::None) {
          double boost = double(-w->second);                        // control_test.go:19-26
          if (boost < current_factor) boost = current_factor;
          r += boost;
        }
      }
    }
    r = r - current_factor;
    return r;
  }

  bool less(const std::string& a, const std::string& b) const {      // plan.go:617-628
    double si = score(a), sj = score(b);
    if (si < sj) return true;
    if (si > sj) return false;
    auto pa = node_positions->find(a), pb = node_positions->find(b);
    int64_t ia = pa == node_positions->end() ? 0 : pa->second;
    int64_t ib = pb == node_positions->end() ? 0 : pb->second;
    return ia < ib;
  }
};

// plan.go:723-734
Strs include_exclude_nodes(const std::string& node, int64_t include_level, int64_t exclude_level,
                           const StrMap& parents, const std::unordered_map<std::string, Strs>& children) {
  Strs inc = find_leaves(find_ancestor(node, parents, include_level), children);
  Strs exc = find_leaves(find_ancestor(node, parents, exclude_level), children);
  return strings_remove_strings(inc, exc);
}

// plan.go:738-753
Strs include_exclude_nodes_intersect(const Strs& nodes, int64_t include_level, int64_t exclude_level,
                                     const StrMap& parents,
                                     const std::unordered_map<std::string, Strs>& children) {
  Strs rv;
  for (const auto& node : nodes) {
    Strs res = include_exclude_nodes(node, include_level, exclude_level, parents, children);
    if (rv.empty()) { rv = std::move(res); continue; }
    rv = strings_intersect_strings(rv, res);
  }
  return rv;
}

struct InnerResult { PartitionMap next; Warnings warnings; int64_t steps = 0; };

// plan.go:60-331
InnerResult plan_next_map_inner(const PartitionMap& prev_map, const PartitionMap& partitions_to_assign,
                                const Strs& nodes_all, const OptStrs& nodes_to_remove,
                                const OptStrs& nodes_to_add, const PartitionModel& model,
                                const Options& opts) {
  InnerResult out;
  Warnings& partition_warnings = out.warnings;

  IntMap node_positions;                                            // plan.go:72-75
  for (size_t i = 0; i < nodes_all.size(); ++i) node_positions[nodes_all[i]] = int64_t(i);

  const Strs nodes_next = strings_remove_strings(nodes_all, nodes_to_remove);   // plan.go:77

  const StrMap empty_parents;
  const StrMap& parents = opts.node_hierarchy ? *opts.node_hierarchy : empty_parents;
  const auto hierarchy_children = map_parents_to_map_children(parents);          // plan.go:79

  // plan.go:83-89: deep copy minus the to-be-removed nodes, sorted by name.
  std::vector<PartitionPtr> next_partitions;
  next_partitions.reserve(partitions_to_assign.size());
  for (const auto& kv : partitions_to_assign) {
    auto p = std::make_shared<Partition>();
    p->name = kv.second->name;
    p->nodes_by_state = copy_nodes_by_state(kv.second->nodes_by_state);
    p->nodes_by_state = remove_nodes_cb(p->nodes_by_state, nodes_to_remove,
                                        [](const std::string&, const Strs&) {}, false);
    next_partitions.push_back(std::move(p));
  }
  {
    PartitionSorter by_name;   // stateName "", no prevMap/add/remove/weights
    if (opts.memoize_partition_scores) {
      std::vector<std::pair<std::vector<std::string>, PartitionPtr>> keyed;
      keyed.reserve(next_partitions.size());
      for (auto& p : next_partitions) keyed.emplace_back(by_name.score(*p), p);
      std::sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) {
        if (a.first != b.first) return a.first < b.first;
        return a.second->name < b.second->name;
      });
      for (size_t x = 0; x < next_partitions.size(); ++x) next_partitions[x] = keyed[x].second;
    } else
    std::sort(next_partitions.begin(), next_partitions.end(),
              [&](const PartitionPtr& a, const PartitionPtr& b) { return by_name.less(*a, *b); });
  }

  StateNodeCounts state_node_counts = count_state_nodes(prev_map, opts.partition_weights);   // plan.go:94

  // D1: the reference takes whichever minimum-priority state Go's map order
  // yields first; fixed here as the first such state in ascending name order.
  std::string top_priority_state_name;
  {
    Strs names;
    for (const auto& kv : model) names.push_back(kv.first);
    std::sort(names.begin(), names.end());
    bool have = false;
    for (const auto& n : names) {                                    // plan.go:126-132
      if (!have || model.at(n).priority < model.at(top_priority_state_name).priority) {
        top_priority_state_name = n;
        have = true;
      }
    }
  }

  auto find_best_nodes = [&](Partition& partition, const std::string& state_name, int64_t constraints,
                             std::unordered_map<std::string, IntMap>& node_to_node_counts) -> OptStrs {
    out.steps++;
    double stickiness = 1.5;                                        // plan.go:104-115
    if (opts.partition_weights) {
      auto w = opts.partition_weights->find(partition.name);
      if (w != opts.partition_weights->end()) {
        stickiness = double(w->second);
      } else if (opts.state_stickiness) {
        auto s = opts.state_stickiness->find(state_name);
        if (s != opts.state_stickiness->end()) stickiness = double(s->second);
      }
    }

    IntMap node_partition_counts;                                   // plan.go:118-124
    for (const auto& sc : state_node_counts)
      for (const auto& nc : sc.second) node_partition_counts[nc.first] += nc.second;

    std::string top_priority_node;                                  // plan.go:134-138
    {
      auto it = partition.nodes_by_state.find(top_priority_state_name);
      if (it != partition.nodes_by_state.end() && !deref(it->second).empty())
        top_priority_node = deref(it->second)[0];
    }

    const int64_t state_priority = model.at(state_name).priority;   // plan.go:140

    // plan.go:142: append([]string(nil), nodesNext...) is nil when nodesNext is empty.
    OptStrs candidate_nodes;
    if (!nodes_next.empty()) candidate_nodes = nodes_next;

    auto exclude_higher_priority_nodes = [&](OptStrs remaining) -> OptStrs {   // plan.go:146-154
      for (const auto& sn : sorted_keys(partition.nodes_by_state)) {
        auto ms = model.find(sn);
        if (ms == model.end()) {
          std::fprintf(stderr, "blance oracle: state %s of partition %s is not in the model "
                               "(the reference panics)\n", sn.c_str(), partition.name.c_str());
          std::abort();
        }
        if (ms->second.priority < state_priority)
          remaining = strings_remove_strings(deref(remaining), partition.nodes_by_state.at(sn));
      }
      return remaining;
    };
    candidate_nodes = exclude_higher_priority_nodes(candidate_nodes);   // plan.go:156

    NodeSorter ns{&state_name, &partition, int64_t(prev_map.size()), &top_priority_node,
                  &state_node_counts, &node_to_node_counts, &node_partition_counts, &node_positions,
                  &opts.node_weights, stickiness, opts.booster};
    auto sort_nodes = [&](Strs& v) {                                // plan.go:171-172 / 211-212
      std::sort(v.begin(), v.end(),
                [&](const std::string& a, const std::string& b) { return ns.less(a, b); });
    };
    if (candidate_nodes) sort_nodes(*candidate_nodes);

    if (opts.hierarchy_rules) {                                     // plan.go:174-226
      Strs hierarchy_nodes;
      auto rules = opts.hierarchy_rules->find(state_name);
      if (rules != opts.hierarchy_rules->end()) {
        for (const auto& rule : rules->second) {
          std::string h = top_priority_node;
          if (h.empty() && !hierarchy_nodes.empty()) h = hierarchy_nodes[0];
          for (int64_t i = 0; i < constraints; ++i) {
            Strs anchors;
            anchors.push_back(h);
            anchors.insert(anchors.end(), hierarchy_nodes.begin(), hierarchy_nodes.end());
            Strs hc = include_exclude_nodes_intersect(anchors, rule.include_level, rule.exclude_level,
                                                      parents, hierarchy_children);
            hc = strings_intersect_strings(hc, nodes_next);
            hc = deref(exclude_higher_priority_nodes(OptStrs(hc)));
            sort_nodes(hc);
            if (!hc.empty()) hierarchy_nodes.push_back(hc[0]);
            else if (!deref(candidate_nodes).empty()) hierarchy_nodes.push_back(deref(candidate_nodes)[0]);
          }
        }
      }
      Strs merged = hierarchy_nodes;                                // plan.go:224-225
      merged.insert(merged.end(), deref(candidate_nodes).begin(), deref(candidate_nodes).end());
      candidate_nodes = strings_deduplicate(merged);
    }

    if (int64_t(deref(candidate_nodes).size()) >= constraints) {    // plan.go:228-235
      candidate_nodes = Strs(candidate_nodes->begin(), candidate_nodes->begin() + constraints);
    } else {
      char buf[64];
      std::snprintf(buf, sizeof buf, "%lld", (long long)constraints);
      partition_warnings[partition.name].push_back(
          std::string("could not meet constraints: ") + buf + ", stateName: " + state_name +
          ", partitionName: " + partition.name);
    }

    for (const auto& c : deref(candidate_nodes)) {                  // plan.go:238-245
      IntMap& m = node_to_node_counts[top_priority_node];
      m[c] = m[c] + 1;
    }
    return candidate_nodes;
  };

  auto assign_state_to_partitions = [&](const std::string& state_name, int64_t constraints) {
    PartitionSorter ps;                                              // plan.go:255-263
    ps.state_name = state_name;
    ps.prev_map = &prev_map;
    ps.nodes_to_remove = &nodes_to_remove;
    ps.nodes_to_add = &nodes_to_add;
    ps.partition_weights = &opts.partition_weights;
    std::vector<PartitionPtr> order = next_partitions;
    if (opts.memoize_partition_scores) {
      std::vector<std::pair<std::vector<std::string>, PartitionPtr>> keyed;
      keyed.reserve(order.size());
      for (auto& p : order) keyed.emplace_back(ps.score(*p), p);
      std::sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) {
        if (a.first != b.first) return a.first < b.first;            // element-wise, then by length: plan.go:498-511
        return a.second->name < b.second->name;
      });
      for (size_t x = 0; x < order.size(); ++x) order[x] = keyed[x].second;
    } else
    std::sort(order.begin(), order.end(),
              [&](const PartitionPtr& a, const PartitionPtr& b) { return ps.less(*a, *b); });

    std::unordered_map<std::string, IntMap> node_to_node_counts;    // plan.go:266

    int64_t done = 0;
    auto slice_t0 = std::chrono::steady_clock::now();
    int64_t in_slice = 0;
    for (auto& partition : order) {                                  // plan.go:268-302
      if (opts.max_steps_per_pass >= 0 && done++ >= opts.max_steps_per_pass) break;
      if (opts.slice_seconds && opts.slice_steps > 0 && in_slice == opts.slice_steps) {
        const auto now = std::chrono::steady_clock::now();
        opts.slice_seconds->push_back(std::chrono::duration<double>(now - slice_t0).count());
        slice_t0 = now;
        in_slice = 0;
      }
      ++in_slice;
      int64_t partition_weight = 1;
      if (opts.partition_weights) {
        auto w = opts.partition_weights->find(partition->name);
        if (w != opts.partition_weights->end()) partition_weight = w->second;
      }
      auto dec = [&](const std::string& sn, const Strs& nodes) {
        adjust_state_node_counts(state_node_counts, sn, nodes, -partition_weight);
      };

      OptStrs nodes_to_assign = find_best_nodes(*partition, state_name, constraints, node_to_node_counts);

      OptStrs old_nodes;   // partition.NodesByState[stateName]; nil when the key is absent
      {
        auto it = partition->nodes_by_state.find(state_name);
        if (it != partition->nodes_by_state.end()) old_nodes = it->second;
      }
      partition->nodes_by_state = remove_nodes_cb(partition->nodes_by_state, old_nodes, dec, true);
      partition->nodes_by_state = remove_nodes_cb(partition->nodes_by_state, nodes_to_assign, dec, true);
      partition->nodes_by_state[state_name] = nodes_to_assign;       // plan.go:299
      adjust_state_node_counts(state_node_counts, state_name, deref(nodes_to_assign), partition_weight);
    }
    if (opts.slice_seconds && opts.slice_steps > 0 && in_slice == opts.slice_steps)
      opts.slice_seconds->push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - slice_t0).count());
  };

  for (const auto& state_name : sort_state_names(model)) {          // plan.go:307-324
    int64_t constraints = model.at(state_name).constraints;
    if (opts.model_state_constraints) {
      auto c = opts.model_state_constraints->find(state_name);
      if (c != opts.model_state_constraints->end()) constraints = c->second;
    }
    if (constraints > 0) assign_state_to_partitions(state_name, constraints);
  }

  for (auto& p : next_partitions) out.next[p->name] = p;            // plan.go:326-330
  return out;
}

}  // namespace

// --- misc.go -----------------------------------------------------------------

Strs strings_remove_strings(const Strs& a, const OptStrs& remove) {   // misc.go:27-36
  std::unordered_set<std::string> rm;
  if (remove) rm.insert(remove->begin(), remove->end());
  Strs rv;
  rv.reserve(a.size());
  for (const auto& s : a)
    if (!rm.count(s)) rv.push_back(s);
  return rv;
}

Strs strings_intersect_strings(const Strs& a, const OptStrs& b) {     // misc.go:40-51
  std::unordered_set<std::string> bm, seen;
  if (b) bm.insert(b->begin(), b->end());
  Strs rv;
  rv.reserve(a.size());
  for (const auto& s : a)
    if (bm.count(s) && !seen.count(s)) { seen.insert(s); rv.push_back(s); }
  return rv;
}

Strs strings_deduplicate(const Strs& a) {                             // misc.go:55-66
  std::unordered_set<std::string> seen;
  Strs rv;
  for (const auto& s : a)
    if (seen.insert(s).second) rv.push_back(s);
  return rv;
}

// --- plan.go helpers -----------------------------------------------------------

Strs flatten_nodes_by_state(const NodesByState& nbs, const Strs& state_order) {   // plan.go:425-431
  Strs rv;
  std::unordered_set<std::string> done;
  for (const auto& s : state_order) {
    auto it = nbs.find(s);
    if (it == nbs.end() || !done.insert(s).second) continue;
    rv.insert(rv.end(), deref(it->second).begin(), deref(it->second).end());
  }
  for (const auto& s : sorted_keys(nbs)) {
    if (done.count(s)) continue;
    const Strs& v = deref(nbs.at(s));
    rv.insert(rv.end(), v.begin(), v.end());
  }
  return rv;
}

NodesByState remove_nodes_from_nodes_by_state(const NodesByState& nbs, const OptStrs& remove) {
  return remove_nodes_cb(nbs, remove, [](const std::string&, const Strs&) {}, false);
}

static bool state_name_less(const PartitionModel* m, const std::string& i, const std::string& j) {
  if (m) {                                                            // plan.go:459-470
    auto a = m->find(i), b = m->find(j);
    if (a != m->end() && b != m->end() && a->second.priority < b->second.priority) return true;
  }
  return i < j;
}

// Go's sort.Sort runs plain insertion sort below 12 elements; the state
// comparator is not a consistent order (D2), so the algorithm is part of the
// observable behaviour for the unit table (plan_test.go:117-176).
void state_name_insertion_sort(const PartitionModel* model, Strs& s) {
  for (size_t i = 1; i < s.size(); ++i)
    for (size_t j = i; j > 0 && state_name_less(model, s[j], s[j - 1]); --j) std::swap(s[j], s[j - 1]);
}

Strs sort_state_names(const PartitionModel& model) {                  // plan.go:437-447
  Strs names;
  for (const auto& kv : model) names.push_back(kv.first);
  std::sort(names.begin(), names.end());   // stands in for Go's random map order (D2)
  state_name_insertion_sort(&model, names);
  return names;
}

std::unordered_map<std::string, IntMap> count_state_nodes(const PartitionMap& m,
                                                          const std::optional<IntMap>& weights) {
  std::unordered_map<std::string, IntMap> rv;                         // plan.go:374-399
  for (const auto& kv : m) {
    for (const auto& sn : kv.second->nodes_by_state) {
      IntMap& s = rv[sn.first];
      for (const auto& node : deref(sn.second)) {
        int64_t w = 1;
        if (weights) {
          auto it = weights->find(kv.first);
          if (it != weights->end()) w = it->second;
        }
        s[node] = s[node] + w;
      }
    }
  }
  return rv;
}

std::unordered_map<std::string, Strs> map_parents_to_map_children(const StrMap& parents) {
  Strs nodes;                                                         // plan.go:703-717
  for (const auto& kv : parents) nodes.push_back(kv.first);
  std::sort(nodes.begin(), nodes.end());
  std::unordered_map<std::string, Strs> rv;
  for (const auto& child : nodes) rv[parents.at(child)].push_back(child);
  return rv;
}

std::string find_ancestor(std::string node, const StrMap& parents, int64_t level) {
  while (level > 0) {                                                 // plan.go:755-762
    auto it = parents.find(node);
    node = it == parents.end() ? std::string() : it->second;
    --level;
  }
  return node;
}

Strs find_leaves(const std::string& node, const std::unordered_map<std::string, Strs>& children) {
  auto it = children.find(node);                                      // plan.go:764-774
  if (it == children.end() || it->second.empty()) return {node};
  Strs rv;
  for (const auto& c : it->second) {
    Strs sub = find_leaves(c, children);
    rv.insert(rv.end(), sub.begin(), sub.end());
  }
  return rv;
}

// --- plan.go:23-58 -------------------------------------------------------------

PlanResult plan_next_map_ex(PartitionMap& prev_map, PartitionMap& partitions_to_assign,
                            Strs nodes_all, OptStrs nodes_to_remove, OptStrs nodes_to_add,
                            const PartitionModel& model, const Options& opts) {
  PlanResult res;
  for (int i = 0; i < opts.max_iterations; ++i) {
    InnerResult inner = plan_next_map_inner(prev_map, partitions_to_assign, nodes_all,
                                            nodes_to_remove, nodes_to_add, model, opts);
    res.next_map = std::move(inner.next);
    res.warnings = std::move(inner.warnings);
    res.iterations = i + 1;
    res.steps += inner.steps;
    bool not_match = false;                                           // plan.go:36-42
    for (const auto& kv : res.next_map) {
      auto it = prev_map.find(kv.first);
      if (it == prev_map.end() || !it->second ||
          !nodes_by_state_deep_equal(kv.second->nodes_by_state, it->second->nodes_by_state) ||
          kv.second->name != it->second->name) {
        not_match = true;
        break;
      }
    }
    if (!not_match) break;
    for (const auto& kv : res.next_map) {                             // plan.go:49-52
      prev_map[kv.first] = kv.second;
      partitions_to_assign[kv.first] = kv.second;
    }
    nodes_all = strings_remove_strings(nodes_all, nodes_to_remove);   // plan.go:53-55
    nodes_to_remove = Strs{};
    nodes_to_add = Strs{};
  }
  return res;
}

// --- moves.go ---------------------------------------------------------------------

static const Strs& nbs_get(const NodesByState& nbs, const std::string& state) {
  auto it = nbs.find(state);
  return it == nbs.end() ? kEmpty : deref(it->second);
}

Strs find_state_changes(int beg_state_idx, int end_state_idx, const std::string& state,
                        const Strs& states, const NodesByState& beg, const NodesByState& end) {
  Strs rv;                                                            // moves.go:121-136
  for (const auto& node : nbs_get(end, state))
    for (int i = beg_state_idx; i < end_state_idx; ++i)
      for (const auto& n : nbs_get(beg, states[size_t(i)]))
        if (n == node) rv.push_back(node);
  return rv;
}

std::vector<NodeStateOp> calc_partition_moves(const Strs& states, const NodesByState& beg,
                                              const NodesByState& end, bool favor_min_nodes) {
  std::vector<NodeStateOp> moves;                                     // moves.go:41-119
  std::unordered_set<std::string> seen;
  auto add_moves = [&](const Strs& nodes, const std::string& state, const char* op) {
    for (const auto& node : nodes)
      if (seen.insert(node).second) moves.push_back({node, state, op});
  };
  Strs beg_nodes = flatten_nodes_by_state(beg, states);
  Strs end_nodes = flatten_nodes_by_state(end, states);
  Strs adds = strings_remove_strings(end_nodes, beg_nodes);
  Strs dels = strings_remove_strings(beg_nodes, end_nodes);
  const int n = int(states.size());
  auto promote = [&](int si) {
    add_moves(find_state_changes(si + 1, n, states[size_t(si)], states, beg, end), states[size_t(si)], "promote");
  };
  auto demote = [&](int si) {
    add_moves(find_state_changes(0, si, states[size_t(si)], states, beg, end), states[size_t(si)], "demote");
  };
  auto clean_add = [&](int si) {
    const std::string& s = states[size_t(si)];
    add_moves(strings_intersect_strings(strings_remove_strings(nbs_get(end, s), nbs_get(beg, s)), adds), s, "add");
  };
  auto clean_del = [&](int si) {
    const std::string& s = states[size_t(si)];
    add_moves(strings_intersect_strings(strings_remove_strings(nbs_get(beg, s), nbs_get(end, s)), dels), "", "del");
  };
  if (!favor_min_nodes) {
    for (int si = 0; si < n; ++si) { promote(si); demote(si); clean_add(si); clean_del(si); }
  } else {
    for (int si = n - 1; si >= 0; --si) { clean_del(si); demote(si); promote(si); clean_add(si); }
  }
  return moves;
}

}  // namespace blance_oracle
