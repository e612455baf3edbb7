// blance_b200/csrc/host_api.cpp — see host_api.hpp.  Interning (strings -> flat
// int32 tables), the calls into the CUDA library, and the way back to maps.
#include "host_api.hpp"

#include <algorithm>
#include <chrono>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string_view>
#include <thread>
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

// string_view -> dense index, open addressing; the views point into the caller's maps, which
// outlive the call.  Built single-threaded, read from many threads.
struct SvTable {
  std::vector<uint32_t> slots;                  // index + 1, 0 = empty
  std::vector<uint64_t> hashes;
  std::vector<std::string_view> keys;
  uint64_t mask = 0;
  static uint64_t hash(std::string_view s) {    // FNV-1a with a final mix
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    h ^= h >> 32; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 29;
    return h;
  }
  void init(size_t n) {
    size_t cap = 16;
    while (cap < n * 2 + 2) cap <<= 1;
    slots.assign(cap, 0u);
    mask = cap - 1;
    hashes.reserve(n);
    keys.reserve(n);
  }
  int32_t find(std::string_view s, uint64_t h) const {
    if (slots.empty()) return -1;
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      const uint32_t e = slots[i];
      if (!e) return -1;
      if (hashes[e - 1] == h && keys[e - 1] == s) return int32_t(e - 1);
    }
  }
  int32_t insert(std::string_view s, uint64_t h) {   // the index of s, new or old
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      const uint32_t e = slots[i];
      if (!e) {
        keys.push_back(s);
        hashes.push_back(h);
        slots[i] = uint32_t(keys.size());
        return int32_t(keys.size() - 1);
      }
      if (hashes[e - 1] == h && keys[e - 1] == s) return int32_t(e - 1);
    }
  }
};

// ---- a little data parallelism for the million-partition maps (the per-partition work is independent) ----
std::atomic<int> g_host_threads{0};             // SetHostThreads; 0 = default

int max_threads() {
  const int forced = g_host_threads.load(std::memory_order_relaxed);
  if (forced >= 1) return forced;
  static const int n = [] {
    if (const char* e = std::getenv("BLANCE_HOST_THREADS")) { const int v = std::atoi(e); if (v >= 1) return std::min(v, 64); }
    const unsigned hc = std::thread::hardware_concurrency();
    return int(std::min(16u, std::max(1u, hc)));
  }();
  return n;
}

// f(begin, end, thread index); runs inline below 32 768 items
template <class F>
void parallel_for(size_t n, F f) {
  const int T = max_threads();
  if (T <= 1 || n < 32768) { if (n) f(size_t(0), n, 0); return; }
  std::vector<std::thread> th;
  const size_t chunk = (n + size_t(T) - 1) / size_t(T);
  for (int t = 1; t < T; ++t) {
    const size_t lo = std::min(n, chunk * size_t(t)), hi = std::min(n, lo + chunk);
    if (lo < hi) th.emplace_back([=] { f(lo, hi, t); });
  }
  f(size_t(0), std::min(n, chunk), 0);
  for (auto& x : th) x.join();
}

// chunk sorts in parallel, then pairwise merges
template <class T, class Less>
void parallel_sort(std::vector<T>& v, Less less) {
  const int TN = max_threads();
  if (TN <= 1 || v.size() < 65536) { std::sort(v.begin(), v.end(), less); return; }
  int parts = 1;
  while (parts * 2 <= TN) parts *= 2;
  const size_t n = v.size();
  std::vector<size_t> cut(size_t(parts) + 1);
  for (int i = 0; i <= parts; ++i) cut[size_t(i)] = n * size_t(i) / size_t(parts);
  {
    std::vector<std::thread> th;
    for (int i = 1; i < parts; ++i) th.emplace_back([&, i] { std::sort(v.begin() + long(cut[size_t(i)]), v.begin() + long(cut[size_t(i) + 1]), less); });
    std::sort(v.begin(), v.begin() + long(cut[1]), less);
    for (auto& x : th) x.join();
  }
  for (int width = 1; width < parts; width *= 2) {
    std::vector<std::thread> th;
    for (int i = 0; i + width < parts; i += 2 * width) {
      const size_t lo = cut[size_t(i)], mid = cut[size_t(i + width)], hi = cut[size_t(std::min(parts, i + 2 * width))];
      th.emplace_back([&, lo, mid, hi] { std::inplace_merge(v.begin() + long(lo), v.begin() + long(mid), v.begin() + long(hi), less); });
    }
    for (auto& x : th) x.join();
  }
}

// first error raised inside a parallel region
struct ErrorSlot {
  std::mutex mu;
  bool has = false;
  std::string msg;
  void set(const std::string& m) { std::lock_guard<std::mutex> g(mu); if (!has) { has = true; msg = m; } }
  void rethrow() { if (has) invalid(msg); }
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
  SvTable node_tab;                              // read-only view of nodesAll for the parallel passes
  node_tab.init(size_t(N));
  for (const auto& n : nodesAll) node_tab.insert(n, SvTable::hash(n));

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
  // the model has a handful of states: comparing names beats hashing them
  auto find_state = [&](const std::string& name) -> int32_t {
    for (int32_t s = 0; s < S; ++s)
      if (ip->state_names[size_t(s)].size() == name.size() && ip->state_names[size_t(s)] == name) return s;
    return -1;
  };

  // ---- partitions, indexed in the name order of the partition sort key (plan.go:519-528, 512).
  // The maps are walked once into pointer arrays; everything per partition after that runs in parallel.
  using Entry = const PartitionMap::value_type*;
  const bool same_map = &prevMap == &partitionsToAssign;
  std::vector<Entry> pv, av;
  auto walk = [&](const PartitionMap& m, std::vector<Entry>& out) {     // bucket ranges in parallel
    out.resize(m.size());
    const size_t B = m.bucket_count();
    const int T = max_threads();
    if (T <= 1 || m.size() < 32768) { size_t i = 0; for (const auto& kv : m) out[i++] = &kv; return; }
    std::vector<std::vector<Entry>> part{size_t(T)};
    parallel_for(B, [&](size_t lo, size_t hi, int t) {
      auto& mine = part[size_t(t)];
      mine.reserve((hi - lo) * m.size() / B + 16);
      for (size_t b = lo; b < hi; ++b)
        for (auto it = m.begin(b); it != m.end(b); ++it) mine.push_back(&*it);
    });
    size_t i = 0;
    for (const auto& v : part) { std::copy(v.begin(), v.end(), out.begin() + long(i)); i += v.size(); }
  };
  walk(prevMap, pv);
  if (!same_map) walk(partitionsToAssign, av);
  auto check_names = [&](const std::vector<Entry>& v) {
    ErrorSlot err;
    parallel_for(v.size(), [&](size_t lo, size_t hi, int) {
      for (size_t i = lo; i < hi; ++i)
        if (!v[i]->second.Name.empty() && v[i]->second.Name != v[i]->first) {
          err.set("Partition.Name '" + v[i]->second.Name + "' differs from its map key '" + v[i]->first + "'");
          return;
        }
    });
    err.rethrow();
  };
  check_names(pv);
  check_names(av);
  // name -> unique index (u), in first-seen order: prevMap's entries, then the new ones of partitionsToAssign
  std::vector<uint64_t> hp(pv.size()), ha(av.size());
  parallel_for(pv.size(), [&](size_t lo, size_t hi, int) { for (size_t i = lo; i < hi; ++i) hp[i] = SvTable::hash(pv[i]->first); });
  parallel_for(av.size(), [&](size_t lo, size_t hi, int) { for (size_t i = lo; i < hi; ++i) ha[i] = SvTable::hash(av[i]->first); });
  SvTable names;
  names.init(pv.size() + av.size());
  for (size_t i = 0; i < pv.size(); ++i) names.insert(pv[i]->first, hp[i]);        // map keys are unique: u == i
  std::vector<int32_t> au(av.size());
  for (size_t i = 0; i < av.size(); ++i) au[i] = names.insert(av[i]->first, ha[i]);
  const size_t U = names.keys.size();
  const int32_t PU = int32_t(U);
  // sort keys: names that are small non-negative integers compare as integers ("%10d" of v < 10^10 is ten
  // characters wide, so the padded strings order like the numbers); anything else compares as the strings do
  struct SortKey { long long v; uint32_t u; uint32_t numeric; };
  std::vector<SortKey> keys(U);
  parallel_for(U, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i) {
      long long v = 0;
      const std::string raw(names.keys[i]);
      const bool num = go_atoi(raw, &v) && v >= 0 && v < 10000000000ll;
      keys[i] = SortKey{num ? v : 0, uint32_t(i), num ? 1u : 0u};
    }
  });
  auto padded_of = [&](const SortKey& k) -> std::string {
    std::string raw(names.keys[k.u]);
    if (k.numeric) return pad10(k.v);
    long long v;
    if (go_atoi(raw, &v) && v >= 0) return pad10(v);
    return raw;
  };
  auto key_less = [&](const SortKey& a, const SortKey& b) {
    if (a.numeric && b.numeric) {
      if (a.v != b.v) return a.v < b.v;
      return names.keys[a.u] < names.keys[b.u];
    }
    const std::string pa = padded_of(a), pb = padded_of(b);
    if (pa != pb) return pa < pb;
    return names.keys[a.u] < names.keys[b.u];
  };
  parallel_sort(keys, key_less);
  std::vector<int32_t> rank(U);                 // unique index -> partition id
  ip->part_names.resize(U);
  parallel_for(U, [&](size_t lo, size_t hi, int) {
    for (size_t p = lo; p < hi; ++p) { rank[keys[p].u] = int32_t(p); ip->part_names[p] = std::string(names.keys[keys[p].u]); }
  });
  auto part_of_prev = [&](size_t i) { return rank[i]; };
  auto part_of_assign = [&](size_t i) { return rank[size_t(au[i])]; };

  // ---- per-partition scalars
  ip->part_in_prev.assign(size_t(PU), 0);
  ip->part_in_assign.assign(size_t(PU), 0);
  ip->part_weight.assign(size_t(PU), 1);
  ip->part_has_weight.assign(size_t(PU), 0);
  ip->part_name_rank.resize(size_t(PU));
  for (int32_t p = 0; p < PU; ++p) ip->part_name_rank[size_t(p)] = p;
  if (options.PartitionWeights)
    for (const auto& kv : *options.PartitionWeights) {
      const int32_t u = names.find(kv.first, SvTable::hash(kv.first));
      if (u < 0) continue;
      ip->part_weight[size_t(rank[size_t(u)])] = kv.second;
      ip->part_has_weight[size_t(rank[size_t(u)])] = 1;
    }

  // ---- slot layout and rows.  A state's range holds max(constraints, longest input list).  The rows are
  // filled in ONE pass over the maps assuming the constraints are wide enough; a longer list (rare) only
  // records the width it needs and the pass is repeated with the right layout.
  std::vector<int32_t> cap(size_t(S), 0);
  for (int32_t s = 0; s < S; ++s) cap[size_t(s)] = std::max(0, ip->state_constraints[size_t(s)]);
  struct Extra { int32_t part; int32_t node; };
  struct Unknown { size_t pos; int which; const std::string* name; };   // a row cell (or extras entry) naming a node outside nodesAll
  std::vector<Extra> extras;   // prevMap entries under non-model states (only feed tot)
  int32_t SL = 0;
  for (int attempt = 0;; ++attempt) {
    ip->state_slot_off.assign(size_t(S) + 1, 0);
    for (int32_t s = 0; s < S; ++s) ip->state_slot_off[size_t(s) + 1] = ip->state_slot_off[size_t(s)] + cap[size_t(s)];
    SL = ip->state_slot_off[size_t(S)];
    extras.clear();
    std::vector<int32_t> need = cap;
    auto fill = [&](const std::vector<Entry>& v, bool from_prev, std::vector<int32_t>& rows, std::vector<uint8_t>& shape,
                    std::vector<uint8_t>& present, bool is_prev, bool must_be_model) {
      const int T = max_threads();
      ErrorSlot err;
      std::vector<std::vector<Extra>> textra{size_t(T)};
      std::vector<std::vector<Unknown>> tunk{size_t(T)};
      std::vector<std::vector<int32_t>> tneed(size_t(T), std::vector<int32_t>(size_t(S), 0));
      parallel_for(v.size(), [&](size_t lo, size_t hi, int t) {
        for (size_t i = lo; i < hi; ++i) {
          const int32_t p = from_prev ? part_of_prev(i) : part_of_assign(i);
          present[size_t(p)] |= 1;
          for (const auto& sn : v[i]->second.NodesByState) {
            const int32_t s = find_state(sn.first);
            if (s < 0) {
              if (must_be_model) {
                err.set("partition '" + v[i]->first + "' has state '" + sn.first + "' that is not in the model (the reference panics, plan.go:148)");
                return;
              }
              if (is_prev) present[size_t(p)] = 3;        // a key outside the model: never DeepEqual (plan.go:38)
              if (is_prev)
                for (const auto& n : deref(sn.second)) {
                  const int32_t id = node_tab.find(n, SvTable::hash(n));
                  if (id < 0) tunk[size_t(t)].push_back({textra[size_t(t)].size(), -1, &n});
                  textra[size_t(t)].push_back({p, id});
                }
              continue;
            }
            shape[size_t(p) * size_t(S) + size_t(s)] = sn.second ? BLANCE_SHAPE_LIST : BLANCE_SHAPE_NIL;
            const Strs& list = deref(sn.second);
            if (int32_t(list.size()) > cap[size_t(s)]) { tneed[size_t(t)][size_t(s)] = std::max(tneed[size_t(t)][size_t(s)], int32_t(list.size())); continue; }
            size_t pos = size_t(p) * size_t(SL) + size_t(ip->state_slot_off[size_t(s)]);
            for (const auto& n : list) {
              const int32_t id = node_tab.find(n, SvTable::hash(n));
              if (id >= 0) rows[pos] = id;
              else tunk[size_t(t)].push_back({pos, 0, &n});
              ++pos;
            }
          }
        }
      });
      err.rethrow();
      // names outside nodesAll get their ids here, in a fixed order (thread, then position)
      for (int t = 0; t < T; ++t) {
        for (const auto& u : tunk[size_t(t)]) {
          const int32_t id = nodes.get(*u.name);
          if (u.which >= 0) rows[u.pos] = id;
          else textra[size_t(t)][u.pos].node = id;
        }
        extras.insert(extras.end(), textra[size_t(t)].begin(), textra[size_t(t)].end());
        for (int32_t s = 0; s < S; ++s) need[size_t(s)] = std::max(need[size_t(s)], tneed[size_t(t)][size_t(s)]);
      }
    };
    ip->prev_rows.assign(size_t(PU) * size_t(SL), BLANCE_NO_NODE);
    ip->prev_shape.assign(size_t(PU) * size_t(S), BLANCE_SHAPE_ABSENT);
    fill(pv, true, ip->prev_rows, ip->prev_shape, ip->part_in_prev, true, same_map);
    if (!same_map) {
      ip->cur_rows.assign(size_t(PU) * size_t(SL), BLANCE_NO_NODE);
      ip->cur_shape.assign(size_t(PU) * size_t(S), BLANCE_SHAPE_ABSENT);
      fill(av, false, ip->cur_rows, ip->cur_shape, ip->part_in_assign, false, true);
    }
    if (need == cap) break;
    if (attempt >= 1) invalid("internal: slot layout did not settle");
    cap = need;
  }
  if (same_map) {
    ip->cur_rows = ip->prev_rows;
    ip->cur_shape = ip->prev_shape;
    ip->part_in_assign = ip->part_in_prev;
  }

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
  // the assigned partitions (plan.go:326-330), built in parallel, then moved into the map
  std::vector<int32_t> ids;
  ids.reserve(size_t(in.n_parts));
  for (int32_t p = 0; p < in.n_parts; ++p)
    if (ip.part_in_assign[size_t(p)]) ids.push_back(p);
  std::vector<Partition> parts(ids.size());
  parallel_for(ids.size(), [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i) {
      const int32_t p = ids[i];
      Partition& part = parts[i];
      part.Name = ip.part_names[size_t(p)];
      const int32_t* row = ob.next_rows.data() + size_t(p) * size_t(in.n_slots);
      for (int32_t s = 0; s < in.n_states; ++s) {
        const uint8_t sh = ob.next_shape[size_t(p) * size_t(in.n_states) + size_t(s)];
        if (sh == BLANCE_SHAPE_ABSENT) continue;
        if (sh == BLANCE_SHAPE_NIL) { part.NodesByState[ip.state_names[size_t(s)]] = std::nullopt; continue; }
        Strs list;
        for (int32_t j = ip.state_slot_off[size_t(s)]; j < ip.state_slot_off[size_t(s) + 1] && row[j] != BLANCE_NO_NODE; ++j)
          list.push_back(ip.node_names[size_t(row[j])]);
        part.NodesByState[ip.state_names[size_t(s)]] = std::move(list);
      }
    }
  });
  PartitionMap next;
  next.reserve(ids.size());
  for (size_t i = 0; i < ids.size(); ++i) {
    const int32_t p = ids[i];
    if (warnings)
      for (int32_t s = 0; s < in.n_states; ++s)
        if (ob.warn[size_t(p) * size_t(in.n_states) + size_t(s)]) {
          char buf[32];                                              // plan.go:231-234
          std::snprintf(buf, sizeof buf, "%d", ip.state_constraints[size_t(s)]);
          (*warnings)[parts[i].Name].push_back(std::string("could not meet constraints: ") + buf +
                                               ", stateName: " + ip.state_names[size_t(s)] +
                                               ", partitionName: " + parts[i].Name);
        }
    std::string key = parts[i].Name;
    next.emplace(std::move(key), std::move(parts[i]));
  }
  return next;
}

// ---- the JSON wire form (api.go:30,35) --------------------------------------------------------------------
// encoding/json, default options: strings are quoted with ", \\ and control characters escaped (\n \r \t short
// forms, others \u00XX), <, > and & as \u003c \u003e \u0026 (HTML-safe), U+2028 / U+2029 as \u2028 / \u2029,
// invalid UTF-8 as U+FFFD; map keys are sorted by their bytes; a nil slice is null.
static void json_string(std::string& out, const std::string& v) {
  static const char* hex = "0123456789abcdef";
  out.push_back('"');
  const unsigned char* p = reinterpret_cast<const unsigned char*>(v.data());
  const size_t n = v.size();
  for (size_t i = 0; i < n;) {
    const unsigned char c = p[i];
    if (c < 0x80) {
      if (c == '"' || c == '\\') { out.push_back('\\'); out.push_back(char(c)); }
      else if (c == '\n') out += "\\n";
      else if (c == '\r') out += "\\r";
      else if (c == '\t') out += "\\t";
      else if (c < 0x20 || c == '<' || c == '>' || c == '&') { out += "\\u00"; out.push_back(hex[c >> 4]); out.push_back(hex[c & 15]); }
      else out.push_back(char(c));
      ++i;
      continue;
    }
    // decode one UTF-8 sequence the way Go's utf8.DecodeRuneInString does (shortest form, no surrogates, <= U+10FFFF)
    size_t len = 0;
    uint32_t cp = 0;
    if (c >= 0xC2 && c <= 0xDF) { len = 2; cp = c & 0x1F; }
    else if (c >= 0xE0 && c <= 0xEF) { len = 3; cp = c & 0x0F; }
    else if (c >= 0xF0 && c <= 0xF4) { len = 4; cp = c & 0x07; }
    bool ok = len != 0 && i + len <= n;
    for (size_t k = 1; ok && k < len; ++k) {
      const unsigned char d = p[i + k];
      unsigned char lo = 0x80, hi = 0xBF;
      if (k == 1) {
        if (c == 0xE0) lo = 0xA0;
        if (c == 0xED) hi = 0x9F;
        if (c == 0xF0) lo = 0x90;
        if (c == 0xF4) hi = 0x8F;
      }
      if (d < lo || d > hi) ok = false;
      cp = (cp << 6) | (d & 0x3F);
    }
    if (!ok) { out += "\\ufffd"; ++i; continue; }
    if (cp == 0x2028 || cp == 0x2029) { out += "\\u202"; out.push_back(hex[cp & 15]); }
    else out.append(v, i, len);
    i += len;
  }
  out.push_back('"');
}

template <class GetList>   // GetList(state index) -> (present, is_nil, list writer)
static void json_partition(std::string& out, const std::string& name, const std::vector<std::pair<const std::string*, const OptStrs*>>& states) {
  out += "{\"name\":";
  json_string(out, name);
  out += ",\"nodesByState\":{";
  bool first = true;
  for (const auto& st : states) {
    if (!first) out.push_back(',');
    first = false;
    json_string(out, *st.first);
    out.push_back(':');
    if (!*st.second) { out += "null"; continue; }
    out.push_back('[');
    bool f2 = true;
    for (const auto& n : **st.second) {
      if (!f2) out.push_back(',');
      f2 = false;
      json_string(out, n);
    }
    out.push_back(']');
  }
  out += "}}";
}

std::string PartitionMapToJSON(const PartitionMap& m) {
  std::vector<const std::pair<const std::string, Partition>*> entries;
  entries.reserve(m.size());
  for (const auto& kv : m) entries.push_back(&kv);
  std::sort(entries.begin(), entries.end(), [](auto* a, auto* b) { return a->first < b->first; });   // byte order = Go's key order
  std::vector<std::string> chunks(entries.size());
  parallel_for(entries.size(), [&](size_t lo, size_t hi, int) {
    std::vector<std::pair<const std::string*, const OptStrs*>> states;
    for (size_t i = lo; i < hi; ++i) {
      const Partition& part = entries[i]->second;
      states.clear();
      for (const auto& kv : part.NodesByState) states.push_back({&kv.first, &kv.second});
      std::sort(states.begin(), states.end(), [](const auto& a, const auto& b) { return *a.first < *b.first; });
      std::string& out = chunks[i];
      json_string(out, entries[i]->first);
      out.push_back(':');
      json_partition<int>(out, part.Name, states);
    }
  });
  size_t total = 2;
  for (const auto& c : chunks) total += c.size() + 1;
  std::string out;
  out.reserve(total);
  out.push_back('{');
  for (size_t i = 0; i < chunks.size(); ++i) {
    if (i) out.push_back(',');
    out += chunks[i];
  }
  out.push_back('}');
  return out;
}

// rows -> JSON directly (the next map of a plan; same bytes as PartitionMapToJSON(UninternPlan(...)))
std::string PlanResultToJSON(const InternedPlan& ip, const PlanOutBuffers& ob) {
  const blance_plan_in& in = ip.in;
  std::vector<int32_t> ids;
  for (int32_t p = 0; p < in.n_parts; ++p)
    if (ip.part_in_assign[size_t(p)]) ids.push_back(p);
  std::sort(ids.begin(), ids.end(), [&](int32_t a, int32_t b) { return ip.part_names[size_t(a)] < ip.part_names[size_t(b)]; });
  std::vector<int32_t> state_order(size_t(in.n_states));
  for (int32_t s = 0; s < in.n_states; ++s) state_order[size_t(s)] = s;
  std::sort(state_order.begin(), state_order.end(), [&](int32_t a, int32_t b) { return ip.state_names[size_t(a)] < ip.state_names[size_t(b)]; });
  std::vector<std::string> chunks(ids.size());
  parallel_for(ids.size(), [&](size_t lo, size_t hi, int) {
    std::vector<OptStrs> lists(size_t(in.n_states));
    std::vector<std::pair<const std::string*, const OptStrs*>> states;
    for (size_t i = lo; i < hi; ++i) {
      const int32_t p = ids[i];
      const int32_t* row = ob.next_rows.data() + size_t(p) * size_t(in.n_slots);
      states.clear();
      for (int32_t s : state_order) {
        const uint8_t sh = ob.next_shape[size_t(p) * size_t(in.n_states) + size_t(s)];
        if (sh == BLANCE_SHAPE_ABSENT) continue;
        OptStrs& l = lists[size_t(s)];
        if (sh == BLANCE_SHAPE_NIL) l = std::nullopt;
        else {
          l = Strs{};
          for (int32_t j = ip.state_slot_off[size_t(s)]; j < ip.state_slot_off[size_t(s) + 1] && row[j] != BLANCE_NO_NODE; ++j)
            l->push_back(ip.node_names[size_t(row[j])]);
        }
        states.push_back({&ip.state_names[size_t(s)], &l});
      }
      std::string& out = chunks[i];
      json_string(out, ip.part_names[size_t(p)]);
      out.push_back(':');
      json_partition<int>(out, ip.part_names[size_t(p)], states);
    }
  });
  std::string out;
  size_t total = 2;
  for (const auto& c : chunks) total += c.size() + 1;
  out.reserve(total);
  out.push_back('{');
  for (size_t i = 0; i < chunks.size(); ++i) {
    if (i) out.push_back(',');
    out += chunks[i];
  }
  out.push_back('}');
  return out;
}

// plan.go:49-52: after any non-matching iteration the caller's maps hold the new partitions; when the loop
// ends their content equals the returned map.  Map surgery (new keys) is serial, the deep copies are not.
void ReplayCallerMutation(const PartitionMap& next, PartitionMap& prevMap, PartitionMap& partitionsToAssign) {
  const bool same = &prevMap == &partitionsToAssign;
  std::vector<const Partition*> src;
  std::vector<Partition*> dst_prev, dst_assign;
  src.reserve(next.size());
  dst_prev.reserve(next.size());
  if (!same) dst_assign.reserve(next.size());
  for (const auto& kv : next) {
    src.push_back(&kv.second);
    dst_prev.push_back(&prevMap[kv.first]);
    if (!same) dst_assign.push_back(&partitionsToAssign[kv.first]);
  }
  parallel_for(src.size(), [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; ++i) {
      *dst_prev[i] = *src[i];
      if (!same) *dst_assign[i] = *src[i];
    }
  });
}

void SetHostThreads(int n) { g_host_threads.store(n < 0 ? 0 : (n > 64 ? 64 : n), std::memory_order_relaxed); }
int HostThreads() { return max_threads(); }

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
  using clk = std::chrono::steady_clock;
  auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  const auto t0 = clk::now();
  auto ip = InternPlan(prevMap, partitionsToAssign, nodesAll, nodesToRemove, nodesToAdd, model, options);
  PlanOutBuffers ob(*ip);
  const auto t1 = clk::now();
  blance_ctx* ctx = DefaultContext();
  int st = blance_plan_next_map(ctx, &ip->in, &ob.out);
  if (st != BLANCE_OK) throw BlanceError(st, std::string("blance_plan_next_map failed: ") + blance_last_error(ctx));
  const auto t2 = clk::now();
  if (stats) {
    stats->iters_run = ob.out.iters_run; stats->converged = ob.out.converged; stats->steps = ob.out.steps;
    stats->device_ms = ob.out.device_ms; stats->kernel_ms = ob.out.kernel_ms; stats->pass_ms = ob.out.pass_ms;
    stats->intern_ms = ms(t0, t1); stats->call_ms = ms(t1, t2);
  }
  if (ob.out.iters_run <= 0) return PartitionMap{};                  // MaxIterationsPerPlan <= 0: plan.go:32,57
  PartitionMap next = UninternPlan(*ip, ob, warnings);
  const auto t3 = clk::now();
  if (ob.out.iters_run >= 2 || !ob.out.converged) ReplayCallerMutation(next, prevMap, partitionsToAssign);
  if (stats) { stats->unintern_ms = ms(t2, t3); stats->mutate_ms = ms(t3, clk::now()); }
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
