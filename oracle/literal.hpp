// oracle/literal.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// "Literal" CPU oracle: a statement-for-statement C++ restatement of the
// reference planner (couchbase/blance, Go) using the SAME data structures the
// reference uses — string-keyed hash maps, slices of strings, a comparison
// sort whose comparator re-evaluates the score on every call — so that it has
// the reference's asymptotics and can stand in as the "reference-equivalent CPU
// path" (there is no Go toolchain in this image or on the GPU box, so the Go
// code itself can never be executed here).
//
// Parity status: PINNED by the reference's own golden vectors — 69 active
// PlanNextMap cases, 9 findStateChanges + 29 CalcPartitionMoves cases and the
// helper unit tables, transcribed by tests/golden/make_fixtures.py and checked
// in tests/test_oracle_golden.py.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may include, link or execute anything in this directory.
//
// Each function cites the reference file:line it follows (paths relative to the
// reference tree).
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <unordered_map>
#include <vector>

namespace blance_oracle {

using Strs = std::vector<std::string>;
// Go distinguishes a nil slice from an empty one (reflect.DeepEqual does, and
// plan.go:554 tests nodesToAdd != nil), so nil-able slices are optionals.
using OptStrs = std::optional<Strs>;
using NodesByState = std::unordered_map<std::string, OptStrs>;

struct Partition {                       // api.go:28-36
  std::string name;
  NodesByState nodes_by_state;
};
using PartitionPtr = std::shared_ptr<Partition>;
using PartitionMap = std::unordered_map<std::string, PartitionPtr>;   // api.go:24

struct ModelState { int64_t priority = 0; int64_t constraints = 0; };  // api.go:46-62
using PartitionModel = std::unordered_map<std::string, ModelState>;    // api.go:41

struct HierarchyRule { int64_t include_level = 0; int64_t exclude_level = 0; };  // api.go:95-105
using HierarchyRules = std::unordered_map<std::string, std::vector<HierarchyRule>>;  // api.go:75

using IntMap = std::unordered_map<std::string, int64_t>;
using StrMap = std::unordered_map<std::string, std::string>;

enum class Booster { None = 0, CbgtMax = 1 };   // plan.go:693-697; control_test.go:19-26

struct Options {                         // api.go:183-190 (+ the package-level hooks)
  std::optional<IntMap> model_state_constraints;
  std::optional<IntMap> partition_weights;
  std::optional<IntMap> state_stickiness;
  std::optional<IntMap> node_weights;
  std::optional<StrMap> node_hierarchy;
  std::optional<HierarchyRules> hierarchy_rules;
  Booster booster = Booster::None;       // NodeScoreBooster, plan.go:693
  int max_iterations = 10;               // MaxIterationsPerPlan, plan.go:21
  // Optional cap on how many partitions each state pass processes (bench.py's
  // bounded cpu_baseline sample; <0 = no cap).  Not part of the reference.
  int64_t max_steps_per_pass = -1;
  // Optional timing of the greedy chain in slices of `slice_steps` consecutive findBestNodes steps (bench.py's
  // reference arm: every slice is one bounded sample of the workload; the copies and sorts around the passes are
  // outside the slices).  slice_seconds receives one entry per completed slice.  Not part of the reference.
  int64_t slice_steps = 0;
  // Evaluate partitionSorter.Score once per partition and sort the precomputed keys instead of recomputing the
  // two Sprintf strings inside every comparison (plan.go:495-562).  Same order (Score is a function of the
  // partition alone during a sort); only bench.py's reference arm sets it, to keep its untimed set-up short.
  bool memoize_partition_scores = false;
  std::vector<double>* slice_seconds = nullptr;
};

using Warnings = std::unordered_map<std::string, Strs>;

struct PlanResult {
  PartitionMap next_map;
  Warnings warnings;
  int iterations = 0;          // inner plans executed
  int64_t steps = 0;           // findBestNodes calls executed (all iterations)
};

// misc.go
Strs strings_remove_strings(const Strs& a, const OptStrs& remove);      // misc.go:27-36
Strs strings_intersect_strings(const Strs& a, const OptStrs& b);        // misc.go:40-51
Strs strings_deduplicate(const Strs& a);                                // misc.go:55-66

// plan.go helpers (exposed for the unit tables)
Strs flatten_nodes_by_state(const NodesByState& nbs, const Strs& state_order);          // plan.go:425-431
NodesByState remove_nodes_from_nodes_by_state(const NodesByState& nbs, const OptStrs& remove);  // plan.go:408-421
void state_name_insertion_sort(const PartitionModel* model, Strs& names);               // plan.go:450-474
Strs sort_state_names(const PartitionModel& model);                                     // plan.go:437-447
std::unordered_map<std::string, IntMap> count_state_nodes(const PartitionMap& m,
                                                          const std::optional<IntMap>& weights);  // plan.go:374-399
std::unordered_map<std::string, Strs> map_parents_to_map_children(const StrMap& parents);   // plan.go:703-717
std::string find_ancestor(std::string node, const StrMap& parents, int64_t level);      // plan.go:755-762
Strs find_leaves(const std::string& node, const std::unordered_map<std::string, Strs>& children);  // plan.go:764-774

// plan.go:23-58.  prev_map and partitions_to_assign are mutated exactly as the
// reference mutates its arguments (plan.go:49-52).  Pass the same object for
// both to reproduce the "Vis" harness (plan_test.go:1716-1718).
PlanResult plan_next_map_ex(PartitionMap& prev_map, PartitionMap& partitions_to_assign,
                            Strs nodes_all, OptStrs nodes_to_remove, OptStrs nodes_to_add,
                            const PartitionModel& model, const Options& opts);

struct NodeStateOp { std::string node, state, op; };   // moves.go:17-21
Strs find_state_changes(int beg_state_idx, int end_state_idx, const std::string& state,
                        const Strs& states, const NodesByState& beg, const NodesByState& end);  // moves.go:121-136
std::vector<NodeStateOp> calc_partition_moves(const Strs& states, const NodesByState& beg,
                                              const NodesByState& end, bool favor_min_nodes);   // moves.go:41-119

}  // namespace blance_oracle
