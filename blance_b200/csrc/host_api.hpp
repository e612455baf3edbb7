// blance_b200/csrc/host_api.hpp — host-side mirror of blance's public planner API.
//
// The reference is Go and no Go toolchain exists in this image, so the host side
// above the C ABI (include/blance_b200.h) is written in C++ and mirrors api.go
// name for name:
//
//   PartitionMap / Partition            api.go:24-36
//   PartitionModel / ...State           api.go:41-62
//   HierarchyRules / HierarchyRule      api.go:75-105
//   PlanNextMapOptions                  api.go:183-190
//   PlanNextMap / PlanNextMapEx         api.go:109-157   (body -> blance_plan_next_map)
//   NodeStateOp / CalcPartitionMoves    moves.go:17-46   (body -> blance_calc_partition_moves)
//
// It does what the cgo shim of INTEGRATION.md does in Go: intern strings into the
// flat int32 tables of the C ABI, call the CUDA library, rebuild the maps, format
// the warning strings (plan.go:232-234) and replay the caller-map mutation of
// plan.go:49-52.  There is no CPU fallback: the calls throw BlanceError when the
// CUDA library reports a failure (e.g. no device).
#pragma once

#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "blance_b200.h"

namespace blance {

using Strs = std::vector<std::string>;
using OptStrs = std::optional<Strs>;   // Go's nil slice = nullopt (plan.go:554, reflect.DeepEqual)
using NodesByState = std::unordered_map<std::string, OptStrs>;

struct Partition {                     // api.go:28-36
  std::string Name;
  ::blance::NodesByState NodesByState;
};
using PartitionMap = std::unordered_map<std::string, Partition>;   // api.go:24 (values, not pointers)

// The JSON wire form of a PartitionMap, byte for byte what Go's encoding/json produces for
// map[string]*Partition with the tags of api.go:30,35 (`json:"name"`, `json:"nodesByState"`): object keys sorted
// by their bytes, a nil slice as null, strings escaped with encoding/json's default (HTML-safe) rules.
// The per-partition objects are rendered in parallel for large maps.
std::string PartitionMapToJSON(const PartitionMap& m);
// The same for the flat result of a plan (rows -> JSON without building the PartitionMap first).
struct InternedPlan;
struct PlanOutBuffers;
std::string PlanResultToJSON(const InternedPlan& ip, const PlanOutBuffers& ob);

struct PartitionModelState { int Priority = 0; int Constraints = 0; };   // api.go:46-62
using PartitionModel = std::unordered_map<std::string, PartitionModelState>;

struct HierarchyRule { int IncludeLevel = 0; int ExcludeLevel = 0; };     // api.go:95-105
using HierarchyRules = std::unordered_map<std::string, std::vector<HierarchyRule>>;

struct PlanNextMapOptions {            // api.go:183-190
  std::optional<std::unordered_map<std::string, int>> ModelStateConstraints;
  std::optional<std::unordered_map<std::string, int>> PartitionWeights;
  std::optional<std::unordered_map<std::string, int>> StateStickiness;
  std::optional<std::unordered_map<std::string, int>> NodeWeights;
  std::optional<std::unordered_map<std::string, std::string>> NodeHierarchy;
  std::optional<::blance::HierarchyRules> HierarchyRules;
  // The package-level knobs of plan.go, which a C ABI cannot read from Go globals:
  int MaxIterationsPerPlan = 10;       // plan.go:21
  int NodeScoreBooster = BLANCE_BOOSTER_NONE;   // plan.go:693; enum blance_booster
  int Engine = BLANCE_ENGINE_AUTO;     // enum blance_engine (not in the reference; results do not depend on it)
};

using Warnings = std::unordered_map<std::string, Strs>;

struct BlanceError : std::runtime_error {
  int status;
  BlanceError(int st, const std::string& what) : std::runtime_error(what), status(st) {}
};

struct PlanStats {                     // not in the reference; what the GPU did
  int iters_run = 0;
  int converged = 0;
  int64_t steps = 0;
  float device_ms = 0, kernel_ms = 0, pass_ms = 0;
  // host wall time of the stages of PlanNextMapEx: maps -> tables, the C ABI call, tables -> map, plan.go:49-52
  double intern_ms = 0, call_ms = 0, unintern_ms = 0, mutate_ms = 0;
};

// api.go:147-157.  prevMap and partitionsToAssign are mutated as plan.go:49-52
// mutates them (they may be the same object).
PartitionMap PlanNextMapEx(PartitionMap& prevMap, PartitionMap& partitionsToAssign, const Strs& nodesAll,
                           const OptStrs& nodesToRemove, const OptStrs& nodesToAdd,
                           const PartitionModel& model, const PlanNextMapOptions& options,
                           Warnings* warnings, PlanStats* stats = nullptr);

struct NodeStateOp { std::string Node, State, Op; };   // moves.go:17-21

// moves.go:41-46 for one partition (a batch of one on the device).
std::vector<NodeStateOp> CalcPartitionMoves(const Strs& states, const NodesByState& begNodesByState,
                                            const NodesByState& endNodesByState, bool favorMinNodes);

// The vectorised form: every partition of `beg` U `end` in one launch.
std::unordered_map<std::string, std::vector<NodeStateOp>> CalcPartitionMovesMap(
    const Strs& states, const PartitionMap& beg, const PartitionMap& end, bool favorMinNodes);

// ---------------------------------------------------------------------------------
// The interning layer, exposed so that tests can drive the SAME tables through the
// CPU oracle and compare array for array.

struct InternedPlan {
  // name tables
  Strs node_names;        // [n_node_ids]
  Strs state_names;       // [n_states] in sortStateNames order
  Strs part_names;        // [n_parts] in the name-rule order of plan.go:519-528,512
  // tables (owning storage for the pointers in `in`)
  std::vector<int32_t> state_priority, state_constraints, state_slot_off, state_stickiness;
  std::vector<uint8_t> state_has_stickiness;
  std::vector<uint8_t> node_removed, node_added, node_has_weight;
  std::vector<int32_t> node_weight;
  std::vector<uint8_t> part_in_prev, part_in_assign, part_has_weight;
  std::vector<int32_t> part_weight, part_name_rank;
  std::vector<int32_t> prev_rows, cur_rows;
  std::vector<uint8_t> prev_shape, cur_shape;
  std::vector<int32_t> extra_tot_first, extra_tot_rest;
  std::vector<int32_t> rule_off;
  std::vector<uint32_t> ie_mask;
  blance_plan_in in{};    // points into the vectors above
};

std::unique_ptr<InternedPlan> InternPlan(const PartitionMap& prevMap, const PartitionMap& partitionsToAssign,
                                         const Strs& nodesAll, const OptStrs& nodesToRemove,
                                         const OptStrs& nodesToAdd, const PartitionModel& model,
                                         const PlanNextMapOptions& options);

struct PlanOutBuffers {
  std::vector<int32_t> next_rows;
  std::vector<uint8_t> next_shape, warn;
  blance_plan_out out{};
  explicit PlanOutBuffers(const InternedPlan& ip);
};

// rows -> PartitionMap of the assigned partitions, plus the warning strings.
PartitionMap UninternPlan(const InternedPlan& ip, const PlanOutBuffers& ob, Warnings* warnings);

// plan.go:49-52: store the partitions of `next` into the caller's maps (they may be the same object).
void ReplayCallerMutation(const PartitionMap& next, PartitionMap& prevMap, PartitionMap& partitionsToAssign);

// The marshalling layer (InternPlan, UninternPlan, ReplayCallerMutation) splits maps of 32 768 partitions
// or more over host threads: min(16, hardware threads) by default, BLANCE_HOST_THREADS in the environment,
// or this call (0 = back to the default).  Results do not depend on the thread count.
void SetHostThreads(int n);
int HostThreads();

// The process-wide context the host API runs on (created on first use).
blance_ctx* DefaultContext();

}  // namespace blance
