// oracle/literal_py.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// pybind11 face of the literal oracle (oracle/literal.cpp) so that tests/ and
// bench.py's cpu_baseline leg can drive it with plain dicts/lists.  A
// PartitionMap crosses as {partitionName: {stateName: [node, ...] | None}}
// (Partition.Name == its key, as in every reference fixture).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <chrono>

#include "literal.hpp"

namespace py = pybind11;
using namespace blance_oracle;

using PyPartitionMap = std::unordered_map<std::string, NodesByState>;

static PartitionMap to_map(const PyPartitionMap& m) {
  PartitionMap out;
  for (const auto& kv : m) {
    auto p = std::make_shared<Partition>();
    p->name = kv.first;
    p->nodes_by_state = kv.second;
    out[kv.first] = std::move(p);
  }
  return out;
}

static PyPartitionMap from_map(const PartitionMap& m) {
  PyPartitionMap out;
  for (const auto& kv : m) out[kv.first] = kv.second->nodes_by_state;
  return out;
}

static Options make_options(const std::optional<IntMap>& msc, const std::optional<IntMap>& pw,
                            const std::optional<IntMap>& ss, const std::optional<IntMap>& nw,
                            const std::optional<StrMap>& nh,
                            const std::optional<std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>>>& hr,
                            int booster, int max_iterations, int64_t max_steps_per_pass) {
  Options o;
  o.model_state_constraints = msc;
  o.partition_weights = pw;
  o.state_stickiness = ss;
  o.node_weights = nw;
  o.node_hierarchy = nh;
  if (hr) {
    HierarchyRules rules;
    for (const auto& kv : *hr) {
      auto& v = rules[kv.first];
      for (const auto& r : kv.second) v.push_back({r.first, r.second});
    }
    o.hierarchy_rules = std::move(rules);
  }
  o.booster = booster == 1 ? Booster::CbgtMax : Booster::None;
  o.max_iterations = max_iterations;
  o.max_steps_per_pass = max_steps_per_pass;
  return o;
}

PYBIND11_MODULE(_literal, m) {
  m.doc() = "literal CPU oracle of couchbase/blance's planner (test infrastructure)";

  m.def(
      "plan_next_map_ex",
      [](const PyPartitionMap& prev, const std::optional<PyPartitionMap>& assign, const Strs& nodes_all,
         const OptStrs& nodes_to_remove, const OptStrs& nodes_to_add,
         const std::unordered_map<std::string, std::pair<int64_t, int64_t>>& model,
         const std::optional<IntMap>& msc, const std::optional<IntMap>& pw, const std::optional<IntMap>& ss,
         const std::optional<IntMap>& nw, const std::optional<StrMap>& nh,
         const std::optional<std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>>>& hr,
         int booster, int max_iterations, int64_t max_steps_per_pass, int64_t slice_steps, bool memoize_partition_scores) {
        PartitionModel pm;
        for (const auto& kv : model) pm[kv.first] = {kv.second.first, kv.second.second};
        Options o = make_options(msc, pw, ss, nw, nh, hr, booster, max_iterations, max_steps_per_pass);
        std::vector<double> slices;
        o.slice_steps = slice_steps;
        o.memoize_partition_scores = memoize_partition_scores;
        if (slice_steps > 0) o.slice_seconds = &slices;
        PartitionMap prev_map = to_map(prev);
        PartitionMap assign_map;
        // None = the caller passed the SAME map object twice (plan_test.go:1716-1718)
        PartitionMap* assign_ptr = &prev_map;
        if (assign) { assign_map = to_map(*assign); assign_ptr = &assign_map; }
        PlanResult r;
        double secs;
        {
          py::gil_scoped_release rel;
          auto t0 = std::chrono::steady_clock::now();
          r = plan_next_map_ex(prev_map, *assign_ptr, nodes_all, nodes_to_remove, nodes_to_add, pm, o);
          secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        py::dict out;
        out["next_map"] = from_map(r.next_map);
        out["warnings"] = r.warnings;
        out["prev_map"] = from_map(prev_map);
        out["partitions_to_assign"] = from_map(*assign_ptr);
        out["iterations"] = r.iterations;
        out["steps"] = r.steps;
        out["seconds"] = secs;
        out["slice_seconds"] = slices;
        return out;
      },
      py::arg("prev_map"), py::arg("partitions_to_assign"), py::arg("nodes_all"), py::arg("nodes_to_remove"),
      py::arg("nodes_to_add"), py::arg("model"), py::arg("model_state_constraints") = py::none(),
      py::arg("partition_weights") = py::none(), py::arg("state_stickiness") = py::none(),
      py::arg("node_weights") = py::none(), py::arg("node_hierarchy") = py::none(),
      py::arg("hierarchy_rules") = py::none(), py::arg("booster") = 0, py::arg("max_iterations") = 10,
      py::arg("max_steps_per_pass") = -1, py::arg("slice_steps") = 0,
      py::arg("memoize_partition_scores") = false);

  m.def("calc_partition_moves",
        [](const Strs& states, const NodesByState& beg, const NodesByState& end, bool favor_min_nodes) {
          std::vector<std::tuple<std::string, std::string, std::string>> out;
          for (const auto& op : calc_partition_moves(states, beg, end, favor_min_nodes))
            out.emplace_back(op.node, op.state, op.op);
          return out;
        });
  m.def("find_state_changes", &find_state_changes);
  m.def("strings_remove_strings", &strings_remove_strings);
  m.def("strings_intersect_strings", &strings_intersect_strings);
  m.def("strings_deduplicate", &strings_deduplicate);
  m.def("flatten_nodes_by_state", &flatten_nodes_by_state);
  m.def("remove_nodes_from_nodes_by_state", &remove_nodes_from_nodes_by_state);
  m.def("state_name_sort",
        [](const std::optional<std::unordered_map<std::string, std::pair<int64_t, int64_t>>>& model, Strs names) {
          PartitionModel pm;
          if (model)
            for (const auto& kv : *model) pm[kv.first] = {kv.second.first, kv.second.second};
          state_name_insertion_sort(model ? &pm : nullptr, names);
          return names;
        });
  m.def("count_state_nodes", [](const PyPartitionMap& pmap, const std::optional<IntMap>& w) {
    return count_state_nodes(to_map(pmap), w);
  });
  m.def("map_parents_to_map_children", &map_parents_to_map_children);
  m.def("find_ancestor", &find_ancestor);
  m.def("find_leaves", &find_leaves);
}
