// tools/bench_marshal.cpp — host-side cost of the marshalling layer: PartitionMap ->
// flat tables (InternPlan) and rows -> PartitionMap (UninternPlan) on a cfg-4-shaped cluster.  Pure host
// code: no GPU is touched (the plan itself is not run; the "result" rows are the input rows).
//
//   g++ -O2 -std=c++17 -pthread -Iinclude -Iblance_b200/csrc tools/bench_marshal.cpp blance_b200/csrc/host_api.cpp \
//       -Lblance_b200/lib -lblance_b200 -Wl,-rpath,$PWD/blance_b200/lib -o /tmp/bench_marshal && /tmp/bench_marshal 1048576 1024
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "host_api.hpp"

using namespace blance;
using clk = std::chrono::steady_clock;

static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int main(int argc, char** argv) {
  const int P = argc > 1 ? std::atoi(argv[1]) : 1048576;
  const int N = argc > 2 ? std::atoi(argv[2]) : 1024;
  Strs nodes;
  char buf[32];
  for (int i = 0; i < N; ++i) { std::snprintf(buf, sizeof buf, "n%04d", i); nodes.push_back(buf); }
  unsigned long long x = 0xB1A9CE04ull;
  auto rnd = [&]() { x += 0x9E3779B97F4A7C15ull; unsigned long long z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                     z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
  auto t0 = clk::now();
  PartitionMap prev;
  prev.reserve(P);
  for (int p = 0; p < P; ++p) {
    Partition part;
    part.Name = std::to_string(p);
    const int a = p % (N - 16), b = (a + 1 + int(rnd() % 7)) % (N - 16), c = (b + 1 + int(rnd() % 7)) % (N - 16);
    part.NodesByState["primary"] = Strs{nodes[a]};
    part.NodesByState["replica"] = Strs{nodes[b], nodes[c == a ? (c + 1) % (N - 16) : c]};
    prev.emplace(part.Name, std::move(part));
  }
  PartitionMap assign = prev;
  PartitionModel model{{"primary", {0, 1}}, {"replica", {1, 2}}};
  PlanNextMapOptions opt;
  opt.PartitionWeights.emplace();
  for (int p = 0; p < P; p += 4) (*opt.PartitionWeights)[std::to_string(p)] = 1 + int(rnd() % 8);
  opt.StateStickiness.emplace(std::unordered_map<std::string, int>{{"primary", 3}, {"replica", 2}});
  opt.NodeWeights.emplace();
  for (int i = 0; i < N; ++i) (*opt.NodeWeights)[nodes[i]] = 1 + int(rnd() % 4);
  OptStrs rm = Strs(nodes.begin(), nodes.begin() + 16), add = Strs(nodes.end() - 16, nodes.end());
  auto t1 = clk::now();
  std::printf("build maps: %.3f s (%d partitions x %d nodes)\n", secs(t0, t1), P, N);
  for (int rep = 0; rep < 3; ++rep) {
    auto a0 = clk::now();
    auto ip = InternPlan(prev, assign, nodes, rm, add, model, opt);
    auto a1 = clk::now();
    PlanOutBuffers ob(*ip);
    std::memcpy(ob.next_rows.data(), ip->cur_rows.data(), ob.next_rows.size() * sizeof(int32_t));
    std::memcpy(ob.next_shape.data(), ip->cur_shape.data(), ob.next_shape.size());
    auto a2 = clk::now();
    Warnings w;
    PartitionMap next = UninternPlan(*ip, ob, &w);
    auto a3 = clk::now();
    ReplayCallerMutation(next, prev, assign);
    auto a4 = clk::now();
    const double tot = secs(a0, a1) + secs(a2, a4);
    std::printf("InternPlan %.3f s | UninternPlan %.3f s | caller-map mutation %.3f s | total %.3f s (%.0f partitions/s through the string maps, %d threads), %zu partitions back\n",
                secs(a0, a1), secs(a2, a3), secs(a3, a4), tot, P / tot, HostThreads(), next.size());
  }
  return 0;
}
