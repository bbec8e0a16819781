// kNN tables by shortest-path distance: the step before the hot path (utils.py:183-204).
//
// The reference runs a FULL networkx Dijkstra from every node and keeps the k+1 closest
// (O(N * E log N); ~27 min at N = 20 000).  Here the search stops after k+1 nodes are settled.
// Bit-identical tables need networkx's exact visiting order:
//   * heap entries are (distance, push counter, node) — ties in distance pop in push order;
//   * successors are relaxed in adjacency insertion order; nx.DiGraph.add_edge on a repeated
//     (u, v) keeps the original position and overwrites the weight (utils.py:190-192);
//   * a node is pushed when unseen or strictly closer than its best tentative distance;
//   * distances accumulate in float64 and are stored to a float32 table (utils.py:196, 203);
//   * heapq.nsmallest(k+1, ..., key=dist) is stable and the settle order is already sorted.
#include <cmath>
#include <cstdint>
#include <limits>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>

#include "mga.h"

namespace mga { void set_error(const std::string& msg); }

namespace {
struct Entry {
  double d;
  int64_t cnt;
  int node;
};
struct Later {
  bool operator()(const Entry& a, const Entry& b) const { return a.d > b.d || (a.d == b.d && a.cnt > b.cnt); }
};
}  // namespace

extern "C" int mga_knn_build(int32_t n_nodes, int64_t n_edges, const int64_t* edges, const double* dists, int32_t k,
                             int32_t* out_nodes, float* out_dists) {
  if (n_nodes <= 0 || n_edges < 0 || k < 0 || !edges || !dists || !out_nodes || !out_dists) {
    mga::set_error("mga_knn_build: bad argument");
    return MGA_ERR_INVALID;
  }
  std::vector<std::vector<std::pair<int, double>>> succ(n_nodes);
  std::vector<char> present(n_nodes, 0);
  std::unordered_map<int64_t, int> where;   // (u, v) -> position in succ[u]
  where.reserve((size_t)n_edges * 2);
  for (int64_t e = 0; e < n_edges; ++e) {
    const int64_t u = edges[2 * e], v = edges[2 * e + 1];
    if (u < 0 || u >= n_nodes || v < 0 || v >= n_nodes) {
      mga::set_error("mga_knn_build: edge endpoint outside [0, n_nodes)");
      return MGA_ERR_INDEX;
    }
    present[u] = present[v] = 1;
    const int64_t key = u * (int64_t)n_nodes + v;
    auto it = where.find(key);
    if (it == where.end()) {
      where.emplace(key, (int)succ[u].size());
      succ[u].emplace_back((int)v, dists[e]);
    } else {
      succ[u][it->second].second = dists[e];
    }
  }
  const int K1 = k + 1;
  const float inf = std::numeric_limits<float>::infinity();
  std::vector<double> seen(n_nodes);
  std::vector<int> seen_tag(n_nodes, -1), done_tag(n_nodes, -1);
  for (int s = 0; s < n_nodes; ++s) {
    if (!present[s]) {
      mga::set_error("Node " + std::to_string(s) + " not found in graph");
      return MGA_ERR_INDEX;
    }
    int32_t* row_n = out_nodes + (size_t)s * K1;
    float* row_d = out_dists + (size_t)s * K1;
    for (int j = 0; j < K1; ++j) { row_n[j] = -1; row_d[j] = inf; }
    std::priority_queue<Entry, std::vector<Entry>, Later> heap;
    int64_t cnt = 0;
    heap.push({0.0, cnt++, s});
    seen[s] = 0.0;
    seen_tag[s] = s;
    int settled = 0;
    while (!heap.empty() && settled < K1) {
      const Entry top = heap.top();
      heap.pop();
      if (done_tag[top.node] == s) continue;
      done_tag[top.node] = s;
      row_n[settled] = top.node;
      row_d[settled] = (float)top.d;
      ++settled;
      for (const auto& nb : succ[top.node]) {
        const int u = nb.first;
        if (done_tag[u] == s) continue;
        const double nd = top.d + nb.second;
        if (seen_tag[u] != s || nd < seen[u]) {
          seen[u] = nd;
          seen_tag[u] = s;
          heap.push({nd, cnt++, u});
        }
      }
    }
  }
  return MGA_OK;
}
