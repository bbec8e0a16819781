// Plan-time scheduling of the resident kernel's shared-memory gathers (host only).
//
// The resident kernel stages vectors node-major (buf[node * TP + t], TP/4 odd) and every gather
// is a 128-bit load; the 8 lanes of a quarter-warp are served in one wavefront only if their
// target nodes fall into distinct 16-byte bank groups, i.e. are distinct mod 8 (or identical).
// Nothing in the algorithm fixes (a) how nodes are numbered inside the kernel, (b) in which
// order a row's K neighbours are visited, or (c) in which order a node's in-list is visited —
// so all three are chosen here to keep quarter-warps conflict-free:
//   (a) reverse Cuthill-McKee on the symmetrised kNN graph: neighbours get nearby numbers, so the
//       8 targets of 8 consecutive rows sit in a narrow band of node ids;
//   (b) per 8-row group, a greedy assignment of each row's neighbours to the K visit slots that
//       avoids two different targets with the same id mod 8 in one slot;
//   (c) the in-list becomes a per-warp ELL (step-major, one (row offset, weight) pair per lane and
//       step, read conflict-free), padded with zero-weight entries on the zero row to the warp's
//       maximum in-degree, with the same greedy placement per 8-row group.
// Measured on the PEMS04-shaped graph: wavefronts per quarter-phase 2.29 -> 1.22 (forward tables)
// and 1.78 -> 1.11 (in-list).  Only the ORDER of floating-point additions inside a row sum
// changes (the reference's own torch reductions do not pin one either).
#include <algorithm>
#include <numeric>
#include <queue>
#include <set>
#include <vector>

#include "mga_schedule.h"

namespace mga {

static std::vector<int> rcm_order(int N, const std::vector<std::vector<int>>& adj) {
  std::vector<int> deg(N), order;
  for (int i = 0; i < N; ++i) deg[i] = (int)adj[i].size();
  std::vector<char> seen(N, 0);
  std::vector<int> by_deg(N);
  std::iota(by_deg.begin(), by_deg.end(), 0);
  std::stable_sort(by_deg.begin(), by_deg.end(), [&](int a, int b) { return deg[a] < deg[b]; });
  order.reserve(N);
  for (int start : by_deg) {
    if (seen[start]) continue;
    std::queue<int> q;
    q.push(start);
    seen[start] = 1;
    while (!q.empty()) {
      const int v = q.front();
      q.pop();
      order.push_back(v);
      std::vector<int> nb;
      for (int u : adj[v]) if (!seen[u]) { seen[u] = 1; nb.push_back(u); }
      std::stable_sort(nb.begin(), nb.end(), [&](int a, int b) { return deg[a] < deg[b]; });
      for (int u : nb) q.push(u);
    }
  }
  std::reverse(order.begin(), order.end());
  return order;   // order[new] = old
}

namespace {
struct Cand {
  int node;     // internal id, or N for the zero row
  float w;
};
// number of DIFFERENT nodes already using this node's bank group in the current slot
inline int clash(const std::vector<std::set<int>>& used, int node, int N) {
  if (node >= N) return 0;                 // zero row: every lane reads the same address (broadcast)
  const std::set<int>& s = used[node & 7];
  return (int)s.size() - (int)s.count(node);
}
inline void take(std::vector<std::set<int>>& used, int node, int N) {
  if (node < N) used[node & 7].insert(node);
}
}  // namespace

// rows: per internal row its K candidates; writes the slot order in place
static void assign_slots(int N, int K, std::vector<std::vector<Cand>>& rows) {
  const int R = (int)rows.size();
  for (int q0 = 0; q0 < R; q0 += 8) {
    const int q1 = std::min(R, q0 + 8);
    std::vector<std::vector<Cand>> rem(rows.begin() + q0, rows.begin() + q1);
    for (int j = 0; j < K; ++j) {
      std::vector<std::set<int>> used(8);
      std::vector<int> idx(q1 - q0);
      std::iota(idx.begin(), idx.end(), 0);
      auto distinct = [&](int r) { std::set<int> s; for (auto& c : rem[r]) s.insert(c.node); return (int)s.size(); };
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return distinct(a) < distinct(b); });
      for (int r : idx) {
        int best = 0, bc = 1 << 30;
        for (int c = 0; c < (int)rem[r].size(); ++c) {
          const int cl = clash(used, rem[r][c].node, N);
          if (cl < bc) { bc = cl; best = c; }
        }
        rows[q0 + r][j] = rem[r][best];
        take(used, rem[r][best].node, N);
        rem[r].erase(rem[r].begin() + best);
      }
    }
  }
}

void build_resident_schedule(int N, int kd, const int* nbr_d, const float* d_w, int ku, const int* nbr_u,
                             const float* u_w, const int* csr_ptr, const int* csr_src, const float* csr_w,
                             ResidentSchedule* out) {
  ResidentSchedule& S = *out;
  S.N = N;
  // ---- self links leave the tables: the thread that owns node i holds p_i and q_i in registers, so
  // the self term of L_d (and of L_d^T: the in-list entry (i -> i) carries the same weight) is one
  // multiply instead of a shared-memory gather.  kNN tables always list the node itself first
  // (utils.py:199-203), i.e. one of K forward gathers and one in-list entry per node go away.
  S.w_self.assign(N, 0.f);
  std::vector<std::vector<Cand>> fwd_d(N), fwd_u(N), in_l(N);     // original numbering
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < kd; ++j) {
      const int nb = nbr_d[(size_t)i * kd + j];
      if (nb < 0) continue;                                      // "-1 = no neighbour" contributes 0 (quirk Q6)
      if (nb == i) S.w_self[i] += d_w[(size_t)i * kd + j];
      else fwd_d[i].push_back(Cand{nb, d_w[(size_t)i * kd + j]});
    }
    for (int j = 0; j < ku; ++j) {
      const int nb = nbr_u[(size_t)i * ku + j];
      if (nb >= 0) fwd_u[i].push_back(Cand{nb, u_w[(size_t)i * ku + j]});
    }
    for (int e = csr_ptr[i]; e < csr_ptr[i + 1]; ++e)
      if (csr_src[e] != i) in_l[i].push_back(Cand{csr_src[e], csr_w[e]});
  }
  S.kd = 0; S.ku = 0;
  for (int i = 0; i < N; ++i) {
    S.kd = std::max(S.kd, (int)fwd_d[i].size());
    S.ku = std::max(S.ku, (int)fwd_u[i].size());
  }
  // ---- (a) node order: reverse Cuthill-McKee, then stably by in-degree (largest first) so that
  // the 32 rows of a warp have in-lists of similar length (the per-warp ELL pads to the longest)
  std::vector<std::vector<int>> adj(N);
  auto link = [&](int a, int b) { if (a != b && a >= 0 && b >= 0) { adj[a].push_back(b); adj[b].push_back(a); } };
  for (int i = 0; i < N; ++i) {
    for (auto& c : fwd_d[i]) link(i, c.node);
    for (auto& c : fwd_u[i]) link(i, c.node);
    for (auto& c : in_l[i]) link(i, c.node);
  }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  S.perm = rcm_order(N, adj);
  std::stable_sort(S.perm.begin(), S.perm.end(), [&](int a, int b) { return in_l[a].size() > in_l[b].size(); });
  S.inv.assign(N, 0);
  for (int p = 0; p < N; ++p) S.inv[S.perm[p]] = p;
  {
    std::vector<float> ws(N);
    for (int p = 0; p < N; ++p) ws[p] = S.w_self[S.perm[p]];
    S.w_self.swap(ws);                                           // internal order from here on
  }
  // ---- (b) forward tables, padded to the longest row with zero-weight entries on the zero row
  auto forward = [&](int K, const std::vector<std::vector<Cand>>& src, std::vector<int>& o_n, std::vector<float>& o_w) {
    std::vector<std::vector<Cand>> rows(N, std::vector<Cand>(K, Cand{N, 0.f}));
    for (int p = 0; p < N; ++p) {
      const std::vector<Cand>& r = src[S.perm[p]];
      for (int j = 0; j < (int)r.size(); ++j) rows[p][j] = Cand{S.inv[r[j].node], r[j].w};
    }
    if (K > 0) assign_slots(N, K, rows);
    o_n.resize((size_t)N * K);
    o_w.resize((size_t)N * K);
    for (int p = 0; p < N; ++p)
      for (int j = 0; j < K; ++j) { o_n[(size_t)p * K + j] = rows[p][j].node; o_w[(size_t)p * K + j] = rows[p][j].w; }
  };
  forward(S.kd, fwd_d, S.nbr_d, S.w_d);
  forward(S.ku, fwd_u, S.nbr_u, S.w_u);
  // ---- (c) in-list as per-warp ELL
  const int NT = ((N + 31) / 32) * 32, n_warps = NT / 32;
  S.ell_ptr.assign(n_warps + 1, 0);
  std::vector<std::vector<Cand>> lists(NT);
  for (int p = 0; p < N; ++p)
    for (auto& c : in_l[S.perm[p]]) lists[p].push_back(Cand{S.inv[c.node], c.w});
  for (int w = 0; w < n_warps; ++w) {
    int m = 0;
    for (int l = 0; l < 32; ++l) m = std::max(m, (int)lists[w * 32 + l].size());
    S.ell_ptr[w + 1] = S.ell_ptr[w] + m;
  }
  S.ell_node.assign((size_t)S.ell_ptr[n_warps] * 32, N);
  S.ell_w.assign((size_t)S.ell_ptr[n_warps] * 32, 0.f);
  for (int w = 0; w < n_warps; ++w) {
    const int steps = S.ell_ptr[w + 1] - S.ell_ptr[w];
    for (int q0 = 0; q0 < 32; q0 += 8) {
      std::vector<std::vector<Cand>> rem(8);
      for (int r = 0; r < 8; ++r) rem[r] = lists[w * 32 + q0 + r];
      for (int e = 0; e < steps; ++e) {
        std::vector<std::set<int>> used(8);
        std::vector<int> idx(8);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return rem[a].size() > rem[b].size(); });
        for (int pass = 0; pass < 2; ++pass) {     // pass 0: rows that must place now; pass 1: free riders
          for (int r : idx) {
            if (rem[r].empty()) continue;
            const bool forced = (int)rem[r].size() >= steps - e;
            if ((pass == 0) != forced) continue;
            size_t at = (size_t)(S.ell_ptr[w] + e) * 32 + q0 + r;
            if (S.ell_node[at] != N || S.ell_w[at] != 0.f) continue;
            int best = -1, bc = 1 << 30;
            for (int c = 0; c < (int)rem[r].size(); ++c) {
              const int cl = clash(used, rem[r][c].node, N);
              if (cl < bc) { bc = cl; best = c; }
            }
            if (!forced && bc > 0) continue;     // wait for a conflict-free step
            S.ell_node[at] = rem[r][best].node;
            S.ell_w[at] = rem[r][best].w;
            take(used, rem[r][best].node, N);
            rem[r].erase(rem[r].begin() + best);
          }
        }
      }
    }
  }
}

}  // namespace mga
