// Plan-time scheduling of the resident kernel's shared-memory gathers (host only).
//
// The resident kernel stages vectors node-major (buf[node * TP + t], TP/4 odd) and every gather
// is a 128-bit load; the 8 lanes of a quarter-warp are served in one wavefront only if their
// target nodes fall into distinct 16-byte bank groups, i.e. are distinct mod 8 (or identical).
// Nothing in the algorithm fixes (a) how nodes are numbered inside the kernel, (b) in which
// order a row's neighbours are visited, or (c) in which order a node's in-list is visited —
// so all three are chosen here:
//   (0) self links leave the tables (the owning thread has the value in registers);
//   (a) warps: nodes sorted by in-degree (ties in reverse Cuthill-McKee order), so the 32 rows of
//       a warp have in-lists of similar length — the per-warp ELL pads to the longest;
//       inside a warp, a local search permutes the nodes (= picks each node's bank class, id mod
//       8, and its quarter-warp) until, for every quarter-warp and table, no class is the target
//       of more gathers than the table has slots — the condition under which (b)/(c) can be
//       conflict-free;
//   (b)/(c) per quarter-warp, the (row, target) pairs are the edges of a bipartite multigraph
//       rows x bank classes; a proper edge colouring with `slots` colours (König: exists when no
//       vertex has more edges than colours; alternating-path construction) is the visit order:
//       one colour = one slot in which the 8 rows hit 8 different bank groups.
// Measured on the PEMS04-shaped graph (307 nodes, k = 6): in-list steps per CTA 139 -> 65,
// wavefronts per quarter-warp gather phase see tests/test_cabi.py.  Only the ORDER of
// floating-point additions inside a row sum changes (the reference's own torch reductions do
// not pin one either).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <functional>
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

// node order for the streaming path: reverse Cuthill-McKee on the symmetrised union of both tables
std::vector<int> graph_rcm_order(int N, int kd, const int* nbr_d, int ku, const int* nbr_u) {
  std::vector<std::vector<int>> adj(N);
  auto link = [&](int a, int b) { if (a != b && a >= 0 && b >= 0) { adj[a].push_back(b); adj[b].push_back(a); } };
  for (int i = 0; i < N; ++i) {
    for (int j = 0; j < kd; ++j) link(i, nbr_d[(size_t)i * kd + j]);
    for (int j = 0; j < ku; ++j) link(i, nbr_u[(size_t)i * ku + j]);
  }
  for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
  return rcm_order(N, adj);
}

namespace {
struct Cand {
  int node;     // node id (original numbering before placement, internal after), or N for the zero row
  float w;
};

// ---- (a) bank-class balancing ---------------------------------------------------------------
// load[t][q][c] = number of gathers of table t issued by the rows of quarter-warp q whose target
// has bank class c; a slot assignment without conflicts needs load <= slots(t, q).
struct Balancer {
  int N, NT, nq;
  const std::vector<std::vector<Cand>>* tab[3];
  std::vector<std::vector<int>> refs[3];          // refs[t][v] = rows that gather v in table t
  std::vector<int> slots[3];                      // per quarter
  std::vector<int> load[3];                       // nq * 8
  std::vector<int> pos;                           // pos[node]
  long total = 0;

  // cost of one (table, quarter): the wavefronts a perfect visit order still needs beyond one per slot
  // are max_c load - slots (two classes one over can share their extra wavefront), weighted heavily;
  // the sum over classes is kept as a small term that gives the search a slope on plateaus
  long cost(int t, int q) const {
    int mx = 0, sum = 0;
    for (int c = 0; c < 8; ++c) {
      const int o = std::max(0, load[t][q * 8 + c] - slots[t][q]);
      mx = std::max(mx, o);
      sum += o;
    }
    return 16L * mx + sum;
  }
  void bump(int t, int q, int c, int d) {
    total -= cost(t, q);
    load[t][q * 8 + c] += d;
    total += cost(t, q);
  }
  void row(int v, int d) {                        // add / remove all gathers issued by row v
    const int q = pos[v] >> 3;
    for (int t = 0; t < 3; ++t)
      for (const Cand& c : (*tab[t])[v]) bump(t, q, pos[c.node] & 7, d);
  }
  void reclass(int v, int from, int to, int skip_a, int skip_b) {   // v's class changes for every row that gathers it
    for (int t = 0; t < 3; ++t)
      for (int r : refs[t][v]) {
        if (r == skip_a || r == skip_b) continue;
        const int q = pos[r] >> 3;
        bump(t, q, from, -1);
        bump(t, q, to, +1);
      }
  }
  void swap(int u, int v) {
    row(u, -1); row(v, -1);
    const int cu = pos[u] & 7, cv = pos[v] & 7;
    if (cu != cv) { reclass(u, cu, cv, u, v); reclass(v, cv, cu, u, v); }
    std::swap(pos[u], pos[v]);
    row(u, +1); row(v, +1);
  }
};

// ---- (b)/(c) visit order of one quarter-warp = edge colouring of rows x bank classes -----------
// rows[r] = the gathers of row r (<= S of them); on return rows[r] has exactly S entries, entry j
// being the gather of slot j (a zero row with zero weight where the row has none).
//
// Slots are peeled off one at a time.  With m slots left, a class that is still the target of L
// gathers must place at least L - (m - 1) of them in this slot, or a later slot cannot be
// conflict-free; these demands, one gather per row, are a bipartite matching (rows -> classes with
// capacities), found by augmenting paths.  A class whose demand is 2 makes this slot cost two
// wavefronts — the surplus of ALL over-subscribed classes is taken by the same (first) slot, which
// is what meets the bound max(S, max_c L_c) wavefronts per quarter-warp.
static void colour_quarter(int N, int S, std::vector<std::vector<Cand>>& rows) {
  const int R = (int)rows.size();
  std::vector<std::vector<Cand>> rem = rows;
  for (int r = 0; r < R; ++r) rows[r].assign(S, Cand{N, 0.f});
  for (int m = S; m >= 1; --m) {
    const int slot = S - m;
    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0}, need[8];
    for (int r = 0; r < R; ++r) for (const Cand& c : rem[r]) load[c.node & 7]++;
    for (int c = 0; c < 8; ++c) need[c] = std::max(0, load[c] - (m - 1));
    std::vector<int> pick(R, -1);                       // pick[r] = class row r gives to this slot
    auto has = [&](int r, int c) { for (const Cand& x : rem[r]) if ((x.node & 7) == c) return true; return false; };
    // phase 1 — the demands: Kuhn's algorithm, demand units on the left, rows on the right
    std::vector<int> unit_cls, match_row(R, -1);
    for (int c = 0; c < 8; ++c) for (int k = 0; k < need[c]; ++k) unit_cls.push_back(c);
    std::vector<char> vis;
    std::function<bool(int)> try_unit = [&](int u) {
      for (int r = 0; r < R; ++r) {
        if (vis[r] || !has(r, unit_cls[u])) continue;
        vis[r] = 1;
        if (match_row[r] < 0 || try_unit(match_row[r])) { match_row[r] = u; return true; }
      }
      return false;
    };
    for (int u = 0; u < (int)unit_cls.size(); ++u) { vis.assign(R, 0); try_unit(u); }
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cap[8];
    for (int r = 0; r < R; ++r) if (match_row[r] >= 0) { pick[r] = unit_cls[match_row[r]]; cnt[pick[r]]++; }
    for (int c = 0; c < 8; ++c) cap[c] = std::max(1, need[c]);
    // phase 2 — rows with exactly m gathers left must use this slot as well: augmenting paths from the row
    // (row -> class with room, or class -> one of its rows -> that row's other class ...); counts of the
    // classes on the way do not change, so the demands stay met
    std::vector<char> visc;
    std::function<bool(int)> try_row = [&](int r) {
      for (const Cand& x : rem[r]) {
        const int c = x.node & 7;
        if (visc[c]) continue;
        visc[c] = 1;
        if (cnt[c] < cap[c]) { pick[r] = c; cnt[c]++; return true; }
        for (int r2 = 0; r2 < R; ++r2)
          if (r2 != r && pick[r2] == c && try_row(r2)) { pick[r] = c; return true; }
      }
      return false;
    };
    for (int r = 0; r < R; ++r) {
      if (pick[r] >= 0 || (int)rem[r].size() < m || rem[r].empty()) continue;
      visc.assign(8, 0);
      if (try_row(r)) continue;
      int best = -1;                                     // no room anywhere: the class least used in this slot
      for (const Cand& x : rem[r]) {
        const int cc = x.node & 7;
        if (best < 0 || cnt[cc] < cnt[best] || (cnt[cc] == cnt[best] && load[cc] > load[best])) best = cc;
      }
      pick[r] = best;
      cnt[best]++;
    }
    for (int r = 0; r < R; ++r) {
      if (pick[r] < 0) continue;
      for (size_t k = 0; k < rem[r].size(); ++k)
        if ((rem[r][k].node & 7) == pick[r]) {
          rows[r][slot] = rem[r][k];
          rem[r].erase(rem[r].begin() + k);
          break;
        }
    }
  }
  // polish: exchange two slots of one row while that lowers the wavefront count (where a class is the
  // target of more gathers than there are slots, this stacks the surplus of several classes into the
  // same slot instead of spoiling one slot each)
  auto wavefronts = [&](int s) {
    int seen[8][8], cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < R; ++r) {
      const int n = rows[r][s].node;
      if (n >= N) continue;
      const int c = n & 7;
      bool dup = false;
      for (int k = 0; k < cnt[c]; ++k) dup |= (seen[c][k] == n);
      if (!dup) seen[c][cnt[c]++] = n;
    }
    int m = 1;
    for (int c = 0; c < 8; ++c) m = std::max(m, cnt[c]);
    return m;
  };
  // strictly improving exchanges first; then sideways ones as well (a fixed pseudo-random subset, so the
  // walk does not cycle), keeping the best arrangement seen
  auto total_wf = [&]() { int t = 0; for (int s = 0; s < S; ++s) t += wavefronts(s); return t; };
  std::vector<std::vector<Cand>> best_rows = rows;
  int best_total = total_wf();
  uint32_t lcg = 12345u;
  for (int sweep = 0; sweep < 48 && best_total > S; ++sweep) {
    const bool sideways = sweep >= 4;
    bool moved = false;
    for (int r = 0; r < R; ++r)
      for (int s1 = 0; s1 < S; ++s1)
        for (int s2 = s1 + 1; s2 < S; ++s2) {
          if (rows[r][s1].node == rows[r][s2].node) continue;
          const int before = wavefronts(s1) + wavefronts(s2);
          if (before == 2) continue;
          std::swap(rows[r][s1], rows[r][s2]);
          const int after = wavefronts(s1) + wavefronts(s2);
          lcg = lcg * 1664525u + 1013904223u;
          if (after < before || (sideways && after == before && (lcg >> 28) < 5)) moved = true;
          else std::swap(rows[r][s1], rows[r][s2]);
        }
    const int t = total_wf();
    if (t < best_total) { best_total = t; best_rows = rows; }
    if (!moved) break;
  }
  rows = best_rows;
  // padding entries read a zero row: rows N .. N+7 are all zero, one per bank class, and the pads of a
  // slot take the one whose class no real target of that slot uses (all pads of a slot share it: broadcast)
  for (int s = 0; s < S; ++s) {
    bool used[8] = {false, false, false, false, false, false, false, false};
    for (int r = 0; r < R; ++r) if (rows[r][s].node < N) used[rows[r][s].node & 7] = true;
    int z = 0;
    for (int k = 0; k < 8; ++k) if (!used[(N + k) & 7]) { z = k; break; }
    for (int r = 0; r < R; ++r) if (rows[r][s].node >= N) rows[r][s].node = N + z;
  }
}
}  // namespace

void build_resident_schedule(int N, int kd, const int* nbr_d, const float* d_w, int ku, const int* nbr_u,
                             const float* u_w, const int* csr_ptr, const int* csr_src, const float* csr_w,
                             ResidentSchedule* out) {
  ResidentSchedule& S = *out;
  S.N = N;
  // ---- (0) self links leave the tables: the thread that owns node i holds p_i and q_i in registers,
  // so the self term of L_d (and of L_d^T: the in-list entry (i -> i) carries the same weight) is one
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
  // ---- (a) warps by in-degree
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
  const int NT = ((N + 31) / 32) * 32, n_warps = NT / 32, nq = NT / 8;
  S.ell_ptr.assign(n_warps + 1, 0);
  for (int w = 0; w < n_warps; ++w) {
    int m = 0;
    for (int p = w * 32; p < std::min(N, w * 32 + 32); ++p) m = std::max(m, (int)in_l[S.perm[p]].size());
    S.ell_ptr[w + 1] = S.ell_ptr[w] + m;
  }
  // ---- (a) continued: bank classes and quarter-warps inside each warp
  {
    Balancer B;
    B.N = N; B.NT = NT; B.nq = nq;
    B.tab[0] = &fwd_d; B.tab[1] = &fwd_u; B.tab[2] = &in_l;
    B.pos.assign(N, 0);
    for (int p = 0; p < N; ++p) B.pos[S.perm[p]] = p;
    for (int t = 0; t < 3; ++t) {
      B.refs[t].assign(N, {});
      for (int v = 0; v < N; ++v)
        for (const Cand& c : (*B.tab[t])[v]) B.refs[t][c.node].push_back(v);
      B.slots[t].assign(nq, 0);
      B.load[t].assign((size_t)nq * 8, 0);
    }
    for (int q = 0; q < nq; ++q) {
      B.slots[0][q] = S.kd; B.slots[1][q] = S.ku;
      B.slots[2][q] = S.ell_ptr[q / 4 + 1] - S.ell_ptr[q / 4];
    }
    for (int v = 0; v < N; ++v) B.row(v, +1);
    const long start_total = B.total;
    uint64_t rng = 0x9E3779B97F4A7C15ull;                        // fixed seed: the schedule is deterministic
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    std::vector<int> at(NT, -1);                                  // at[pos] = node
    for (int v = 0; v < N; ++v) at[B.pos[v]] = v;
    long max_trials = 40L * N;      // measured at PEMS04 shape: 496k / 507k / 522k / 524k windows/s with 1k / 3k / 10k / 29k trials
    if (const char* e = std::getenv("MGA_SCHED_TRIALS")) max_trials = std::atol(e);
    double t_start = 4.0;
    if (const char* e = std::getenv("MGA_SCHED_TEMP")) t_start = std::atof(e);
    long best_total = B.total;
    std::vector<int> best_at = at;
    for (long trial = 0; trial < max_trials && B.total > 0; ++trial) {
      const int w = (int)(next() % n_warps);
      const int pa = w * 32 + (int)(next() % 32), pb = w * 32 + (int)(next() % 32);
      if (pa == pb || at[pa] < 0 || at[pb] < 0) continue;
      const int u = at[pa], v = at[pb];
      const long before = B.total;
      B.swap(u, v);
      // annealing: uphill moves are accepted with a probability that falls as the search proceeds
      const double temp = t_start * (1.0 - (double)trial / (double)max_trials);
      const long delta = B.total - before;
      bool accept = delta <= 0;
      if (!accept && temp > 0.0) accept = (double)(next() >> 11) * (1.0 / 9007199254740992.0) < std::exp(-(double)delta / temp);
      if (accept) {
        std::swap(at[pa], at[pb]);
        if (B.total < best_total) { best_total = B.total; best_at = at; }
      } else {
        B.swap(u, v);
      }
    }
    if (best_total < B.total) {                                   // rebuild the loads of the best placement seen
      at = best_at;
      for (int p = 0; p < NT; ++p) if (at[p] >= 0) B.pos[at[p]] = p;
      for (int t = 0; t < 3; ++t) std::fill(B.load[t].begin(), B.load[t].end(), 0);
      B.total = 0;
      for (int v = 0; v < N; ++v) B.row(v, +1);
    }
    for (int p = 0; p < NT; ++p) if (at[p] >= 0) S.perm[p] = at[p];
    S.overload = 0;
    for (int t = 0; t < 3; ++t) for (int q = 0; q < nq; ++q) S.overload += (int)(B.cost(t, q) / 16);
    if (std::getenv("MGA_SCHED_VERBOSE"))
      for (int t = 0; t < 3; ++t) {
        long bound = 0, base = 0;
        int hist[6] = {0, 0, 0, 0, 0, 0};
        for (int q = 0; q < nq; ++q) {
          int mx = 0;
          for (int c = 0; c < 8; ++c) mx = std::max(mx, B.load[t][q * 8 + c]);
          bound += std::max(B.slots[t][q], mx); base += B.slots[t][q];
          hist[std::min(5, std::max(0, mx - B.slots[t][q]))]++;
        }
        std::fprintf(stderr, "[mga]   table %d: slots %ld, wavefront bound %ld (%.3f), extra-per-quarter hist %d %d %d %d %d %d\n", t, base,
                     bound, base ? (double)bound / base : 0.0, hist[0], hist[1], hist[2], hist[3], hist[4], hist[5]);
      }
    if (std::getenv("MGA_SCHED_VERBOSE"))
      std::fprintf(stderr, "[mga] schedule: N %d, slots d %d u %d, in-list steps %d, overload %ld -> %ld\n", N, S.kd,
                   S.ku, S.ell_ptr[n_warps], start_total, B.total);
  }
  S.inv.assign(N, 0);
  for (int p = 0; p < N; ++p) S.inv[S.perm[p]] = p;
  {
    std::vector<float> ws(N);
    for (int p = 0; p < N; ++p) ws[p] = S.w_self[S.perm[p]];
    S.w_self.swap(ws);                                           // internal order from here on
  }
  // ---- (b) forward tables: one edge colouring per quarter-warp
  auto forward = [&](int K, const std::vector<std::vector<Cand>>& src, std::vector<int>& o_n, std::vector<float>& o_w) {
    o_n.assign((size_t)N * K, N);          // (pads: one of the zero rows N .. N+7)
    o_w.assign((size_t)N * K, 0.f);
    if (K == 0) return;
    for (int q0 = 0; q0 < N; q0 += 8) {
      const int q1 = std::min(N, q0 + 8);
      std::vector<std::vector<Cand>> rows(q1 - q0);
      for (int p = q0; p < q1; ++p)
        for (const Cand& c : src[S.perm[p]]) rows[p - q0].push_back(Cand{S.inv[c.node], c.w});
      colour_quarter(N, K, rows);
      for (int p = q0; p < q1; ++p)
        for (int j = 0; j < K; ++j) { o_n[(size_t)p * K + j] = rows[p - q0][j].node; o_w[(size_t)p * K + j] = rows[p - q0][j].w; }
    }
  };
  forward(S.kd, fwd_d, S.nbr_d, S.w_d);
  forward(S.ku, fwd_u, S.nbr_u, S.w_u);
  // ---- (c) in-list as per-warp ELL, step-major, one edge colouring per quarter-warp
  S.ell_node.assign((size_t)S.ell_ptr[n_warps] * 32, N);
  S.ell_w.assign((size_t)S.ell_ptr[n_warps] * 32, 0.f);
  for (int w = 0; w < n_warps; ++w) {
    const int steps = S.ell_ptr[w + 1] - S.ell_ptr[w];
    if (steps == 0) continue;
    for (int q0 = 0; q0 < 32; q0 += 8) {
      std::vector<std::vector<Cand>> rows(8);
      for (int r = 0; r < 8; ++r) {
        const int p = w * 32 + q0 + r;
        if (p < N) for (const Cand& c : in_l[S.perm[p]]) rows[r].push_back(Cand{S.inv[c.node], c.w});
      }
      colour_quarter(N, steps, rows);
      for (int r = 0; r < 8; ++r)
        for (int e = 0; e < steps; ++e) {
          const size_t at = (size_t)(S.ell_ptr[w] + e) * 32 + q0 + r;
          S.ell_node[at] = rows[r][e].node;
          S.ell_w[at] = rows[r][e].w;
        }
    }
  }
}

}  // namespace mga
