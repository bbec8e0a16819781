// Plan construction (host) and the C-ABI dispatch layer of libmga.
//
// The plan replaces the graph state ADMM_algorithm.__init__ leaves behind (ADMM.py:15-98):
// it validates the neighbour indices once (the reference re-validates on every Ldr_T call,
// ADMM.py:204-206), narrows them to int32, detects time-invariant weight tables (the
// reference stores T identical copies, utils.py:294-295), and builds the in-list (transposed
// CSR) that turns the reference's scatter_add into an atomics-free gather.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mga_common.cuh"
#include "mga_schedule.h"

namespace mga {

static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { t_error = msg; }

int cuda_fail(cudaError_t e, const char* what) {
  t_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
  return MGA_ERR_CUDA;
}

int ensure_workspace(mga_plan* plan, Workspace& ws, size_t bytes) {
  if (ws.bytes >= bytes) return MGA_OK;
  if (ws.base) {
    MGA_CUDA(cudaDeviceSynchronize());
    MGA_CUDA(cudaFree(ws.base));
    ws.base = nullptr;
    ws.bytes = 0;
  }
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&ws.base, want);
  if (e != cudaSuccess) {
    want = bytes;
    e = cudaMalloc(&ws.base, want);
  }
  if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(workspace)");
  ws.bytes = want;
  (void)plan;
  return MGA_OK;
}

template <typename T>
static int upload(mga_plan* p, const std::vector<T>& h, const T** out) {
  void* d = nullptr;
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  MGA_CUDA(cudaMalloc(&d, bytes));
  p->owned.push_back(d);
  if (!h.empty()) MGA_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = static_cast<const T*>(d);
  return MGA_OK;
}

// A weight table stored as nT identical slices collapses to one (quirk Q8 / utils.py:294-295).
static int collapse_time(const float* w, int nT, size_t slice, std::vector<float>& out) {
  bool same = true;
  for (int t = 1; t < nT && same; ++t) same = std::memcmp(w, w + (size_t)t * slice, slice * sizeof(float)) == 0;
  int keep = same ? 1 : nT;
  out.assign(w, w + (size_t)keep * slice);
  return keep;
}

}  // namespace mga

using namespace mga;

// Everything of plan construction that needs no device: index validation, narrowing, collapse of
// time-expanded weights, the in-list.  `slot` receives, per in-list entry, its index into d_w.
static int validate_desc(const mga_graph_desc* d) {
  const int N = d->n_nodes, T = d->T;
  if (N <= 0 || T < 2 || d->t_in < 1 || d->t_in > T || d->ku < 0 || !d->nbr_u && d->ku > 0) {
    set_error("mga_plan_create: bad shape (need N>0, T>=2, 1<=t_in<=T)");
    return MGA_ERR_INVALID;
  }
  if (d->temporal < MGA_TEMPORAL_GRAPH || d->temporal > MGA_TEMPORAL_BAND) {
    set_error("mga_plan_create: bad temporal kind");
    return MGA_ERR_INVALID;
  }
  if (d->temporal == MGA_TEMPORAL_GRAPH && (d->kd <= 0 || !d->nbr_d || !d->d_w)) {
    set_error("mga_plan_create: temporal graph needs nbr_d / d_w");
    return MGA_ERR_INVALID;
  }
  if (d->ku > 0 && (!d->u_w || (d->u_w_T != 1 && d->u_w_T != T))) {
    set_error("mga_plan_create: u_w must be (N,ku) or (T,N,ku)");
    return MGA_ERR_INVALID;
  }
  return MGA_OK;
}

static int host_tables(const mga_graph_desc* d, mga_plan* p, std::vector<int>& slot) {
  const int N = d->n_nodes, T = d->T;
  GraphDev& g = p->g;
  g.N = N; g.T = T; g.t_in = d->t_in; g.ku = d->ku;
  g.temporal = d->temporal; g.skip = 0;

  // ---- spatial table
  p->h_nbr_u.resize((size_t)N * d->ku);
  for (size_t k = 0; k < p->h_nbr_u.size(); ++k) {
    int64_t v = d->nbr_u[k];
    if (v < -1 || v >= N) { set_error("Index out of bounds"); return (MGA_ERR_INDEX); }
    p->h_nbr_u[k] = (int)v;
  }
  g.u_wT = d->ku > 0 ? collapse_time(d->u_w, d->u_w_T, (size_t)N * d->ku, p->h_u_w) : 1;

  // ---- temporal table + in-list
  if (d->temporal == MGA_TEMPORAL_GRAPH) {
    if (d->d_w_T != 1 && d->d_w_T != T - 1) {
      set_error("mga_plan_create: d_w must be (N,kd) or (T-1,N,kd)");
      return (MGA_ERR_INVALID);
    }
    g.kd = d->kd;
    g.q1 = 1;  // ADMM.py:220-222
    p->h_nbr_d.resize((size_t)N * g.kd);
    for (size_t k = 0; k < p->h_nbr_d.size(); ++k) {
      int64_t v = d->nbr_d[k];
      if (v < -1 || v >= N) { set_error("Index out of bounds"); return (MGA_ERR_INDEX); }
      p->h_nbr_d[k] = (int)v;
    }
    g.d_wT = collapse_time(d->d_w, d->d_w_T, (size_t)N * g.kd, p->h_d_w);
    p->h_csr_ptr.assign(N + 1, 0);
    if (d->ldrt_mode == MGA_LDRT_SCATTER) {
      for (size_t k = 0; k < p->h_nbr_d.size(); ++k)
        if (p->h_nbr_d[k] >= 0) p->h_csr_ptr[p->h_nbr_d[k] + 1]++;
      for (int c = 0; c < N; ++c) p->h_csr_ptr[c + 1] += p->h_csr_ptr[c];
      std::vector<int> fill(p->h_csr_ptr.begin(), p->h_csr_ptr.end() - 1);
      p->h_csr_src.resize(p->h_csr_ptr[N]);
      slot.resize(p->h_csr_ptr[N]);
      for (int i = 0; i < N; ++i)        // ascending (i, j) == ascending slot: scatter_add order
        for (int j = 0; j < g.kd; ++j) {
          int c = p->h_nbr_d[(size_t)i * g.kd + j];
          if (c < 0) continue;
          int pos = fill[c]++;
          p->h_csr_src[pos] = i;
          slot[pos] = i * g.kd + j;
        }
    } else if (d->ldrt_mode == MGA_LDRT_GATHER) {
      for (int i = 0; i < N; ++i) {
        int cnt = 0;
        for (int j = 0; j < g.kd; ++j) cnt += p->h_nbr_d[(size_t)i * g.kd + j] >= 0;
        p->h_csr_ptr[i + 1] = p->h_csr_ptr[i] + cnt;
      }
      p->h_csr_src.resize(p->h_csr_ptr[N]);
      slot.resize(p->h_csr_ptr[N]);
      int pos = 0;
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < g.kd; ++j) {
          int c = p->h_nbr_d[(size_t)i * g.kd + j];
          if (c < 0) continue;
          p->h_csr_src[pos] = c;
          slot[pos] = i * g.kd + j;
          ++pos;
        }
    } else {
      set_error("mga_plan_create: bad ldrt_mode");
      return (MGA_ERR_INVALID);
    }
  } else if (d->temporal == MGA_TEMPORAL_LINE) {
    // first difference in time (ADMM.py:153-157, 182-186): self link with unit weight, no Q1 term
    g.kd = 1;
    g.q1 = 0;
    g.d_wT = 1;
    p->h_nbr_d.resize(N);
    p->h_d_w.assign(N, 1.0f);
    p->h_csr_ptr.resize(N + 1);
    p->h_csr_src.resize(N);
    slot.resize(N);
    for (int i = 0; i < N; ++i) { p->h_nbr_d[i] = i; p->h_csr_ptr[i] = i; p->h_csr_src[i] = i; slot[i] = i; }
    p->h_csr_ptr[N] = N;
  } else {
    // banded temporal stencil (ADMM.py:41-52, 158-164, 187-194): weights (T, skip, N)
    if (d->kd < 1 || !d->d_w) { set_error("mga_plan_create: BAND needs d_w (T,skip,N), kd=skip"); return (MGA_ERR_INVALID); }
    g.kd = 0;
    g.q1 = 0;
    g.d_wT = 1;
    g.skip = d->kd;
    p->h_csr_ptr.assign(N + 1, 0);
  }
  g.nnz = (int)p->h_csr_src.size();
  g.max_in_deg = 0;
  for (int c = 0; c < N; ++c) g.max_in_deg = std::max(g.max_in_deg, p->h_csr_ptr[c + 1] - p->h_csr_ptr[c]);
  p->h_csr_w.resize(g.nnz);
  for (int k = 0; k < g.nnz; ++k) p->h_csr_w[k] = p->h_d_w[slot[k]];

  return MGA_OK;
}

extern "C" {

int mga_version(void) { return MGA_VERSION; }
const char* mga_last_error(void) { return t_error.c_str(); }
int64_t mga_launch_count(void) { return g_launches.load(); }

int mga_plan_create(const mga_graph_desc* d, int device, mga_plan** out) {
  if (!d || !out) { set_error("mga_plan_create: NULL argument"); return MGA_ERR_INVALID; }
  *out = nullptr;
  int rc = validate_desc(d);
  if (rc) return rc;
  const int N = d->n_nodes, T = d->T;
  int ndev = 0;
  MGA_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_error("mga_plan_create: no such CUDA device"); return MGA_ERR_CUDA; }
  MGA_CUDA(cudaSetDevice(device));

  mga_plan* p = new mga_plan();
  p->device = device;
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaGetDeviceProperties"); }
  p->sm_count = prop.multiProcessorCount;
  p->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  p->max_smem_sm = (int)prop.sharedMemPerMultiprocessor;
  p->l2_bytes = (size_t)prop.l2CacheSize;

  GraphDev& g = p->g;
  auto fail = [&](int code) { mga_plan_destroy(p); return code; };
  std::vector<int> slot;
  if ((rc = host_tables(d, p, slot))) return fail(rc);
  p->ldrt_gather = d->temporal == MGA_TEMPORAL_GRAPH && d->ldrt_mode == MGA_LDRT_GATHER;
  if ((rc = upload(p, p->h_nbr_u, &g.nbr_u))) return fail(rc);
  if ((rc = upload(p, p->h_u_w, &g.u_w))) return fail(rc);
  if ((rc = upload(p, p->h_nbr_d, &g.nbr_d))) return fail(rc);
  if ((rc = upload(p, p->h_d_w, &g.d_w))) return fail(rc);
  if ((rc = upload(p, p->h_csr_ptr, &g.csr_ptr))) return fail(rc);
  if ((rc = upload(p, p->h_csr_src, &g.csr_src))) return fail(rc);
  if ((rc = upload(p, slot, &g.csr_slot))) return fail(rc);
  if ((rc = upload(p, p->h_csr_w, &g.csr_w))) return fail(rc);
  g.band_w = nullptr;
  if (d->temporal == MGA_TEMPORAL_BAND) {
    std::vector<float> bw(d->d_w, d->d_w + (size_t)T * g.skip * N);
    if ((rc = upload(p, bw, &g.band_w))) return fail(rc);
    // the reference builds these weights identical for every node (ADMM.py:44-50): keep one (T, skip) copy then
    bool uniform = true;
    std::vector<float> bu((size_t)T * g.skip);
    for (size_t k = 0; k < bu.size() && uniform; ++k) {
      bu[k] = bw[k * N];
      for (int i = 1; i < N && uniform; ++i) uniform = bw[k * N + i] == bu[k];
    }
    p->band_uniform = nullptr;
    if (uniform && (rc = upload(p, bu, &p->band_uniform))) return fail(rc);
  }
  // resident-kernel schedule: only for time-invariant tables of a size one CTA can own
  if (g.u_wT == 1 && g.d_wT == 1 && N <= 1024) {      // (banded line graph: empty temporal tables, the stencil is per node)
    ResidentSchedule sc;
    build_resident_schedule(N, g.kd, p->h_nbr_d.data(), p->h_d_w.data(), g.ku, p->h_nbr_u.data(), p->h_u_w.data(),
                            p->h_csr_ptr.data(), p->h_csr_src.data(), p->h_csr_w.data(), &sc);
    std::vector<int> ent(sc.ell_node.size() * 2);
    for (size_t k = 0; k < sc.ell_node.size(); ++k) {
      ent[2 * k] = sc.ell_node[k];
      std::memcpy(&ent[2 * k + 1], &sc.ell_w[k], sizeof(float));
    }
    p->r_ell_total = (int)sc.ell_node.size();
    p->r_kd = sc.kd; p->r_ku = sc.ku;
    if ((rc = upload(p, sc.w_self, &p->r_w_self))) return fail(rc);
    if ((rc = upload(p, sc.perm, &p->r_perm))) return fail(rc);
    if ((rc = upload(p, sc.nbr_d, &p->r_nbr_d))) return fail(rc);
    if ((rc = upload(p, sc.w_d, &p->r_w_d))) return fail(rc);
    if ((rc = upload(p, sc.nbr_u, &p->r_nbr_u))) return fail(rc);
    if ((rc = upload(p, sc.w_u, &p->r_w_u))) return fail(rc);
    if ((rc = upload(p, sc.ell_ptr, &p->r_ell_ptr))) return fail(rc);
    if ((rc = upload(p, ent, &p->r_ell_ent))) return fail(rc);
    void* ctr = nullptr;
    e = cudaMalloc(&ctr, mga_plan::kCounters * 32 * sizeof(int));
    if (e != cudaSuccess) return fail(cuda_fail(e, "cudaMalloc(window counters)"));
    p->owned.push_back(ctr);
    p->r_counters = static_cast<int*>(ctr);
    p->has_sched = true;
  }
  // chunked streaming path: time-invariant tables in reverse-Cuthill-McKee node order
  if (d->temporal != MGA_TEMPORAL_BAND && g.u_wT == 1 && g.d_wT == 1) {
    std::vector<int> perm = graph_rcm_order(N, g.kd, p->h_nbr_d.data(), g.ku, p->h_nbr_u.data());
    std::vector<int> inv(N);
    for (int k = 0; k < N; ++k) inv[perm[k]] = k;
    auto remap = [&](const std::vector<int>& nbr, const std::vector<float>& w, int K, std::vector<int>& o_n, std::vector<float>& o_w) {
      o_n.resize((size_t)N * K);
      o_w.resize((size_t)N * K);
      for (int k = 0; k < N; ++k)
        for (int j = 0; j < K; ++j) {
          const int nb = nbr[(size_t)perm[k] * K + j];
          o_n[(size_t)k * K + j] = nb >= 0 ? inv[nb] : -1;
          o_w[(size_t)k * K + j] = w[(size_t)perm[k] * K + j];
        }
    };
    std::vector<int> nd, nu, ip(N + 1, 0), is, isl;
    std::vector<float> wd, wu, iw;
    remap(p->h_nbr_d, p->h_d_w, g.kd, nd, wd);
    remap(p->h_nbr_u, p->h_u_w, g.ku, nu, wu);
    for (int k = 0; k < N; ++k) {
      const int o = perm[k];
      for (int e = p->h_csr_ptr[o]; e < p->h_csr_ptr[o + 1]; ++e) {
        is.push_back(inv[p->h_csr_src[e]]);
        iw.push_back(p->h_csr_w[e]);
        isl.push_back(inv[slot[e] / g.kd] * g.kd + slot[e] % g.kd);      // the entry's place in the reordered forward table
      }
      ip[k + 1] = (int)is.size();
    }
    Graph2& g2 = p->g2;
    g2.N = N; g2.T = T; g2.t_in = d->t_in; g2.C4 = (T + 3) / 4; g2.kd = g.kd; g2.ku = g.ku; g2.q1 = g.q1;
    stream2_tiling(&g2);
    // tables of the time-tiled kernels (staged in shared memory by every CTA): entries (byte offset of the
    // neighbour's row inside the tile, weight bits), so a gather is LDS.64 -> LDS.128 with no integer math and
    // no "-1" branch - a missing neighbour points at the own row with weight 0.  The self link of the temporal
    // graph leaves both the forward table and the in-list (the owner already holds that value): tab_wself.
    if (g2.CB3 > 0) {
      const int row_bytes = g2.CB3 * 16;
      // Node tiles (graphs too large for one CTA, whole rows of <= 8 chunks): a CTA owns NT3 RCM-consecutive nodes and
      // stages, besides them, the EXTERNAL rows their table entries point to (RCM keeps those few: ~90 per 256-512
      // nodes on the 20 000-node road graph).  Table entries hold LOCAL row offsets: own rows first, then the tile's
      // external rows in ascending order.  Single-tile plans: NT3 = N, no external rows.
      int NT = g2.NT3 > 0 ? std::min(g2.NT3, N) : N;
      const int ntile = (N + NT - 1) / NT;
      NT = (N + ntile - 1) / ntile;                 // balanced tiles: 883 nodes -> 4 x 221, not 3 x 256 + 115
      g2.NT3 = NT;
      g2.ntile3 = ntile;
      std::vector<float> wsd(N, 0.f);
      // per-node reference lists -> per-tile external rows + a local index for every reference
      struct Ext { std::vector<int> ptr, rows; };
      auto build_ext = [&](const std::vector<std::vector<std::pair<int, float>>>& refs, Ext& ex,
                           std::vector<std::vector<int>>& local) {
        ex.ptr.assign(ntile + 1, 0);
        ex.rows.clear();
        local.assign(N, {});
        for (int j = 0; j < ntile; ++j) {
          const int n0 = j * NT, n1 = std::min(N, n0 + NT);
          std::vector<int> ext;
          for (int k = n0; k < n1; ++k)
            for (auto& r : refs[k]) if (r.first < n0 || r.first >= n1) ext.push_back(r.first);
          std::sort(ext.begin(), ext.end());
          ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
          for (int k = n0; k < n1; ++k)
            for (auto& r : refs[k]) {
              const int m = r.first;
              local[k].push_back((m >= n0 && m < n1) ? m - n0
                                                     : (n1 - n0) + (int)(std::lower_bound(ext.begin(), ext.end(), m) - ext.begin()));
            }
          ex.rows.insert(ex.rows.end(), ext.begin(), ext.end());
          ex.ptr[j + 1] = (int)ex.rows.size();
        }
      };
      auto rows_of = [&](const std::vector<int>& nb, const std::vector<float>& w, int K, bool drop_self) {
        std::vector<std::vector<std::pair<int, float>>> rows(N);
        for (int k = 0; k < N; ++k)
          for (int j = 0; j < K; ++j) {
            const int m = nb[(size_t)k * K + j];
            if (m < 0) continue;
            if (drop_self && m == k) { wsd[k] += w[(size_t)k * K + j]; continue; }
            rows[k].push_back({m, w[(size_t)k * K + j]});
          }
        return rows;
      };
      auto pack = [&](const std::vector<std::vector<std::pair<int, float>>>& rows, const std::vector<std::vector<int>>& local,
                      std::vector<int>& tab) {
        int kmax = 0;
        for (int k = 0; k < N; ++k) kmax = std::max(kmax, (int)rows[k].size());
        tab.assign((size_t)N * kmax * 2, 0);
        for (int k = 0; k < N; ++k)
          for (int j = 0; j < kmax; ++j) {
            const bool has = j < (int)rows[k].size();
            const float wj = has ? rows[k][j].second : 0.f;
            tab[((size_t)k * kmax + j) * 2] = (has ? local[k][j] : k % NT) * row_bytes;      // pad: own row, weight 0
            std::memcpy(&tab[((size_t)k * kmax + j) * 2 + 1], &wj, sizeof(float));
          }
        return kmax;
      };
      std::vector<int> td, tu, ti, ip3(N + 1, 0);
      Ext ex_d, ex_u, ex_in;
      std::vector<std::vector<int>> loc_d, loc_u, loc_in;
      const auto rows_d = rows_of(nd, wd, g.kd, true);
      const auto rows_u = rows_of(nu, wu, g.ku, false);
      build_ext(rows_d, ex_d, loc_d);
      build_ext(rows_u, ex_u, loc_u);
      g2.kd3 = pack(rows_d, loc_d, td);
      g2.ku3 = pack(rows_u, loc_u, tu);
      // in-list without the self entries.  The forward table's self weight is also the in-list's only when the
      // in-list is the transpose of the forward table (not for MGA_LDRT_GATHER): check entry by entry
      bool self_ok = true;
      for (int k = 0; k < N; ++k) {
        float ws_in = 0.f;
        for (int e = ip[k]; e < ip[k + 1]; ++e)
          if (is[e] == k) ws_in += iw[e];
        if (ws_in != wsd[k]) self_ok = false;
      }
      std::vector<std::vector<std::pair<int, float>>> rows_in(N);
      for (int k = 0; k < N; ++k)
        for (int e = ip[k]; e < ip[k + 1]; ++e)
          if (!(self_ok && is[e] == k)) rows_in[k].push_back({is[e], iw[e]});     // !self_ok: every entry stays, self included
      build_ext(rows_in, ex_in, loc_in);
      int in_max = 0;
      for (int k = 0; k < N; ++k) {
        for (size_t e = 0; e < rows_in[k].size(); ++e) {
          int wbits;
          std::memcpy(&wbits, &rows_in[k][e].second, sizeof(float));
          ti.push_back(loc_in[k][e] * row_bytes);
          ti.push_back(wbits);
        }
        ip3[k + 1] = (int)ti.size() / 2;
      }
      int rmax = 0;
      for (int j = 0; j < ntile; ++j) {
        const int n0 = j * NT, n1 = std::min(N, n0 + NT);
        in_max = std::max(in_max, ip3[n1] - ip3[n0]);
        for (const Ext* ex : {&ex_d, &ex_u, &ex_in}) rmax = std::max(rmax, (n1 - n0) + ex->ptr[j + 1] - ex->ptr[j]);
      }
      g2.R3 = rmax;
      g2.in_max3 = in_max;
      // rows of k3_ldrt_lhs in descending in-list length (inside each node tile): the 4-16 nodes a warp walks together
      // have similar lengths (the in-degree of a kNN graph is far from uniform: hubs), so the gather loop is not
      // padded to the longest list of an arbitrary group
      std::vector<int> ord(N);
      for (int k = 0; k < N; ++k) ord[k] = k;
      // (rows of < 64 B stay in natural order: scattering them costs more DRAM sectors than the balance saves -
      // measured T = 12: 56.6 vs 55.6 ms per step; T = 288: 38.4 vs 39.9)
      const char* es = std::getenv("MGA_S3_SORT");
      if (es ? std::atoi(es) != 0 : g2.CB3 >= 4)
        for (int j = 0; j < ntile; ++j)
          std::stable_sort(ord.begin() + j * NT, ord.begin() + std::min(N, (j + 1) * NT),
                           [&](int x, int y) { return ip3[x + 1] - ip3[x] > ip3[y + 1] - ip3[y]; });
      if ((rc = upload(p, ord, &g2.ord3))) return fail(rc);
      {
        // k4_cg walks the rows in passes of NC / 8 rows, 4 rows per warp and pass: deal the groups of 4 (similar list
        // lengths inside a group) to the warps forwards and backwards in turn, so that no warp collects the long lists
        // of every pass (measured on the PEMS04 graph: longest / mean warp 54 / 39 list steps -> 41 / 39)
        int cons;
        k4_env(&cons);
        const int nw = cons / 32;
        std::vector<int> ord4(ord);
        if (ntile == 1) {
          const int ngrp = (N + 3) / 4;
          for (int pass = 0; pass * nw < ngrp; ++pass) {
            const int g0 = pass * nw, cnt = std::min(nw, ngrp - g0);
            // (a partial last pass stays in order: its short group is the tail; so does a full last pass whose last group
            // holds fewer than 4 rows, N % 4 != 0 - mirroring it would index past the N rows)
            if (!(pass & 1) || cnt < nw || (g0 + cnt) * 4 > N) continue;
            for (int w = 0; w < cnt; ++w)
              for (int r = 0; r < 4; ++r) ord4[(g0 + w) * 4 + r] = ord[(g0 + cnt - 1 - w) * 4 + r];
          }
        }
        if ((rc = upload(p, ord4, &g2.ord4))) return fail(rc);
      }
      g2.in_self3 = self_ok ? 1 : 0;
      g2.in_ptr3_total = (int)ti.size() / 2;
      const int* dev_tab = nullptr;
      if ((rc = upload(p, td, &dev_tab))) return fail(rc);
      g2.tab_d = reinterpret_cast<const int2*>(dev_tab);
      if ((rc = upload(p, tu, &dev_tab))) return fail(rc);
      g2.tab_u = reinterpret_cast<const int2*>(dev_tab);
      if ((rc = upload(p, ti, &dev_tab))) return fail(rc);
      g2.tab_in3 = reinterpret_cast<const int2*>(dev_tab);
      if ((rc = upload(p, ip3, &g2.in_ptr3))) return fail(rc);
      if ((rc = upload(p, wsd, &g2.wself_d))) return fail(rc);
      if ((rc = upload(p, ex_d.ptr, &g2.extp_d)) || (rc = upload(p, ex_d.rows, &g2.ext_d))) return fail(rc);
      if ((rc = upload(p, ex_u.ptr, &g2.extp_u)) || (rc = upload(p, ex_u.rows, &g2.ext_u))) return fail(rc);
      if ((rc = upload(p, ex_in.ptr, &g2.extp_in)) || (rc = upload(p, ex_in.rows, &g2.ext_in))) return fail(rc);
      // dynamic shared memory of the three kernels: R3 rows of (tile, tile2, halo) per buffer + halo2 + ext list, NT3 rows
      // of wself (+ in-list offsets, row order) + the tile's table slice
      const int nb3 = g2.db3 ? 2 : 1;
      const size_t tile_b = (size_t)rmax * (2 * nb3 * (size_t)row_bytes + (nb3 + 2) * 4) + (size_t)NT * 4 + 8;
      g2.smem3_d = (int)(tile_b + (size_t)NT * g2.kd3 * 8);
      g2.smem3_u = (int)(tile_b + (size_t)NT * g2.ku3 * 8);
      g2.smem3_in = (int)(tile_b + (size_t)((NT + 2) & ~1) * 4 + (size_t)((NT + 1) & ~1) * 4 + (size_t)in_max * 8);
      const int smem_max = std::max(g2.smem3_d, std::max(g2.smem3_u, g2.smem3_in));
      if (smem_max > p->max_smem_optin - 1024) g2.CB3 = 0;
      else if (ntile > 1 && smem_max > (228 * 1024) / 2 - 1024) { g2.one3 = 1; stream2_threads3(&g2); }   // one CTA of 1024 threads per SM
      if (std::getenv("MGA_S3_VERBOSE"))
        std::fprintf(stderr, "[mga] time-tiled kernels: N=%d C4=%d db=%d CB3=%d NB3t=%d tiles3=%d NT3=%d ntile3=%d R3=%d kd3=%d ku3=%d in=%d self=%d smem d/u/in=%d/%d/%d\n",
                     N, g2.C4, g2.db3, g2.CB3, g2.NB3t, g2.tiles3, g2.NT3, g2.ntile3, g2.R3, g2.kd3, g2.ku3, g2.in_ptr3_total, g2.in_self3, g2.smem3_d, g2.smem3_u,
                     g2.smem3_in);
    }
    if ((rc = upload(p, perm, &g2.perm))) return fail(rc);
    if ((rc = upload(p, nd, &g2.nbr_d))) return fail(rc);
    if ((rc = upload(p, wd, &g2.w_d))) return fail(rc);
    if ((rc = upload(p, nu, &g2.nbr_u))) return fail(rc);
    if ((rc = upload(p, wu, &g2.w_u))) return fail(rc);
    if ((rc = upload(p, ip, &g2.in_ptr))) return fail(rc);
    if ((rc = upload(p, is, &g2.in_src))) return fail(rc);
    if ((rc = upload(p, iw, &g2.in_w))) return fail(rc);
    if ((rc = upload(p, isl, &g2.in_slot))) return fail(rc);
    p->has_s2 = true;
  }
  p->pinned_bytes = 1 << 16;
  e = cudaMallocHost(&p->pinned, p->pinned_bytes);
  if (e != cudaSuccess) { int c = cuda_fail(e, "cudaMallocHost"); return fail(c); }
  *out = p;
  return MGA_OK;
}

// Host-only: build the tables and the resident schedule for `d` and verify that the schedule is a
// pure re-ordering (node permutation, per-row visit order, in-list placement) of the input.
int mga_schedule_selfcheck(const mga_graph_desc* d, double* stats) {
  if (!d) { set_error("mga_schedule_selfcheck: NULL argument"); return MGA_ERR_INVALID; }
  int rc = validate_desc(d);
  if (rc) return rc;
  if (d->temporal == MGA_TEMPORAL_BAND) { set_error("no resident schedule for the banded stencil"); return MGA_ERR_UNSUPPORTED; }
  mga_plan tmp;
  std::vector<int> slot;
  if ((rc = host_tables(d, &tmp, slot))) return rc;
  const GraphDev& g = tmp.g;
  if (g.u_wT != 1 || g.d_wT != 1) { set_error("no resident schedule for time-varying weights"); return MGA_ERR_UNSUPPORTED; }
  const int N = g.N;
  ResidentSchedule sc;
  build_resident_schedule(N, g.kd, tmp.h_nbr_d.data(), tmp.h_d_w.data(), g.ku, tmp.h_nbr_u.data(), tmp.h_u_w.data(),
                          tmp.h_csr_ptr.data(), tmp.h_csr_src.data(), tmp.h_csr_w.data(), &sc);
  auto bad = [&](const char* what) { set_error(std::string("schedule self-check failed: ") + what); return MGA_ERR_INVALID; };
  // (a) permutation
  std::vector<char> hit(N, 0);
  for (int p = 0; p < N; ++p) {
    const int o = sc.perm[p];
    if (o < 0 || o >= N || hit[o]) return bad("perm is not a permutation");
    hit[o] = 1;
    if (sc.inv[o] != p) return bad("inv is not the inverse of perm");
  }
  typedef std::pair<int, float> Ent;
  auto same_multiset = [](std::vector<Ent> a, std::vector<Ent> b) {
    std::sort(a.begin(), a.end());
    std::sort(b.begin(), b.end());
    return a == b;
  };
  // (b) forward rows: same (neighbour, weight) multiset once "-1" slots (weight 0 on the zero row in
  // the schedule) and, for the temporal table, self links (kept as w_self) are set aside
  auto check_fwd = [&](int K, const std::vector<int>& nbr, const std::vector<float>& w, int Ks,
                       const std::vector<int>& s_n, const std::vector<float>& s_w, bool self_out) {
    for (int p = 0; p < N; ++p) {
      const int o = sc.perm[p];
      std::vector<Ent> a, b;
      float wself = 0.f;
      for (int j = 0; j < K; ++j) {
        const int nb = nbr[(size_t)o * K + j];
        if (nb < 0) continue;
        if (self_out && nb == o) wself += w[(size_t)o * K + j];
        else a.emplace_back(nb, w[(size_t)o * K + j]);
      }
      for (int j = 0; j < Ks; ++j) {
        const int sn = s_n[(size_t)p * Ks + j];
        if (sn < 0 || sn >= N + 8) return false;
        if (sn >= N) { if (s_w[(size_t)p * Ks + j] != 0.f) return false; continue; }
        b.emplace_back(sc.perm[sn], s_w[(size_t)p * Ks + j]);
      }
      if (!same_multiset(a, b)) return false;
      if (self_out && wself != sc.w_self[p]) return false;
    }
    return true;
  };
  if (!check_fwd(g.kd, tmp.h_nbr_d, tmp.h_d_w, sc.kd, sc.nbr_d, sc.w_d, true)) return bad("temporal table rows differ");
  if (!check_fwd(g.ku, tmp.h_nbr_u, tmp.h_u_w, sc.ku, sc.nbr_u, sc.w_u, false)) return bad("spatial table rows differ");
  // (c) in-list: every node's ELL column holds exactly its CSR row minus the self link (+ zero-weight padding)
  const int n_warps = (N + 31) / 32;
  for (int p = 0; p < n_warps * 32; ++p) {
    const int w = p / 32, l = p % 32;
    std::vector<Ent> a, b;
    if (p < N) {
      const int o = sc.perm[p];
      float wself = 0.f;
      for (int e = tmp.h_csr_ptr[o]; e < tmp.h_csr_ptr[o + 1]; ++e) {
        if (tmp.h_csr_src[e] == o) wself += tmp.h_csr_w[e];
        else a.emplace_back(tmp.h_csr_src[e], tmp.h_csr_w[e]);
      }
      if (wself != sc.w_self[p]) return bad("in-list self link differs from the forward one");
    }
    for (int e = sc.ell_ptr[w]; e < sc.ell_ptr[w + 1]; ++e) {
      const int sn = sc.ell_node[(size_t)e * 32 + l];
      const float sw = sc.ell_w[(size_t)e * 32 + l];
      if (sn < 0 || sn >= N + 8) return bad("in-list node out of range");
      if (sn >= N) { if (sw != 0.f) return bad("padding entry with weight"); continue; }
      b.emplace_back(sc.perm[sn], sw);
    }
    if (!same_multiset(a, b)) return bad("in-list differs");
  }
  if (stats) {
    // average wavefronts per quarter-warp phase, before (input order) and after scheduling
    auto phase = [&](const std::vector<int>& nodes) {
      int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      std::vector<int> u(nodes);
      std::sort(u.begin(), u.end());
      u.erase(std::unique(u.begin(), u.end()), u.end());
      int m = 1;
      for (int n : u) if (n >= 0) m = std::max(m, ++cnt[n & 7]);     // zero rows (>= N) occupy a bank group too
      return m;
    };
    auto fwd_cost = [&](int K, const std::vector<int>& tab) {
      double tot = 0; int c = 0;
      for (int q0 = 0; q0 < N; q0 += 8)
        for (int j = 0; j < K; ++j) {
          std::vector<int> t;
          for (int r = q0; r < std::min(N, q0 + 8); ++r) t.push_back(tab[(size_t)r * K + j]);
          tot += phase(t); ++c;
        }
      return c ? tot / c : 1.0;
    };
    stats[0] = fwd_cost(g.kd, tmp.h_nbr_d);
    stats[1] = fwd_cost(sc.kd, sc.nbr_d);
    double tot = 0, tot2 = 0; int c = 0, c2 = 0;
    for (int q0 = 0; q0 < N; q0 += 8) {
      int m = 0;
      for (int r = q0; r < std::min(N, q0 + 8); ++r) m = std::max(m, tmp.h_csr_ptr[r + 1] - tmp.h_csr_ptr[r]);
      for (int e = 0; e < m; ++e) {
        std::vector<int> t;
        for (int r = q0; r < std::min(N, q0 + 8); ++r)
          if (tmp.h_csr_ptr[r] + e < tmp.h_csr_ptr[r + 1]) t.push_back(tmp.h_csr_src[tmp.h_csr_ptr[r] + e]);
        tot += phase(t); ++c;
      }
    }
    for (int w = 0; w < n_warps; ++w)
      for (int e = sc.ell_ptr[w]; e < sc.ell_ptr[w + 1]; ++e)
        for (int q0 = 0; q0 < 32; q0 += 8) {
          std::vector<int> t;
          for (int l = q0; l < q0 + 8; ++l) t.push_back(sc.ell_node[(size_t)e * 32 + l]);
          tot2 += phase(t); ++c2;
        }
    stats[2] = c ? tot / c : 1.0;
    stats[3] = c2 ? tot2 / c2 : 1.0;
    // in-list steps a warp walks, summed over warps: input order with self links vs the schedule
    int before = 0;
    for (int w = 0; w < n_warps; ++w) {
      int m = 0;
      for (int r = w * 32; r < std::min(N, w * 32 + 32); ++r) m = std::max(m, tmp.h_csr_ptr[r + 1] - tmp.h_csr_ptr[r]);
      before += m;
    }
    stats[4] = before;
    stats[5] = sc.ell_ptr[n_warps];
  }
  return MGA_OK;
}

void mga_plan_destroy(mga_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  for (void* d : p->owned) cudaFree(d);
  if (p->ws.base) cudaFree(p->ws.base);
  if (p->ws_host_io.base) cudaFree(p->ws_host_io.base);
  if (p->pinned) cudaFreeHost(p->pinned);
  if (p->pipe_host) cudaFreeHost(p->pipe_host);
  for (auto& s : p->io_streams) if (s) cudaStreamDestroy(s);
  for (auto& ev : p->io_events) if (ev) cudaEventDestroy(ev);
  delete p;
}

int mga_plan_resident_eligible(const mga_plan* p, int dtype) { return p && resident_eligible(p, dtype) ? 1 : 0; }

int mga_plan_info(const mga_plan* p, int32_t* sm_count, int32_t* smem, int32_t* threads, int32_t* max_in) {
  if (!p) { set_error("mga_plan_info: NULL plan"); return MGA_ERR_INVALID; }
  if (sm_count) *sm_count = p->sm_count;
  int th = 0;
  int sb = resident_eligible(p, MGA_F32) ? resident_smem_bytes(p, &th) : 0;
  if (smem) *smem = sb;
  if (threads) *threads = th;
  if (max_in) *max_in = p->g.max_in_deg;
  return MGA_OK;
}

static int check_common(const mga_plan* p, const mga_params* prm, int64_t B, int dtype, const char* who) {
  if (!p) { set_error(std::string(who) + ": NULL plan"); return MGA_ERR_INVALID; }
  if (B <= 0) { set_error(std::string(who) + ": batch must be positive"); return MGA_ERR_INVALID; }
  if (dtype != MGA_F32 && dtype != MGA_F64) { set_error(std::string(who) + ": dtype must be MGA_F32 or MGA_F64"); return MGA_ERR_INVALID; }
  if (prm && (prm->ablation < MGA_ABL_NONE || prm->ablation > MGA_ABL_UT)) { set_error(std::string(who) + ": bad ablation"); return MGA_ERR_INVALID; }
  cudaError_t e = cudaSetDevice(p->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  return MGA_OK;
}

int mga_apply(mga_plan* p, int op, const mga_params* prm, const void* x, void* y, const void* mask, int64_t B,
              int dtype, void* stream) {
  int rc = check_common(p, prm, B, dtype, "mga_apply");
  if (rc) return rc;
  if (!x || !y || x == y) { set_error("mga_apply: x and y must be distinct non-NULL device pointers"); return MGA_ERR_INVALID; }
  if (op < MGA_OP_LU || op > MGA_OP_LHS_ZD) { set_error("mga_apply: bad op"); return MGA_ERR_INVALID; }
  if (op >= MGA_OP_LHS_X && !prm) { set_error("mga_apply: LHS operators need params"); return MGA_ERR_INVALID; }
  return stream_apply(p, op, prm, x, y, mask, B, dtype, (cudaStream_t)stream);
}

int mga_cg_solve(mga_plan* p, int system, const mga_params* prm, const void* rhs, void* x, const void* mask_first,
                 int64_t B, int dtype, int max_iter, double tol, int32_t* iters_out, void* alpha, void* beta,
                 void* stream) {
  int rc = check_common(p, prm, B, dtype, "mga_cg_solve");
  if (rc) return rc;
  if (!prm || !rhs || !x || max_iter < 0 || system < MGA_SYS_X || system > MGA_SYS_ZD) {
    set_error("mga_cg_solve: bad argument");
    return MGA_ERR_INVALID;
  }
  // fixed iteration count, forecasting operator, a window that fits one CTA: the whole solve in one launch
  const bool can_res = tol <= 0 && !mask_first && max_iter > 0 && prm->ablation == MGA_ABL_NONE && resident_eligible(p, dtype) &&
                       p->g.temporal != MGA_TEMPORAL_BAND;
  if (p->cg_mode == MGA_MODE_RESIDENT && !can_res) {
    set_error("mga_cg_solve: resident mode needs fp32, tol <= 0, no mask, ablation None and a resident-eligible plan");
    return MGA_ERR_UNSUPPORTED;
  }
  if (p->cg_mode != MGA_MODE_STREAMING && p->cg_mode != MGA_MODE_STREAMING_POINT && can_res) {
    if (iters_out) *iters_out = -1;          // ran all iterations: "not converged" in the reference's terms (ADMM.py:368)
    return resident_cg(p, system, prm, rhs, x, B, max_iter, alpha, beta, (cudaStream_t)stream);
  }
  if (p->cg_mode != MGA_MODE_STREAMING_POINT && tol <= 0 && !mask_first && max_iter > 0 && prm->ablation == MGA_ABL_NONE &&
      stream2_eligible(p, dtype)) {
    if (iters_out) *iters_out = -1;
    return stream2_cg(p, system, prm, rhs, x, B, max_iter, alpha, beta, (cudaStream_t)stream);
  }
  return stream_cg(p, system, prm, rhs, x, mask_first, B, dtype, max_iter, tol, iters_out, alpha, beta,
                   (cudaStream_t)stream);
}

int mga_plan_set_cg_mode(mga_plan* p, int mode) {
  if (!p || mode < MGA_MODE_AUTO || mode > MGA_MODE_STREAMING_POINT) { set_error("mga_plan_set_cg_mode: bad argument"); return MGA_ERR_INVALID; }
  p->cg_mode = mode;
  return MGA_OK;
}

int mga_admm_solve(mga_plan* p, const mga_params* prm, const void* y, int y_rows, const void* mask, void* x_out,
                   int64_t B, int dtype, int n_outer, int max_cg, double cg_tol, double admm_tol, double t_mean,
                   double t_var, int want_diag, const mga_admm_outputs* outs, int mode, void* stream) {
  int rc = check_common(p, prm, B, dtype, "mga_admm_solve");
  if (rc) return rc;
  if (!prm || !y || !x_out || n_outer < 0 || max_cg < 0) { set_error("mga_admm_solve: bad argument"); return MGA_ERR_INVALID; }
  const bool forecast = (mask == nullptr);
  if (forecast && y_rows != p->g.t_in) { set_error("mga_admm_solve: y must have t_in rows"); return MGA_ERR_INVALID; }
  if (!forecast && y_rows != p->g.T) { set_error("mga_admm_solve: mask mode needs y with T rows"); return MGA_ERR_INVALID; }
  mga_admm_outputs none{};
  if (!outs) outs = &none;
  const bool fixed = cg_tol <= 0 && admm_tol <= 0;
  const bool can_res = fixed && prm->ablation == MGA_ABL_NONE && resident_eligible(p, dtype) &&     // forecast or mask mode
                       !(mask && p->g.temporal == MGA_TEMPORAL_BAND);
  if (mode == MGA_MODE_RESIDENT && !can_res) {
    set_error("mga_admm_solve: resident mode needs fp32, fixed iteration counts (cg_tol, admm_tol <= 0), ablation None, "
              "time-invariant weights, a window of at most 1024 (node, 12-step slab) threads with T <= 24, at most 10 "
              "neighbours per row after self links are dropped, and no mask on the banded line graph");
    return MGA_ERR_UNSUPPORTED;
  }
  if (mode != MGA_MODE_STREAMING && mode != MGA_MODE_STREAMING_POINT && can_res)
    return resident_admm(p, prm, y, mask, x_out, B, n_outer, max_cg, t_mean, t_var, want_diag, outs, (cudaStream_t)stream);
  // The reference's own call pattern - one window, tolerances, float64 (ADMM.py:76-80) -, float64 batches and plans with
  // time-varying weight tables (T, N, k): one thread-block cluster per window, the stop tests decided on the device
  // (MGA_CLUSTER=0: the general kernels instead)
  {
    static const bool cluster_on = [] { const char* e = std::getenv("MGA_CLUSTER"); return !e || std::atoi(e) != 0; }();
    if (cluster_on && mode != MGA_MODE_STREAMING && mode != MGA_MODE_STREAMING_POINT && prm->ablation == MGA_ABL_NONE &&
        (fixed ? (dtype == MGA_F64 || p->g.u_wT > 1 || p->g.d_wT > 1) : B == 1) && cluster_eligible(p, dtype))
      return cluster_admm(p, prm, y, mask, x_out, B, dtype, n_outer, max_cg, cg_tol, admm_tol, t_mean, t_var, want_diag, outs,
                          (cudaStream_t)stream);
  }
  if (mode != MGA_MODE_STREAMING_POINT && forecast && fixed && prm->ablation == MGA_ABL_NONE && stream2_eligible(p, dtype))
    return stream2_admm(p, prm, y, x_out, B, n_outer, max_cg, t_mean, t_var, want_diag, outs, (cudaStream_t)stream);
  return stream_admm(p, prm, y, y_rows, mask, x_out, B, dtype, n_outer, max_cg, cg_tol, admm_tol, t_mean, t_var,
                     want_diag, outs, (cudaStream_t)stream);
}

// The cluster kernel on its own (mga_admm_solve picks it for tolerance mode at B = 1 and for float64): forecasting mode,
// ablation None; cg_tol / admm_tol > 0 need B = 1.
int mga_cluster_solve(mga_plan* p, const mga_params* prm, const void* y, void* x_out, int64_t B, int dtype, int n_outer,
                      int max_cg, double cg_tol, double admm_tol, double t_mean, double t_var, int want_diag,
                      const mga_admm_outputs* outs, void* stream) {
  int rc = check_common(p, prm, B, dtype, "mga_cluster_solve");
  if (rc) return rc;
  if (!prm || !y || !x_out || n_outer < 0 || max_cg < 0) { set_error("mga_cluster_solve: bad argument"); return MGA_ERR_INVALID; }
  if (prm->ablation != MGA_ABL_NONE || !cluster_eligible(p, dtype)) {
    set_error("mga_cluster_solve: needs ablation None, no banded line graph, T <= 32 and N <= 1024");
    return MGA_ERR_UNSUPPORTED;
  }
  mga_admm_outputs none{};
  if (!outs) outs = &none;
  return cluster_admm(p, prm, y, nullptr, x_out, B, dtype, n_outer, max_cg, cg_tol, admm_tol, t_mean, t_var, want_diag, outs,
                      (cudaStream_t)stream);
}

// ---- end-to-end entry point with host buffers ------------------------------------------------------------------
static int host_io_setup(mga_plan* p) {
  for (auto& s : p->io_streams) if (!s) MGA_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  for (auto& ev : p->io_events) if (!ev) MGA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if (!p->pipe_dev) {
    const size_t nb = (2 * mga_plan::kPipeChunks + 8) * sizeof(int);
    MGA_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->pipe_dev), nb));
    p->owned.push_back(p->pipe_dev);
    MGA_CUDA(cudaMemset(p->pipe_dev, 0, nb));
  }
  if (!p->pipe_host) {
    const size_t nb = 2 * mga_plan::kPipeChunks * sizeof(int);
    MGA_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p->pipe_host), nb, cudaHostAllocMapped));
    std::memset(p->pipe_host, 0, nb);
    MGA_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&p->pipe_host_dev), p->pipe_host, 0));
  }
  return MGA_OK;
}

// Resident mode: ONE persistent launch over the whole batch.  The upload stream copies y chunk by chunk and, in stream
// order behind each chunk, a 4-byte "ready" word; the kernel's CTAs wait for the word of the chunk their next window
// lives in, so the solve starts as soon as the first chunk has landed and always runs on a full grid (chunked launches
// each under-filled the 2-CTAs-per-SM grid and paid a tail per chunk).  The CTA that finishes the last window of a
// chunk raises a flag in mapped host memory; this thread polls the flags and queues the chunk's x download, so copies
// in both directions overlap the solve.  alpha / beta come back in one copy behind the kernel.
static int solve_host_pipelined(mga_plan* p, const mga_params* prm, const char* y_host, int y_rows, char* x_host, int64_t B,
                                int n_outer, int max_cg, double t_mean, double t_var, int want_diag, double* diag_host,
                                double* dx_sum_host, void* alpha_host, void* beta_host, int mode, int64_t chunk) {
  const GraphDev& g = p->g;
  const size_t es = 4;
  const size_t y_win = (size_t)y_rows * g.N * es, x_win = (size_t)g.T * g.N * es;
  const size_t diag_n = (size_t)n_outer * MGA_DIAG_COLS, dx_n = (size_t)n_outer * g.T * g.N;
  const bool want_coef = alpha_host && beta_host && max_cg > 0 && n_outer > 0;
  // windows per launch: the staging of one launch stays below a cap (default 16 GB); larger batches take several launches
  double cap_gb = 16.0;
  if (const char* e = std::getenv("MGA_HOST_STAGING_GB")) cap_gb = std::max(0.001, std::atof(e));
  const int64_t Bs = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)(cap_gb * 1e9 / (double)(y_win + x_win))));
  const size_t coef_rows = (size_t)n_outer * 3 * max_cg;
  const size_t coef_bytes = want_coef ? coef_rows * (size_t)Bs * es : 0;     // (rows, windows of one launch), compact
  // staging: y | x | alpha | beta | diag | dx_sum - the last four back to back, so that a caller who keeps its four
  // result arrays back to back as well gets them in ONE copy behind the kernel
  size_t off_x = ((size_t)Bs * y_win + 255) & ~(size_t)255;
  size_t off_a = (off_x + (size_t)Bs * x_win + 255) & ~(size_t)255;
  size_t off_b = off_a + coef_bytes;
  size_t off_d = (off_b + coef_bytes + 7) & ~(size_t)7;
  int rc = ensure_workspace(p, p->ws_host_io, off_d + (diag_n + dx_n) * sizeof(double));
  if (rc) return rc;
  if ((rc = host_io_setup(p))) return rc;
  char* base = static_cast<char*>(p->ws_host_io.base);
  char* dy = base;
  char* dx = base + off_x;
  double* d_diag = reinterpret_cast<double*>(base + off_d);
  double* d_dx = d_diag + diag_n;
  cudaStream_t s_up = p->io_streams[0], s_run = p->io_streams[1], s_dn = p->io_streams[3];
  mga_admm_outputs outs{};
  if (want_diag) {
    outs.diag = d_diag;
    outs.dx_sum = d_dx;
    MGA_CUDA(cudaMemsetAsync(d_diag, 0, (diag_n + dx_n) * sizeof(double), s_run));
  }
  const bool trace = std::getenv("MGA_HOST_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(); };
  double t_launch = 0, t_uploads = 0, t_first = 0, t_last = 0, t_run = 0;
  bool one_copy = false;
  for (int64_t s0 = 0; s0 < B; s0 += Bs) {
    const int64_t nbs = std::min(Bs, B - s0);
    int64_t ck = chunk > 0 ? chunk : std::max<int64_t>(64, (nbs + 63) / 64);
    if (const char* e = std::getenv("MGA_HOST_CHUNK")) ck = std::max<long>(1, std::atol(e));
    ck = std::min(ck, nbs);
    if ((nbs + ck - 1) / ck > mga_plan::kPipeChunks) ck = (nbs + mga_plan::kPipeChunks - 1) / mga_plan::kPipeChunks;
    const int nchunk = (int)((nbs + ck - 1) / ck);
    // uploads in fewer, larger pieces than downloads (every piece is two API calls ahead of the launch): 8 for small batches
    int64_t ck_up = ck;
    if (nbs <= 4096) ck_up = std::max<int64_t>(ck, (((nbs + 7) / 8 + ck - 1) / ck) * ck);
    const int nup = (int)((nbs + ck_up - 1) / ck_up);
    const int epoch = (p->pipe_epoch = p->pipe_epoch >= (1 << 30) ? 1 : p->pipe_epoch + 1);
    int* host_done = p->pipe_host;
    int* host_word = p->pipe_host + mga_plan::kPipeChunks;      // source of the "ready" copies
    for (int c = 0; c < nchunk; ++c) host_word[c] = epoch;
    HostPipe pipe{};
    pipe.ready = p->pipe_dev;
    pipe.done = p->pipe_dev + mga_plan::kPipeChunks;
    pipe.abort_flag = p->pipe_dev + 2 * mga_plan::kPipeChunks;
    pipe.host_done = p->pipe_host_dev;
    pipe.chunk = (int)ck;
    pipe.chunk_up = (int)ck_up;
    pipe.epoch = epoch;
    if (want_coef) {
      outs.alpha = base + off_a;
      outs.beta = base + off_b;
    }
    auto upload = [&](int c) -> int {
      const int64_t b0 = (int64_t)c * ck_up, nb = std::min(ck_up, nbs - b0);
      MGA_CUDA(cudaMemcpyAsync(dy + (size_t)b0 * y_win, y_host + (size_t)(s0 + b0) * y_win, (size_t)nb * y_win,
                               cudaMemcpyHostToDevice, s_up));
      MGA_CUDA(cudaMemcpyAsync(const_cast<int*>(pipe.ready) + c, host_word + c, sizeof(int), cudaMemcpyHostToDevice, s_up));
      return MGA_OK;
    };
    MGA_CUDA(cudaMemsetAsync(pipe.done, 0, (mga_plan::kPipeChunks + 8) * sizeof(int), s_run));     // counters + abort flag
    // Every upload is queued BEFORE the launch: the kernel waits for data, so nothing it waits for may depend on a host
    // action that follows the launch call - under a profiler or CUDA_LAUNCH_BLOCKING the launch call only returns when the
    // kernel is over (measured under ncu with the uploads queued after the launch: every CTA ran into the 4 s give-up).
    for (int c = 0; c < nup; ++c)
      if ((rc = upload(c))) { cudaDeviceSynchronize(); return rc; }
    t_uploads = us();
    p->pipe = &pipe;
    rc = mga_admm_solve(p, prm, dy, y_rows, nullptr, dx, nbs, MGA_F32, n_outer, max_cg, -1.0, -1.0, t_mean, t_var,
                        want_diag | 2, &outs, mode, s_run);
    p->pipe = nullptr;
    if (rc) { cudaDeviceSynchronize(); return rc; }
    t_launch = us();
    // hand finished chunks to the download stream as the kernel reports them
    bool kernel_over = false;
    for (int c = 0; c < nchunk; ++c) {
      volatile int* flag = host_done + c;
      for (unsigned spin = 0; *flag != epoch; ++spin) {
        if ((spin & 1023u) == 1023u && !kernel_over) {
          const cudaError_t q = cudaStreamQuery(s_run);
          if (q == cudaSuccess) kernel_over = true;                   // one more look at the flag, then it is an error
          else if (q != cudaErrorNotReady) return cuda_fail(q, "resident kernel (host entry)");
        } else if (kernel_over && (spin & 1023u) == 1023u) {
          cudaDeviceSynchronize();
          set_error("mga_admm_solve_host: the solve ended without finishing every chunk (an upload never arrived)");
          return MGA_ERR_CUDA;
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
      }
      if (c == 0) t_first = us();
      if (c == nchunk - 1) t_last = us();
      const int64_t b0 = (int64_t)c * ck, nb = std::min(ck, nbs - b0);
      MGA_CUDA(cudaMemcpyAsync(x_host + (size_t)(s0 + b0) * x_win, dx + (size_t)b0 * x_win, (size_t)nb * x_win,
                               cudaMemcpyDeviceToHost, s_dn));
    }
    MGA_CUDA(cudaStreamSynchronize(s_run));
    t_run = us();
    one_copy = want_diag && diag_host && dx_sum_host && nbs == B && off_d == off_b + coef_bytes &&
               reinterpret_cast<char*>(dx_sum_host) == reinterpret_cast<char*>(diag_host) + diag_n * sizeof(double) &&
               (!want_coef || (static_cast<char*>(beta_host) == static_cast<char*>(alpha_host) + coef_bytes &&
                               reinterpret_cast<char*>(diag_host) == static_cast<char*>(beta_host) + coef_bytes));
    if (one_copy) {
      MGA_CUDA(cudaMemcpyAsync(want_coef ? alpha_host : static_cast<void*>(diag_host), base + (want_coef ? off_a : off_d),
                               (want_coef ? 2 * coef_bytes : 0) + (diag_n + dx_n) * sizeof(double), cudaMemcpyDeviceToHost, s_dn));
    } else if (want_coef) {        // columns [s0, s0 + nbs) of the caller's (rows, B) arrays
      MGA_CUDA(cudaMemcpy2DAsync(static_cast<char*>(alpha_host) + (size_t)s0 * es, (size_t)B * es, base + off_a, (size_t)nbs * es,
                                 (size_t)nbs * es, coef_rows, cudaMemcpyDeviceToHost, s_dn));
      MGA_CUDA(cudaMemcpy2DAsync(static_cast<char*>(beta_host) + (size_t)s0 * es, (size_t)B * es, base + off_b, (size_t)nbs * es,
                                 (size_t)nbs * es, coef_rows, cudaMemcpyDeviceToHost, s_dn));
    }
    if (s0 + nbs < B) MGA_CUDA(cudaStreamSynchronize(s_dn));          // the next launch reuses the staging
  }
  if (want_diag && !one_copy) {
    if (diag_host) MGA_CUDA(cudaMemcpyAsync(diag_host, d_diag, diag_n * sizeof(double), cudaMemcpyDeviceToHost, s_dn));
    if (dx_sum_host) MGA_CUDA(cudaMemcpyAsync(dx_sum_host, d_dx, dx_n * sizeof(double), cudaMemcpyDeviceToHost, s_dn));
  }
  MGA_CUDA(cudaStreamSynchronize(s_dn));
  if (trace)
    std::fprintf(stderr, "[mga] host entry (us): uploads queued %.0f, launched %.0f, first chunk done %.0f, last chunk done %.0f, "
                 "kernel over %.0f, all copies back %.0f\n", t_uploads, t_launch, t_first, t_last, t_run, us());
  return MGA_OK;
}

// Streaming modes (and MGA_HOST_PIPE=0): the batch is cut into chunks; chunk c+1 is uploaded and chunk c-1 downloaded
// while chunk c computes (upload / run / download streams, events for ordering).
int mga_admm_solve_host(mga_plan* p, const mga_params* prm, const void* y_host, int y_rows, void* x_host, int64_t B,
                        int dtype, int n_outer, int max_cg, double t_mean, double t_var, int want_diag,
                        double* diag_host, double* dx_sum_host, void* alpha_host, void* beta_host, int mode,
                        int64_t chunk) {
  int rc = check_common(p, prm, B, dtype, "mga_admm_solve_host");
  if (rc) return rc;
  if (!prm || !y_host || !x_host) { set_error("mga_admm_solve_host: NULL buffer"); return MGA_ERR_INVALID; }
  if (y_rows != p->g.t_in) { set_error("mga_admm_solve_host: y must have t_in rows"); return MGA_ERR_INVALID; }
  if ((alpha_host != nullptr) != (beta_host != nullptr)) { set_error("mga_admm_solve_host: alpha_host and beta_host go together"); return MGA_ERR_INVALID; }
  if (n_outer < 0 || max_cg < 0) { set_error("mga_admm_solve_host: bad iteration counts"); return MGA_ERR_INVALID; }
  const GraphDev& g = p->g;
  const size_t es = dtype == MGA_F32 ? 4 : 8;
  const bool can_res = mode != MGA_MODE_STREAMING && mode != MGA_MODE_STREAMING_POINT && prm->ablation == MGA_ABL_NONE &&
                       resident_eligible(p, dtype);
  if (mode == MGA_MODE_RESIDENT && !can_res) {
    set_error("mga_admm_solve_host: resident mode is not available for this plan / dtype / ablation");
    return MGA_ERR_UNSUPPORTED;
  }
  bool pipelined = can_res && p->g.temporal != MGA_TEMPORAL_BAND;      // (the banded line graph keeps chunked launches)
  if (const char* e = std::getenv("MGA_HOST_PIPE")) pipelined = pipelined && std::atoi(e) != 0;
  if (pipelined)
    return solve_host_pipelined(p, prm, static_cast<const char*>(y_host), y_rows, static_cast<char*>(x_host), B, n_outer,
                                max_cg, t_mean, t_var, want_diag, diag_host, dx_sum_host, alpha_host, beta_host, mode, chunk);
  if (chunk <= 0) {
    int parts = 4;
    if (const char* e = std::getenv("MGA_HOST_CHUNKS")) parts = std::max(1, std::atoi(e));
    chunk = std::max<int64_t>(1, std::min<int64_t>(B, (B + parts - 1) / parts));
  }
  chunk = std::min(chunk, B);
  // chunk sizes: equal parts, or an explicit plan "n1,n2,..." (MGA_HOST_PLAN, must add up to B) for experiments
  std::vector<int64_t> sizes;
  if (const char* e = std::getenv("MGA_HOST_PLAN")) {
    int64_t sum = 0;
    for (const char* q = e; *q;) {
      char* end = nullptr;
      const long v = std::strtol(q, &end, 10);
      if (end == q || v <= 0) { sizes.clear(); break; }
      sizes.push_back(v);
      sum += v;
      q = *end == ',' ? end + 1 : end;
    }
    if (sum != B) sizes.clear();
  }
  if (sizes.empty())
    for (int64_t b0 = 0; b0 < B; b0 += chunk) sizes.push_back(std::min(chunk, B - b0));
  chunk = *std::max_element(sizes.begin(), sizes.end());
  const int64_t nchunk = (int64_t)sizes.size();
  const size_t y_win = (size_t)y_rows * g.N * es, x_win = (size_t)g.T * g.N * es;
  const size_t diag_n = (size_t)n_outer * MGA_DIAG_COLS, dx_n = (size_t)n_outer * g.T * g.N;
  const bool want_coef = alpha_host && beta_host && max_cg > 0 && n_outer > 0;
  const size_t coef_rows = (size_t)n_outer * 3 * max_cg;
  const size_t coef_slot = want_coef ? ((coef_rows * (size_t)chunk * es + 255) & ~(size_t)255) : 0;
  // device staging: kSlots y-slots, kSlots x-slots, (alpha, beta) per slot, diag + dx_sum accumulators.  Three slots: with
  // four chunks the upload of chunk 2 no longer waits for chunk 0's solve (its y slot), so all uploads run back to back
  constexpr int kSlots = 3;
  size_t off_y = 0, off_x = kSlots * (size_t)chunk * y_win, off_c = off_x + kSlots * (size_t)chunk * x_win;
  off_c = (off_c + 255) & ~(size_t)255;
  size_t off_d = off_c + 2 * kSlots * coef_slot;
  size_t total = off_d + (diag_n + dx_n) * sizeof(double);
  rc = ensure_workspace(p, p->ws_host_io, total);
  if (rc) return rc;
  if ((rc = host_io_setup(p))) return rc;
  char* base = static_cast<char*>(p->ws_host_io.base);
  // Resident mode (MGA_HOST_PIPE=0): two chunk solves may run concurrently (one stream per slot), so the tail of one
  // chunk's persistent grid overlaps the head of the next.  Streaming mode shares one workspace: one run stream.
  cudaStream_t s_up = p->io_streams[0], s_dn = p->io_streams[3];
  cudaStream_t s_run[2] = {p->io_streams[1], can_res ? p->io_streams[2] : p->io_streams[1]};
  cudaEvent_t* up_done = &p->io_events[0];    // [kSlots] y slot filled
  cudaEvent_t* run_done = &p->io_events[3];   // [kSlots] x slot filled (= y slot consumed)
  cudaEvent_t* x_free = &p->io_events[6];     // [kSlots] x slot drained
  cudaEvent_t diag_ready = p->io_events[9];
  double* d_diag = reinterpret_cast<double*>(base + off_d);
  double* d_dx = d_diag + diag_n;
  mga_admm_outputs outs{};
  if (want_diag) {
    outs.diag = d_diag;
    outs.dx_sum = d_dx;
    MGA_CUDA(cudaMemsetAsync(d_diag, 0, (diag_n + dx_n) * sizeof(double), s_run[0]));
    MGA_CUDA(cudaEventRecord(diag_ready, s_run[0]));
    if (s_run[1] != s_run[0]) MGA_CUDA(cudaStreamWaitEvent(s_run[1], diag_ready, 0));
  }
  int64_t b_next = 0;
  for (int64_t c = 0; c < nchunk; ++c) {
    const int slot = (int)(c % kSlots), rs = (int)(c & 1);      // staging slot, run stream
    const int64_t b0 = b_next, nb = sizes[c];
    b_next += nb;
    char* dy = base + off_y + (size_t)slot * chunk * y_win;
    char* dx = base + off_x + (size_t)slot * chunk * x_win;
    char* da = base + off_c + (size_t)(2 * slot) * coef_slot;
    char* db = da + coef_slot;
    if (c >= kSlots) MGA_CUDA(cudaStreamWaitEvent(s_up, run_done[slot], 0));
    MGA_CUDA(cudaMemcpyAsync(dy, static_cast<const char*>(y_host) + (size_t)b0 * y_win, (size_t)nb * y_win,
                             cudaMemcpyHostToDevice, s_up));
    MGA_CUDA(cudaEventRecord(up_done[slot], s_up));
    MGA_CUDA(cudaStreamWaitEvent(s_run[rs], up_done[slot], 0));
    if (c >= kSlots) MGA_CUDA(cudaStreamWaitEvent(s_run[rs], x_free[slot], 0));
    // diagnostics accumulate across chunks (the kernels add into diag / dx_sum); the chunk's CG coefficients land in
    // a (rows, nb) block of its slot and go to columns [b0, b0 + nb) of the caller's (rows, B) arrays
    if (want_coef) { outs.alpha = da; outs.beta = db; }
    p->res_slot = rs;
    rc = mga_admm_solve(p, prm, dy, y_rows, nullptr, dx, nb, dtype, n_outer, max_cg, -1.0, -1.0, t_mean, t_var,
                        want_diag | 2, &outs, mode, s_run[rs]);
    p->res_slot = 0;
    if (rc) { cudaDeviceSynchronize(); return rc; }
    MGA_CUDA(cudaEventRecord(run_done[slot], s_run[rs]));
    MGA_CUDA(cudaStreamWaitEvent(s_dn, run_done[slot], 0));
    MGA_CUDA(cudaMemcpyAsync(static_cast<char*>(x_host) + (size_t)b0 * x_win, dx, (size_t)nb * x_win,
                             cudaMemcpyDeviceToHost, s_dn));
    if (want_coef) {
      MGA_CUDA(cudaMemcpy2DAsync(static_cast<char*>(alpha_host) + (size_t)b0 * es, (size_t)B * es, da, (size_t)nb * es,
                                 (size_t)nb * es, coef_rows, cudaMemcpyDeviceToHost, s_dn));
      MGA_CUDA(cudaMemcpy2DAsync(static_cast<char*>(beta_host) + (size_t)b0 * es, (size_t)B * es, db, (size_t)nb * es,
                                 (size_t)nb * es, coef_rows, cudaMemcpyDeviceToHost, s_dn));
    }
    MGA_CUDA(cudaEventRecord(x_free[slot], s_dn));
  }
  cudaStream_t s_run0 = s_run[0], s_run1 = s_run[1];
  MGA_CUDA(cudaStreamSynchronize(s_run0));
  MGA_CUDA(cudaStreamSynchronize(s_run1));
  if (want_diag) {
    if (diag_host) MGA_CUDA(cudaMemcpy(diag_host, d_diag, diag_n * sizeof(double), cudaMemcpyDeviceToHost));
    if (dx_sum_host) MGA_CUDA(cudaMemcpy(dx_sum_host, d_dx, dx_n * sizeof(double), cudaMemcpyDeviceToHost));
  }
  MGA_CUDA(cudaStreamSynchronize(s_dn));
  return MGA_OK;
}

}  // extern "C"
