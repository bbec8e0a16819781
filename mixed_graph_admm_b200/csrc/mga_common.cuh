// Shared declarations of libmga: plan layout, error plumbing, block reductions.
// Everything here is internal; the public surface is include/mga.h.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdlib>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "mga.h"

namespace mga {

// ---------------------------------------------------------------------------------------------
// Device-side view of the graph tables (all pointers are device memory owned by the plan).
//
//  ELL tables (one row per node, -1 = no neighbour; quirk Q6: such a slot contributes exactly 0):
//    nbr_u (N, ku)  u_w (u_wT, N, ku)   — spatial Laplacian L_u        (ADMM.py:138-148)
//    nbr_d (N, kd)  d_w (d_wT, N, kd)   — temporal operator L_d        (ADMM.py:166-177)
//  CSR "in-list" for L_d^T: row c lists (src, slot) pairs with f[t,c] = sum d_w[t][slot] * v[t+1,src]
//    SCATTER (kNN) mode: all (i, i*kd+j) with nbr_d[i,j] == c, ascending slot = the order in which
//                        the reference's scatter_add visits them (ADMM.py:203-208)
//    GATHER mode:        row i's own forward list (ADMM.py:211-215)
//    csr_w (nnz) holds d_w[0][slot] and is used when d_wT == 1.
// ---------------------------------------------------------------------------------------------
struct GraphDev {
  int N, T, t_in;
  int ku, kd;
  int u_wT, d_wT;        // 1 = time-invariant
  int q1;                // 1: row t=0 of Ldr_T keeps the identity term (quirk Q1)
  int temporal;          // mga_temporal_kind
  int skip;              // BAND only
  int nnz, max_in_deg;
  const int* nbr_u;
  const float* u_w;
  const int* nbr_d;
  const float* d_w;
  const int* csr_ptr;
  const int* csr_src;
  const int* csr_slot;
  const float* csr_w;
  const float* band_w;   // (T, skip, N) BAND only
};

// Device view of the tables of the chunked streaming path (mga_stream2.cu): nodes renumbered in
// reverse-Cuthill-McKee order, time-invariant weights, vectors node-major v[b][n][4 * C4].
struct Graph2 {
  int N, T, t_in, C4;      // C4 = ceil(T / 4) chunks of 4 time steps per node row
  int CB, NB, tilesN, tilesC;   // CTA tile = NB nodes x CB chunks; tiles per window = tilesN * tilesC
  int NBt;                      // thread rows of a tile CTA: block = (CB, NBt), NBt >= NB, CB * NBt a multiple of 32
  int one3;                     // time-tiled kernels: the tile only fits once per SM (one CTA of 1024 threads)
  int db3;                      // time-tiled kernels double-buffered (one CTA per SM, next tile copied while this one is gathered)
  int CB3, NB3t, tiles3;        // time-tiled shared-memory kernels (k3_*): CB3 chunks x all nodes per CTA, block = (CB3, NB3t); CB3 = 0: off
  int kd, ku, q1;
  const int* perm;         // perm[internal] = caller's node id
  const int* nbr_d; const float* w_d;     // (N, kd) internal ids, -1 = no neighbour
  const int* nbr_u; const float* w_u;     // (N, ku)
  const int* in_ptr; const int* in_src; const float* in_w;   // in-list of L_d^T (CSR, entries in scatter order)
  const int* in_slot;                                        // in_w[e] == w_d[in_slot[e]]
  // time-tiled kernels: (row byte offset in the tile, weight bits) entries; the self link of the temporal graph is
  // in wself_d instead of the forward table / the in-list (in_self3 = 0: the in-list kept its self entries)
  int kd3, ku3, in_self3, in_ptr3_total;
  // node tiles: NT3 own nodes per tile (N: one tile), R3 = most rows any tile stages (own + external), in_max3 = most
  // in-list entries of a tile; per tile and table the external rows it references (ext*[extp*[j] .. extp*[j+1]))
  int NT3, ntile3, R3, in_max3;
  const int* extp_d; const int* ext_d; const int* extp_u; const int* ext_u; const int* extp_in; const int* ext_in;
  int smem3_d, smem3_u, smem3_in;
  const int2* tab_d; const int2* tab_u; const int2* tab_in3;
  const int* in_ptr3;
  const int* ord3;         // (N) row order of k3_ldrt_lhs: descending in-list length
  const int* ord4;         // (N) row order of k4_cg's phase C: ord3 dealt to the warps of a consumer group in snake order
  const float* wself_d;    // (N)
};

// consumer threads of the fused TMA kernels (mga_stream4.cuh); MGA_S4_NC overrides
inline void k4_env(int* cons) {
  int nc = 512;
  if (const char* e = std::getenv("MGA_S4_NC")) nc = std::atoi(e) >= 768 ? 768 : 512;
  *cons = nc;
}

struct Workspace {
  void* base = nullptr;
  size_t bytes = 0;
};

// Pipelined host entry point: the resident kernel waits for ready[c] == epoch before it touches chunk c of the batch and
// reports finished chunks through host_done[c] (mapped pinned host memory).  See mga_admm_solve_host.
struct HostPipe {
  const int* ready;
  int* done;
  int* host_done;      // device alias of the mapped host array
  int* abort_flag;
  int chunk, epoch;
  int chunk_up;        // windows per upload chunk (a multiple of `chunk`)
};

}  // namespace mga

struct mga_plan {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  int max_smem_sm = 0;         // shared memory of one SM (all resident CTAs together)
  size_t l2_bytes = 0;
  mga::GraphDev g{};
  std::vector<void*> owned;          // device allocations of the tables
  mga::Workspace ws;                 // grown on demand, reused between calls
  mga::Workspace ws_host_io;         // device staging of the *_host entry point
  void* pinned = nullptr;            // small pinned host block (flags, diag read-back)
  size_t pinned_bytes = 0;
  cudaStream_t io_streams[4] = {nullptr, nullptr, nullptr, nullptr};    // upload, run slot 0, run slot 1, download
  cudaEvent_t io_events[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool has_s2 = false;               // tables of the chunked streaming path
  mga::Graph2 g2{};
  std::map<std::pair<const void*, int64_t>, CUtensorMap> tmaps;   // TMA descriptors of workspace vectors, by (base, windows)
  int cg_mode = MGA_MODE_AUTO;       // mga_plan_set_cg_mode
  int res_slot = 0;                  // which half of the resident kernel's parking scratch the next launch uses
                                     // (the host entry point runs two chunk solves concurrently)
  const mga::HostPipe* pipe = nullptr;   // set by mga_admm_solve_host around its resident launch
  static constexpr int kPipeChunks = 256;
  int* pipe_dev = nullptr;           // device: ready[kPipeChunks], done[kPipeChunks], abort
  int* pipe_host = nullptr;          // mapped pinned host: host_done[kPipeChunks], then epoch words for the ready flags
  int* pipe_host_dev = nullptr;      // device alias of pipe_host
  int pipe_epoch = 0;
  // host copies of the tables
  std::vector<int> h_nbr_u, h_nbr_d, h_csr_ptr, h_csr_src;
  std::vector<float> h_u_w, h_d_w, h_csr_w;
  // resident-kernel schedule (mga_schedule.cpp), device copies
  bool has_sched = false;
  const int* r_perm = nullptr;
  const int* r_nbr_d = nullptr;
  const float* r_w_d = nullptr;
  const int* r_nbr_u = nullptr;
  const float* r_w_u = nullptr;
  const int* r_ell_ptr = nullptr;
  const int* r_ell_ent = nullptr;     // int2 pairs (node, weight bits)
  int r_ell_total = 0;
  const float* band_uniform = nullptr;   // (T, skip) band weights when they are the same for every node
  bool ldrt_gather = false;           // MGA_LDRT_GATHER: "L_d^T" is not the transpose of L_d (use_kNN=False, ADMM.py:211-215)
  int r_kd = 0, r_ku = 0;             // slots per row of the scheduled tables (self links and pads dropped)
  const float* r_w_self = nullptr;    // (N) self-link weights, internal order
  // work counters of the resident kernel's dynamic window hand-out: a ring, one per launch in flight
  static constexpr int kCounters = 16;
  int* r_counters = nullptr;          // kCounters x 32 ints (128 B apart)
  unsigned r_counter_next = 0;
};

namespace mga {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
extern std::atomic<int64_t> g_launches;

#define MGA_CUDA(expr)                                         \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::mga::cuda_fail(_e, #expr); \
  } while (0)

#define MGA_LAUNCH_CHECK(name)                                  \
  do {                                                          \
    ::mga::g_launches.fetch_add(1, std::memory_order_relaxed);  \
    cudaError_t _e = cudaGetLastError();                        \
    if (_e != cudaSuccess) return ::mga::cuda_fail(_e, name);   \
  } while (0)

int ensure_workspace(mga_plan* plan, Workspace& ws, size_t bytes);

template <typename S> struct DType;
template <> struct DType<float> { static constexpr int id = MGA_F32; };
template <> struct DType<double> { static constexpr int id = MGA_F64; };

// ---- reductions -------------------------------------------------------------------------------
template <typename S>
__device__ __forceinline__ S warp_sum(S v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over the block; result valid in every thread.  `red` must hold >= 32 S; callers alternate
// between two buffers (or sync) so that back-to-back reductions do not race.
template <typename S>
__device__ __forceinline__ S block_sum(S v, S* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  if (lane == 0) red[w] = v;
  __syncthreads();
  S t = 0;
  for (int k = 0; k < nw; ++k) t += red[k];   // fixed order: every thread gets the same bits
  return t;
}

// streaming-mode entry points implemented in mga_stream.cu
int stream_apply(mga_plan*, int op, const mga_params*, const void* x, void* y, const void* mask, int64_t B,
                 int dtype, cudaStream_t st);
int stream_cg(mga_plan*, int system, const mga_params*, const void* rhs, void* x, const void* mask_first,
              int64_t B, int dtype, int max_iter, double tol, int32_t* iters_out, void* alpha, void* beta,
              cudaStream_t st);
int stream_admm(mga_plan*, const mga_params*, const void* y, int y_rows, const void* mask, void* x_out, int64_t B,
                int dtype, int n_outer, int max_cg, double cg_tol, double admm_tol, double t_mean, double t_var,
                int want_diag, const mga_admm_outputs* outs, cudaStream_t st);
// chunked streaming path (mga_stream2.cu): fp32, forecasting mode, fixed iteration counts, ablation None
bool stream2_eligible(const mga_plan*, int dtype);
void stream2_threads3(Graph2* g);   // block shape of the time-tiled kernels after CB3 / one3 are final
void stream2_tiling(Graph2* g);     // fills CB / NB / tilesN / tilesC for the kernels' block size
int stream2_cg(mga_plan*, int system, const mga_params*, const void* rhs, void* x, int64_t B, int n_cg, void* alpha,
               void* beta, cudaStream_t st);
int stream2_admm(mga_plan*, const mga_params*, const void* y, void* x_out, int64_t B, int n_outer, int max_cg, double t_mean,
                 double t_var, int want_diag, const mga_admm_outputs* outs, cudaStream_t st);
// cluster mode (mga_cluster.cu): one thread-block cluster per window, any dtype, stop tests on the device (B = 1)
bool cluster_eligible(const mga_plan*, int dtype);
int cluster_admm(mga_plan*, const mga_params*, const void* y, const void* mask, void* x_out, int64_t B, int dtype, int n_outer, int max_cg,
                 double cg_tol, double admm_tol, double t_mean, double t_var, int want_diag, const mga_admm_outputs* outs,
                 cudaStream_t st);
// resident mode (mga_resident.cu)
bool resident_eligible(const mga_plan*, int dtype);
int resident_smem_bytes(const mga_plan*, int* threads);
int resident_cg(mga_plan*, int system, const mga_params*, const void* rhs, void* x, int64_t B, int n_cg, void* alpha,
                void* beta, cudaStream_t st);
int resident_admm(mga_plan*, const mga_params*, const void* y, const void* mask, void* x_out, int64_t B, int n_outer, int n_cg,
                  double t_mean, double t_var, int want_diag, const mga_admm_outputs* outs, cudaStream_t st);

}  // namespace mga
