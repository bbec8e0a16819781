// Resident mode: the whole of combined_loop (ADMM.py:528-648) for one window inside one CTA.
//
// Kernels of this file: k_admm_resident (the whole loop; instantiations for forecasting, mask /
// interpolation mode and the banded skip-connection line graph) and k_cg_resident (CG_solver alone).
//
// Work mapping.  The T time steps of a node are cut into chunks of 4; a thread owns CH
// consecutive chunks of ONE node (thread = (node i, slab s), slab-major so a warp is uniform in
// s).  The CG vectors (x, r, p, Ap) of the thread's 4*CH lattice points live in registers.
//
// Shared-memory layout.  Only the vectors other threads gather from are staged: pbuf (p) and
// qbuf (a shifted copy of q = L_d p), NODE-major: buf[node * TP + t], TP = 4 * odd.  A
// neighbour's whole time series is then contiguous, so every gather is an aligned 128-bit load
// (4 time steps per LDS.128), and `TP/4 odd` makes node -> 16-byte bank group a bijection mod 8,
// which keeps the 8 lanes of a quarter-warp on distinct groups when their nodes differ mod 8.
// Rows N .. N+7 are zero rows (one per bank group) that absorb the "-1 = no neighbour" entries (quirk Q6)
// and the padding of the tables.
//
// The shift.  L_d reads p at t-1 and L_d^T reads q at t+1 (ADMM.py:171, 200-208), which would
// misalign the 4-step chunks.  Instead each thread computes qs[k] = q[k+1] for the k it owns
// (neighbour chunks of p at the SAME k, aligned) and qbuf holds qs; the in-list gather of
// L_d^T then needs qs at the thread's own k (aligned again).  Only the thread's own node is
// touched off-chunk: p[k+1] and qs[k-1] across the slab edge, one scalar load each.
//
// Tables.  The thread's rows of the ELL tables (time-invariant) sit in shared memory slot-major, one
// conflict-free 64-bit load per neighbour, each entry = (ABSOLUTE shared address of the neighbour's
// pbuf row at this thread's slab, weight) — so a gather is LDS.64 -> 3 x LDS.128 with immediate
// offsets and no integer arithmetic (MGA_RES_TAB_SMEM=0 keeps them in registers instead: 26 more
// registers, one CTA per SM, measured slower).  The node's own link is not in the tables: its weight
// multiplies the thread's own registers.  The in-list (the transpose that replaces the reference's
// scatter_add) sits in shared memory as a per-warp, step-major ELL of (row address, weight) pairs.
// Node numbering, the visit order of a row's neighbours and of its in-list are chosen at plan time
// (mga_schedule.cpp) so that warps have in-lists of similar length and quarter-warps hit distinct
// bank groups.
//
// Registers are the scarce resource (12 lattice points per thread x {x, r, p, Ap, accumulators}):
// the seven ADMM state vectors are parked between uses — in shared memory when that does not cost
// residency, else in a per-CTA L2-resident scratch — and re-read where needed instead of being
// kept live across a solve.  HBM traffic per window is y in, x out (+ optional iterates /
// diagnostics).
//
// A persistent grid (CTAs/SM x SMs) takes windows from an atomic counter; CTAs never exchange
// data, so windows shard over CTAs — and over GPUs — with no collective.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "mga_common.cuh"

#ifndef MGA_RES_TAB_SMEM
#define MGA_RES_TAB_SMEM 1
#endif
#ifndef MGA_RES_LDCG      // 0: plain (L1-cached) loads of y (experiments; the host entry point needs ld.global.cg)
#define MGA_RES_LDCG 1
#endif
#ifndef MGA_RES_MINB      // CTAs per SM the <= 320-thread instantiations are compiled for
#define MGA_RES_MINB 2
#endif

namespace mga {

enum { ST_X = 0, ST_ZU, ST_ZD, ST_GU, ST_GD, ST_GAM, ST_PHI, ST_COUNT };

struct ResArgs {
  int N, T, t_in, n_outer, n_cg, q1, want_diag, state_in_smem, nnz;
  int NT;      // threads per slab (N rounded up to 32)
  int S;       // slabs (threads per node)
  int TP;      // row stride of the node-major buffers, floats (4 * odd, >= S * 4 * CH)
  int64_t B;
  int kd, ku;
  // scheduled tables of the plan (mga_schedule.cpp): internal node order, conflict-aware slot order
  const int* perm;            // perm[internal] = original node (global memory is in original order)
  const int* nbr_d; const float* d_w; const int* nbr_u; const float* u_w;   // (N, kd) / (N, ku), N = zero row
  const float* w_self;        // (N) weight of each node's own link in the temporal table (not in nbr_d / the in-list)
  const int* ell_ptr;         // (NT/32 + 1) first in-list step of each 32-row warp
  const int2* ell_ent;        // (ell_total) step-major: (internal src or N, weight bits) per lane
  int ell_total;
  const float* y; float* x_out;
  const float* band_w;        // banded line graph: (T, skip, N) weights (caller's node order), else NULL
  const float* band_uniform;  // (T, skip) when the weights do not depend on the node (as the reference builds them)
  int transpose_exact;        // 1: the in-list is the transpose of the forward temporal table (kNN scatter mode, line graph)
  int skip, band_floats;      // band_floats: shared-memory floats reserved for the staged (T, skip) table (multiple of 4)
  const float* mask;          // mask / interpolation mode (ADMM.py:373-376, 783-811): (B, T, N), y then has T rows
  float* out[ST_COUNT];       // optional per-window outputs (index by ST_*; ST_X unused)
  float* scratch;             // gridDim.x * ST_COUNT * N * TP floats when !state_in_smem
  int* next_window;           // zeroed per launch: windows beyond the first gridDim.x are handed out dynamically
  double* diag; double* dx_sum;
  float* alpha; float* beta;
  float rho, rho_u, rho_d, thr;
  float ax, cx, azu, czu, azd, czd;
  float t_mean, t_var;
  // Host-buffer entry point (mga_admm_solve_host): ONE persistent launch consumes the batch while it is still being
  // uploaded and hands finished chunks to the download stream.  Chunk c = windows [c * chunk, (c + 1) * chunk).
  const int* ready;           // device: ready[c] == epoch once the chunk's y (and the flag, in stream order) has landed; NULL: y is there
  int* done;                  // device: finished windows per chunk (left at 0 again by the CTA that completes the chunk)
  int* host_done;             // mapped pinned host memory: host_done[c] = epoch when every x of the chunk is in device memory
  int* abort_flag;            // device: set when a CTA gave up waiting (upload never arrived); the host reports an error
  int chunk, epoch;
  int chunk_up;               // windows per UPLOAD chunk (ready flags); `chunk` is the download granularity (done counters)
};

template <bool CG>
__device__ __forceinline__ float ldy(const float* p) {
#if MGA_RES_LDCG
  if (CG) return __ldcg(p);
#endif
  return *p;
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// The chunk hand-shake of the pipelined host launch, called by one thread between two windows.  Kept out of line on
// purpose: inlined, its 64-bit divisions and spin loops changed the register allocation of the whole kernel (4 % slower).
// pipe_acquire: wait until the chunk of window b has landed; returns 1 when it gave up (an upload never arrived).
static __device__ __noinline__ int pipe_acquire(const int* ready, int chunk, int epoch, long long b, int* abort_flag) {
  const int* flag = ready + (int)(b / chunk);
  if (ld_acquire_gpu(flag) == epoch) return 0;
  const unsigned long long t_start = globaltimer_ns();
  while (ld_acquire_gpu(flag) != epoch) {
    __nanosleep(256);
    if (globaltimer_ns() - t_start > 4000000000ull) { atomicExch(abort_flag, 1); return 1; }   // 4 s: never hang the device
  }
  return 0;
}
// pipe_release: count window b as finished; the thread that completes a chunk publishes it to the host
static __device__ __noinline__ void pipe_release(int* done, int* host_done, int chunk, int epoch, long long b, long long B) {
  const int c = (int)(b / chunk);
  const long long left = B - (long long)c * chunk;
  const int nwin = (int)(left < chunk ? left : chunk);
  if (atomicAdd(done + c, 1) + 1 == nwin) {
    done[c] = 0;
    __threadfence_system();
    *reinterpret_cast<volatile int*>(host_done + c) = epoch;
  }
}

// Shared-memory loads by 32-bit shared address.  The tables hold ABSOLUTE shared addresses (base of
// pbuf / qbuf folded in when the CTA fills them), so a gather is "LDS.128 [entry + 16 c]" with no
// address arithmetic between the table load and the data load.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ int2 lds64(uint32_t a) {
  int2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(a), "n"(OFF) : "memory");
  return v;
}

template <int CH, int K>
struct Ctx {
  static constexpr int TS = 4 * CH;
  int t0, T, t_in;
  bool has_next, has_prev;   // another slab of this node owns t0+TS / t0-1
  int own;                   // i * TP + t0
#if MGA_RES_TAB_SMEM
  uint32_t tabd;             // shared address of this thread's temporal-table column: K entries, stride NTB bytes:
  uint32_t tabu;             //   (shared address of pbuf[nbr][t0], weight bits); tabu: spatial table
  int NTB;
#else
  uint32_t nd[K];            // shared address of pbuf[nbr][t0] (a zero row for padding)
  float wd[K];
  uint32_t nu[K];
  float wu[K];
#endif
  float wself;               // weight of the node's own link in the temporal table (left out of the tables)
  bool qdot;                 // the in-list is the exact transpose of the forward table: <v, L_d^T L_d v> = ||L_d v||^2
  int steps;                 // in-list steps of this thread's warp (padded with zero-weight entries)
  float* pbuf;
  float* qbuf;
  uint32_t ent;              // shared address of this lane's first in-list entry; stride 256 B:
                             //   (shared address of qbuf[src][0], weight bits)
  uint32_t t0b;              // t0 * 4
  float* red;                // 2 x 32
  int red_sel;
  // banded line graph (use_line_graph with skip_connection > 1, ADMM.py:41-52): L_d couples a node only with
  // its own past, weights (T, skip, N); no neighbour gathers, the stencil reads the thread's own staged row
  const float* bw;           // band weights + this thread's node in the caller's numbering
  int skip, bN;              // stencil length; stride between consecutive (t, s) weights

  // Block sum, result in every thread.  Warp partials go to a 32-float row (zero beyond the warp count);
  // after the barrier every thread reads the whole row as broadcast 128-bit loads and adds it up in one
  // fixed tree — shorter than a second shuffle butterfly, and bit-identical in every thread.
  __device__ __forceinline__ float bsum(float v) {
    float* r = red + 32 * red_sel;
    red_sel ^= 1;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum<float>(v);
    if (lane == 0) r[w] = v;
    __syncthreads();
    const int nw4 = (blockDim.x + 127) >> 7;       // float4 groups that hold partials
    float4 t = *reinterpret_cast<const float4*>(r);
    float s = (t.x + t.y) + (t.z + t.w);
    for (int k = 1; k < nw4; ++k) {
      t = *reinterpret_cast<const float4*>(r + 4 * k);
      s += (t.x + t.y) + (t.z + t.w);
    }
    return s;
  }

  // The same reduction in two halves around a barrier the caller has anyway: bsum_post() before it, bsum_read() after.
  __device__ __forceinline__ const float* bsum_post(float v) {
    float* r = red + 32 * red_sel;
    red_sel ^= 1;
    v = warp_sum<float>(v);
    if ((threadIdx.x & 31) == 0) r[threadIdx.x >> 5] = v;
    return r;
  }
  __device__ __forceinline__ float bsum_read(const float* r) const {
    const int nw4 = (blockDim.x + 127) >> 7;
    float4 t = *reinterpret_cast<const float4*>(r);
    float s = (t.x + t.y) + (t.z + t.w);
    for (int k = 1; k < nw4; ++k) {
      t = *reinterpret_cast<const float4*>(r + 4 * k);
      s += (t.x + t.y) + (t.z + t.w);
    }
    return s;
  }

  __device__ __forceinline__ void put(float* buf, const float (&v)[TS]) const {
#pragma unroll
    for (int c = 0; c < CH; ++c)
      *reinterpret_cast<float4*>(buf + own + 4 * c) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
  }
  __device__ __forceinline__ void get(const float* buf, float (&v)[TS]) const {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const float4 g = *reinterpret_cast<const float4*>(buf + own + 4 * c);
      v[4 * c] = g.x; v[4 * c + 1] = g.y; v[4 * c + 2] = g.z; v[4 * c + 3] = g.w;
    }
  }

  // acc[k] += w * row[k], k = 0..TS-1, as CH aligned 128-bit loads from the shared address `row`
  __device__ __forceinline__ void gather_acc(uint32_t row, float w, float (&acc)[TS]) const {
    float4 g[CH];
    g[0] = lds128<0>(row);
    if (CH > 1) g[CH > 1 ? 1 : 0] = lds128<16>(row);
    if (CH > 2) g[CH > 2 ? 2 : 0] = lds128<32>(row);
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      acc[4 * c] += w * g[c].x;
      acc[4 * c + 1] += w * g[c].y;
      acc[4 * c + 2] += w * g[c].z;
      acc[4 * c + 3] += w * g[c].w;
    }
  }

  // acc += sum_j w_j * pbuf[nbr_j] over the temporal / spatial forward table
  __device__ __forceinline__ void fwd_d(float (&acc)[TS]) const {
#if MGA_RES_TAB_SMEM
    uint32_t t = tabd;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int2 en = lds64<0>(t);
      t += NTB;
      gather_acc((uint32_t)en.x, __int_as_float(en.y), acc);
    }
#else
#pragma unroll
    for (int j = 0; j < K; ++j) gather_acc(nd[j], wd[j], acc);
#endif
  }
  __device__ __forceinline__ void fwd_u(float (&acc)[TS]) const {
#if MGA_RES_TAB_SMEM
    uint32_t t = tabu;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int2 en = lds64<0>(t);
      t += NTB;
      gather_acc((uint32_t)en.x, __int_as_float(en.y), acc);
    }
#else
#pragma unroll
    for (int j = 0; j < K; ++j) gather_acc(nu[j], wu[j], acc);
#endif
  }

  // qs[k] = q[t0+k+1] where q = L_d v (ADMM.py:166-177), v already in pbuf (synced).
  __device__ __forceinline__ void shifted_ldr(const float (&v)[TS], float (&qs)[TS]) const {
#pragma unroll
    for (int k = 0; k < TS; ++k) qs[k] = wself * v[k];       // the self link: p_i is in registers
    fwd_d(qs);
    const float vnext = has_next ? pbuf[own + TS] : 0.f;
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      const float up = (k < TS - 1) ? v[k + 1] : vnext;
      qs[k] = (t0 + k + 1 < T) ? up - qs[k] : 0.f;
    }
  }

  // q[k] = (L_d v)[t0 + k] for the banded graph (ADMM.py:158-164): q[t] = v[t] - sum_s w[t][s] v[t-1-s].
  // v is in registers AND staged in `buf` (own row: the columns before t0 belong to this node's earlier slabs).
  // A register window slides one step into the past per stencil tap, so each tap costs one scalar load.
  __device__ __forceinline__ void band_ldr(const float (&v)[TS], const float* buf, float (&q)[TS]) const {
    float win[TS], acc[TS];
#pragma unroll
    for (int k = 0; k < TS; ++k) { win[k] = k > 0 ? v[k - 1] : (t0 > 0 ? buf[own - 1] : 0.f); acc[k] = 0.f; }
    for (int s = 0; s < skip; ++s) {
#pragma unroll
      for (int k = 0; k < TS; ++k) {
        const int t = t0 + k;
        if (t < T && t - 1 - s >= 0) acc[k] += bw[(t * skip + s) * bN] * win[k];
      }
#pragma unroll
      for (int k = TS - 1; k > 0; --k) win[k] = win[k - 1];
      win[0] = (t0 - 2 - s >= 0) ? buf[own - 2 - s] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      const int t = t0 + k;
      q[k] = (t >= 1 && t < T) ? v[k] - acc[k] : 0.f;
    }
  }
  // l[k] = (L_d^T v)[t0 + k] with the row rules of ADMM.py:187-194 (row T-1 keeps v, row 0 has no identity
  // term): l[t] = v[t] - sum_s w[t+1+s][s] v[t+1+s]; the window slides into the future
  __device__ __forceinline__ void band_ldrt(const float (&v)[TS], const float* buf, float (&l)[TS]) const {
    float win[TS], f[TS];
#pragma unroll
    for (int k = 0; k < TS; ++k) { win[k] = k < TS - 1 ? v[k + 1] : (t0 + TS < T ? buf[own + TS] : 0.f); f[k] = 0.f; }
    for (int s = 0; s < skip; ++s) {
#pragma unroll
      for (int k = 0; k < TS; ++k) {
        const int ts = t0 + k + 1 + s;
        if (ts < T) f[k] += bw[(ts * skip + s) * bN] * win[k];
      }
#pragma unroll
      for (int k = 0; k < TS - 1; ++k) win[k] = win[k + 1];
      win[TS - 1] = (t0 + TS + 1 + s < T) ? buf[own + TS + 1 + s] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      const int t = t0 + k;
      l[k] = t >= T ? 0.f : (t == T - 1 ? v[k] : (t == 0 ? -f[k] : v[k] - f[k]));
    }
  }

  // f[k] = sum over the in-list of w * qbuf[src][t0+k]   (ADMM.py:200-209 as a gather; qbuf holds the
  // vector shifted by one time step, so this is the "father" sum at t0+k).  Two entries per trip, the
  // next two (row address, weight) pairs are fetched a trip ahead; the table has two spare steps at
  // its end so the look-ahead never leaves it.
  // ZERO = false: accumulate on top of what f already holds
  template <bool ZERO = true>
  __device__ __forceinline__ void father_sum(float (&f)[TS]) const {
    if (ZERO) {
#pragma unroll
      for (int k = 0; k < TS; ++k) f[k] = 0.f;
    }
    uint32_t ep = ent;
    int2 a = lds64<0>(ep), b = lds64<256>(ep);
    int n = steps;
    for (; n >= 2; n -= 2) {
      const int2 a2 = lds64<512>(ep), b2 = lds64<768>(ep);
      ep += 512;
      gather_acc((uint32_t)a.x + t0b, __int_as_float(a.y), f);
      gather_acc((uint32_t)b.x + t0b, __int_as_float(b.y), f);
      a = a2; b = b2;
    }
    if (n) gather_acc((uint32_t)a.x + t0b, __int_as_float(a.y), f);
  }

  // out = A v for the x / z_d systems: diag(v) + c * L_d^T L_d v  (ADMM.py:371-387, 392-394)
  // MASKED: H = the caller's elementwise mask m (LHS_x(x, mask), ADMM.py:375-376) instead of "rows t < t_in"
  // DOT: also return <v, A v>.  A = D + c L_d^T L_d, so <v, A v> = sum D v^2 + c ||L_d v||^2: every term is known
  // once q = L_d v is (before the in-list gather), the partial sums ride on the barrier that publishes qs, and the
  // reduction's latency hides behind the gather — one barrier fewer per CG iteration than reducing <p, Ap> afterwards.
  // (Same value up to rounding; a sum of non-negative terms instead of one with cancellation.)
  template <bool XSYS, bool MASKED, bool DOT>
  __device__ __forceinline__ void apply_cldr(const float (&v)[TS], float (&out)[TS], float a, float c, const float (&m)[TS],
                                             float& dot) {
    put(pbuf, v);
    __syncthreads();
    const float* part = nullptr;
    {
      // The thread's own q terms enter the accumulator while qs is still in registers: with q[k] = qs[k-1],
      // (L_d^T q)[k] = q[k] - wself qs[k] - f[k]; out starts as wself qs[k] - qs[k-1] and the in-list gather adds
      // f on top, so nothing of qs has to survive (or be re-read after) the gather.
      float qs[TS];
      shifted_ldr(v, qs);
      put(qbuf, qs);
#pragma unroll
      for (int k = 0; k < TS; ++k) out[k] = wself * qs[k] - (k > 0 ? qs[k - 1] : 0.f);
      if (DOT && qdot) {
        float dv = 0.f, dq = 0.f;
#pragma unroll
        for (int k = 0; k < TS; ++k) {
          dv += ((XSYS && t0 + k < t_in) ? v[k] * v[k] : 0.f) + a * (v[k] * v[k]);
          dq += qs[k] * qs[k];                               // the qs of all threads are exactly the q[t >= 1]; q[0] = 0
        }
        part = bsum_post(dv + c * dq);
      }
    }
    __syncthreads();
    if (has_prev) out[0] -= qbuf[own - 1];                   // q[t0] of a later slab; q[0] = 0 (ADMM.py:176)
    father_sum<false>(out);                                  // out = -(L_d^T L_d v); row T-1: f = 0 because qs[T-1] = 0
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      const float l = -out[k];                               // Q1 is moot here as q[0] = 0
      if (XSYS) out[k] = ((MASKED ? v[k] * m[k] : (t0 + k < t_in ? v[k] : 0.f)) + a * v[k]) + c * l;
      else out[k] = c * l + a * v[k];
    }
    if (DOT) {
      if (qdot) {
        dot = bsum_read(part);
      } else {                      // use_kNN=False: "L_d^T" is a gather with the forward table (ADMM.py:211-215), not a transpose
        float loc = 0.f;
#pragma unroll
        for (int k = 0; k < TS; ++k) loc += v[k] * out[k];
        dot = bsum(loc);
      }
    }
  }

  // the same system matrices on the banded line graph
  template <bool XSYS, bool MASKED>
  __device__ __forceinline__ void apply_cldr_band(const float (&v)[TS], float (&out)[TS], float a, float c, const float (&m)[TS]) {
    put(pbuf, v);
    __syncthreads();
    float q[TS];
    band_ldr(v, pbuf, q);
    put(qbuf, q);
    __syncthreads();
    band_ldrt(q, qbuf, out);
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      if (XSYS) out[k] = ((MASKED ? v[k] * m[k] : (t0 + k < t_in ? v[k] : 0.f)) + a * v[k]) + c * out[k];
      else out[k] = c * out[k] + a * v[k];
    }
  }

  // out = mu_u L_u v + (rho_u/2) v  (ADMM.py:389-390)
  __device__ __forceinline__ void apply_lu(const float (&v)[TS], float (&out)[TS], float a, float c) {
    put(pbuf, v);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TS; ++k) out[k] = 0.f;
    fwd_u(out);
#pragma unroll
    for (int k = 0; k < TS; ++k) out[k] = c * (v[k] - out[k]) + a * v[k];
  }

  // out = A v; DOT: dot = <v, A v> as well
  template <int SYS, bool MASKED, bool BAND, bool DOT>
  __device__ __forceinline__ void apply(const float (&v)[TS], float (&out)[TS], float a, float c, const float (&m)[TS], float& dot) {
    if (SYS != MGA_SYS_ZU && !BAND) {
      if (SYS == MGA_SYS_X) apply_cldr<true, MASKED, DOT>(v, out, a, c, m, dot);
      else apply_cldr<false, false, DOT>(v, out, a, c, m, dot);
      return;
    }
    if (SYS == MGA_SYS_ZU) apply_lu(v, out, a, c);
    else if (SYS == MGA_SYS_X) apply_cldr_band<true, MASKED>(v, out, a, c, m);
    else apply_cldr_band<false, false>(v, out, a, c, m);
    if (DOT) {                      // L_u is not symmetric (quirk Q5) and the banded stencil keeps the plain form: <v, A v> as is
      float loc = 0.f;
#pragma unroll
      for (int k = 0; k < TS; ++k) loc += v[k] * out[k];
      dot = bsum(loc);
    }
  }

  // CG_solver, fixed iteration count (ADMM.py:329-368 with an unreachable tolerance; the
  // arithmetic is unguarded like the reference's, quirk Q10).  r holds the right-hand side on
  // entry, x the warm start; x holds the solution on exit.
  // MASK0: the mask goes to the initial residual only, the iterations use H = "rows t < t_in" (quirk Q4, ADMM.py:344-349)
  template <int SYS, bool MASK0, bool BAND = false>
  __device__ __forceinline__ void cg(float (&x)[TS], float (&r)[TS], float a, float c, int n_cg, float* alpha_out,
                                     float* beta_out, int64_t B, const float (&m0)[TS]) {
    float p[TS], ap[TS];
    float pap = 0.f;
    apply<SYS, MASK0, BAND, false>(x, ap, a, c, m0, pap);
    float loc = 0.f;
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      r[k] = r[k] - ap[k];
      p[k] = r[k];
      loc += r[k] * r[k];
    }
    float rr = bsum(loc);
    for (int it = 0; it < n_cg; ++it) {
#if defined(MGA_RES_X) && (MGA_RES_X & 2)      // timing experiment: no <p, Ap> reduction
      apply<SYS, false, BAND, false>(p, ap, a, c, m0, pap);
      pap = rr * 2.f + ap[0] * 1e-30f;
#else
      apply<SYS, false, BAND, true>(p, ap, a, c, m0, pap);
#endif
      const float alpha = rr / pap;
      loc = 0.f;
#pragma unroll
      for (int k = 0; k < TS; ++k) {
        x[k] = x[k] + alpha * p[k];
        r[k] = r[k] - alpha * ap[k];
        loc += r[k] * r[k];
      }
#if defined(MGA_RES_X) && (MGA_RES_X & 1)      // timing experiment: no <r, r> reduction (and its barrier)
      const float rrn = rr * 0.5f + loc * 1e-30f;
#else
      const float rrn = bsum(loc);
#endif
      const float beta = rrn / rr;
      rr = rrn;
      if (alpha_out && threadIdx.x == 0) {
        alpha_out[(size_t)it * B] = alpha;
        beta_out[(size_t)it * B] = beta;
      }
#pragma unroll
      for (int k = 0; k < TS; ++k) p[k] = r[k] + beta * p[k];
    }
  }
};

__device__ __forceinline__ float soft_thr(float s, float d) {
  const float u = fabsf(s) - d;
  const float sg = (float)((s > 0.f) - (s < 0.f));
  return sg * u * (float)(u > 0.f);   // ADMM.py:407-408
}

// Per-CTA setup shared by the kernels of this file: carve shared memory, fill the tables with absolute
// shared addresses, clear the staging buffers.  Ends with a barrier.
template <int CH, int K>
struct Cta {
  Ctx<CH, K> c;
  int i, t0, orig;
  bool active;
  float* pbuf; float* qbuf; float* red; float* dred; float* st_smem;
  float* band_s;             // (T, skip) band weights staged in shared memory (banded line graph with node-invariant weights)

  __device__ __forceinline__ void init(const ResArgs& a, unsigned char* smem_raw) {
    constexpr int TS = 4 * CH;
    const int N = a.N, TP = a.TP;
    const int rows = (N + 8) * TP;
    pbuf = reinterpret_cast<float*>(smem_raw);
    qbuf = pbuf + rows;
    red = qbuf + rows;                                // 64 floats + the next-window slot (68 with padding)
    dred = red + 68;                                  // MGA_DIAG_COLS x 32 floats
    int2* ent = reinterpret_cast<int2*>(dred + MGA_DIAG_COLS * 32);
    float* nxt = reinterpret_cast<float*>(ent + a.ell_total + 64);      // + 2 spare steps for the look-ahead
#if MGA_RES_TAB_SMEM
    int2* tabd = reinterpret_cast<int2*>(nxt);        // K x (S * NT), slot-major, one column per thread
    int2* tabu = tabd + K * a.S * a.NT;               // K x (S * NT)
    nxt = reinterpret_cast<float*>(tabu + K * a.S * a.NT);
#endif
    band_s = nxt;
    nxt += a.band_floats;
    st_smem = nxt;
    const int s = threadIdx.x / a.NT;
    i = threadIdx.x - s * a.NT;
    active = i < N;
    t0 = s * TS;

    c.t0 = t0; c.T = a.T; c.t_in = a.t_in;
    c.has_next = (s + 1 < a.S);
    c.has_prev = (s > 0);
    c.own = (active ? i : N) * TP + t0;     // inactive lanes park on the zero row (they only ever write zeros)
    c.pbuf = pbuf; c.qbuf = qbuf; c.red = red; c.red_sel = 0;
    const uint32_t pb = smem_addr(pbuf), qb = smem_addr(qbuf);
    c.t0b = (uint32_t)t0 * 4u;
#if MGA_RES_TAB_SMEM
    const int ncol = a.S * a.NT;
    c.NTB = ncol * 8; c.tabd = smem_addr(tabd + threadIdx.x); c.tabu = smem_addr(tabu + threadIdx.x);
#endif
#pragma unroll
    for (int j = 0; j < K; ++j) {
      int nb = N;
      float w = 0.f;
      if (active && j < a.kd) { nb = a.nbr_d[i * a.kd + j]; w = a.d_w[i * a.kd + j]; }
#if MGA_RES_TAB_SMEM
      tabd[j * ncol + threadIdx.x] = make_int2((int)(pb + (uint32_t)(nb * TP + t0) * 4u), __float_as_int(w));
#else
      c.nd[j] = pb + (uint32_t)(nb * TP + t0) * 4u;
      c.wd[j] = w;
#endif
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
      int nb = N;
      float w = 0.f;
      if (active && j < a.ku) { nb = a.nbr_u[i * a.ku + j]; w = a.u_w[i * a.ku + j]; }
#if MGA_RES_TAB_SMEM
      tabu[j * ncol + threadIdx.x] = make_int2((int)(pb + (uint32_t)(nb * TP + t0) * 4u), __float_as_int(w));
#else
      c.nu[j] = pb + (uint32_t)(nb * TP + t0) * 4u;
      c.wu[j] = w;
#endif
    }
    c.wself = active ? a.w_self[i] : 0.f;
    c.qdot = a.transpose_exact != 0;
    {
      const int wn = i >> 5;     // warp of this node row (the same for every slab)
      const int first = a.ell_ptr[wn];
      c.steps = a.ell_ptr[wn + 1] - first;
#if defined(MGA_RES_X) && (MGA_RES_X & 12)      // timing experiment: every warp's in-list cut to 7 steps (4) / set to 6 steps (8)
      c.steps = (MGA_RES_X & 8) ? 6 : min(c.steps, 7);
#endif
      c.ent = smem_addr(ent + (size_t)first * 32 + (i & 31));
    }
    for (int e = threadIdx.x; e < a.ell_total + 64; e += blockDim.x) {
      int2 en = make_int2(N, 0);                 // spare steps: a zero row, zero weight
      if (e < a.ell_total) en = a.ell_ent[e];
      ent[e] = make_int2((int)(qb + (uint32_t)(en.x * TP) * 4u), en.y);
    }
    orig = active ? a.perm[i] : 0;   // this thread's node in the caller's numbering
    for (int k = threadIdx.x; k < 2 * rows; k += blockDim.x) pbuf[k] = 0.f;
    for (int k = threadIdx.x; k < 64; k += blockDim.x) red[k] = 0.f;     // bsum() relies on zeros beyond the warp count
    if (a.band_uniform) for (int k = threadIdx.x; k < a.T * a.skip; k += blockDim.x) band_s[k] = a.band_uniform[k];
    __syncthreads();
  }
};

// CG_solver (ADMM.py:329-368) on its own, fixed iteration count, one window per CTA: rhs and the warm
// start come from HBM once, every iteration runs out of registers and shared memory, x goes back once
// (12 B per lattice point and SOLVE instead of 48 B per point and ITERATION).
struct CgArgs {
  int system, n_cg;
  const float* rhs;
  float* x;
  float* alpha; float* beta;   // (n_cg, B) or NULL
  float a, c;                  // diagonal and operator coefficients of the system
};

template <int CH, int K, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k_cg_resident(const ResArgs a, const CgArgs g) {
  constexpr int TS = 4 * CH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta<CH, K> cta;
  cta.init(a, smem_raw);
  Ctx<CH, K>& c = cta.c;
  const int N = a.N, T = a.T, t0 = cta.t0;
  int& s_next = *reinterpret_cast<int*>(cta.red + 64);
  for (int64_t b = blockIdx.x; b < a.B;) {
    float x[TS], r[TS];
#pragma unroll
    for (int k = 0; k < TS; ++k) {
      const bool ok = cta.active && t0 + k < T;
      const size_t at = ((size_t)b * T + t0 + k) * N + cta.orig;
      x[k] = ok ? g.x[at] : 0.f;
      r[k] = ok ? g.rhs[at] : 0.f;
    }
    float* al = g.alpha ? g.alpha + b : nullptr;
    float* be = g.beta ? g.beta + b : nullptr;
    if (g.system == MGA_SYS_X) c.template cg<MGA_SYS_X, false>(x, r, g.a, g.c, g.n_cg, al, be, a.B, r);
    else if (g.system == MGA_SYS_ZU) c.template cg<MGA_SYS_ZU, false>(x, r, g.a, g.c, g.n_cg, al, be, a.B, r);
    else c.template cg<MGA_SYS_ZD, false>(x, r, g.a, g.c, g.n_cg, al, be, a.B, r);
    if (cta.active) {
#pragma unroll
      for (int k = 0; k < TS; ++k)
        if (t0 + k < T) g.x[((size_t)b * T + t0 + k) * N + cta.orig] = x[k];
    }
    if (threadIdx.x == 0) s_next = atomicAdd(a.next_window, 1);
    __syncthreads();
    b = (int64_t)gridDim.x + s_next;
  }
}

// VAR: 0 = forecasting on the kNN / physical / first-difference temporal graph, 1 = mask mode, 2 = banded line graph
// PIPE: the launch of the host entry point (mga_admm_solve_host), which waits for each chunk of y to arrive and reports
// finished chunks.  A separate instantiation: carrying the hand-shake in the kernel of the device entry point cost it
// 4 % (measured: 471 k vs 490 k windows/s at B = 1024) although it only runs once per window.
template <int CH, int K, int MAXT, int MINB, int VAR, bool PIPE>
__global__ void __launch_bounds__(MAXT, MINB) k_admm_resident(const ResArgs a) {
  constexpr int TS = 4 * CH;
  constexpr bool MASKM = VAR == 1, BANDM = VAR == 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cta<CH, K> cta;
  cta.init(a, smem_raw);
  Ctx<CH, K>& c = cta.c;
  const int N = a.N, T = a.T, TP = a.TP, t_in = a.t_in;
  const int i = cta.i, t0 = cta.t0, orig = cta.orig;
  const bool active = cta.active;
  c.bw = BANDM ? (a.band_uniform ? cta.band_s : a.band_w + orig) : nullptr;
  c.skip = a.skip; c.bN = a.band_uniform ? 1 : N;
  float* pbuf = cta.pbuf; float* qbuf = cta.qbuf; float* red = cta.red; float* dred = cta.dred; float* st_smem = cta.st_smem;

  // ---- parked ADMM state: [v][node * TP + t] as 128-bit chunks, in shared memory or in this CTA's slice
  // of an L2-resident scratch (same layout: a warp reads 32 consecutive 48-byte rows)
  float* st_base = a.state_in_smem ? st_smem : a.scratch + (size_t)blockIdx.x * ST_COUNT * N * TP;
  st_base += i * TP + t0;
  const int st_stride = N * TP;
  auto ld_state = [&](int v, float (&o)[TS]) {
    const float4* base = reinterpret_cast<const float4*>(st_base + v * st_stride);
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) g = base[cc];
      o[4 * cc] = g.x; o[4 * cc + 1] = g.y; o[4 * cc + 2] = g.z; o[4 * cc + 3] = g.w;
    }
  };
  auto st_state = [&](int v, const float (&o)[TS]) {
    if (!active) return;
    float4* base = reinterpret_cast<float4*>(st_base + v * st_stride);
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) base[cc] = make_float4(o[4 * cc], o[4 * cc + 1], o[4 * cc + 2], o[4 * cc + 3]);
  };
  // per-thread partial of one diagnostics column: reduced over the warp at once and left in
  // dred[col][warp] until the end of the outer iteration, so no column stays live in registers
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  auto diag_put = [&](int col, float v) {
    v = warp_sum<float>(v);
    if (lane == 0) dred[col * 32 + wp] = v;
  };

  // Windows are handed out dynamically (one atomic per window): CTAs that share an SM with fewer
  // neighbours near the end of the batch run faster and pick up more of the tail.
  int& s_next = *reinterpret_cast<int*>(red + 64);
  int& s_abort = *reinterpret_cast<int*>(red + 65);
  if (PIPE && threadIdx.x == 0) s_abort = 0;
  for (int64_t b = blockIdx.x; b < a.B;) {
    if (PIPE) {
      // the host entry point uploads the batch chunk by chunk while this kernel runs: wait for this window's chunk.
      // y is then read with ld.global.cg (below): a 128-byte L1 line may straddle two windows of different chunks.
      if (threadIdx.x == 0 && pipe_acquire(a.ready, a.chunk_up, a.epoch, b, a.abort_flag)) s_abort = 1;
      __syncthreads();
      if (s_abort) return;
    }
    const int y_rows = MASKM ? T : t_in;
    const float* yw = a.y + (size_t)b * y_rows * N + orig;
    const float* mw = MASKM ? a.mask + (size_t)b * T * N + orig : nullptr;
    // ---- initial_guess (ADMM.py:766-781) / initial_interpolation (ADMM.py:783-811) and initial state (ADMM.py:537-544)
    {
      float x[TS];
      float w, cc;
      if (MASKM) {
        float cnt = 0.f, st = 0.f, sy = 0.f, sty = 0.f, st2 = 0.f;
        for (int t = 0; t < T; ++t) {
          const float m = active ? mw[(size_t)t * N] : 1.f, v = active ? ldy<PIPE>(yw + (size_t)t * N) : 0.f, tt = (float)t;
          cnt += m; st += tt * m; sy += v * m; sty += tt * v * m; st2 += tt * tt * m;
        }
        const float tm = st / cnt, ym = sy / cnt, tym = sty / cnt, t2m = st2 / cnt;
        w = (tym - tm * ym) / (t2m - tm * tm);
        cc = ym - w * tm;
      } else {
        float sy = 0.f, sty = 0.f;
        for (int t = 0; t < t_in; ++t) {
          const float v = active ? ldy<PIPE>(yw + (size_t)t * N) : 0.f;
          sy += v;
          sty += (float)t * v;
        }
        const float my = sy / (float)t_in, mty = sty / (float)t_in;
        w = (mty - a.t_mean * my) / a.t_var;
        cc = my - w * a.t_mean;
      }
      {
        float tenth[TS];
#pragma unroll
        for (int k = 0; k < TS; ++k) {
          const int t = t0 + k;
          float v = 0.f;
          if (active && t < T) {
            if (MASKM) v = (w * (float)t + cc) * (1.f - mw[(size_t)t * N]) + ldy<PIPE>(yw + (size_t)t * N);
            else v = t < t_in ? ldy<PIPE>(yw + (size_t)t * N) : w * (float)t + cc;
          }
          x[k] = v;
          tenth[k] = (active && t < T) ? 0.1f : 0.f;
        }
        st_state(ST_GU, tenth); st_state(ST_GD, tenth); st_state(ST_GAM, tenth);
      }
      st_state(ST_X, x); st_state(ST_ZU, x); st_state(ST_ZD, x);
      // phi = L_d x  (ADMM.py:541)
      c.put(pbuf, x);
      __syncthreads();
      if (BANDM) {
        float q[TS];
        c.band_ldr(x, pbuf, q);
#pragma unroll
        for (int k = 0; k < TS; ++k) x[k] = q[k];
      } else {
        float qs[TS];
        c.shifted_ldr(x, qs);
        c.put(qbuf, qs);
        __syncthreads();
        const float qprev = c.has_prev ? qbuf[c.own - 1] : 0.f;
#pragma unroll
        for (int k = 0; k < TS; ++k) x[k] = (k == 0) ? qprev : qs[k - 1];
      }
      st_state(ST_PHI, x);
      __syncthreads();
    }

    for (int it = 0; it < a.n_outer; ++it) {
      float* al = a.alpha ? a.alpha + ((size_t)it * 3) * a.n_cg * a.B + b : nullptr;
      float* be = a.beta ? a.beta + ((size_t)it * 3) * a.n_cg * a.B + b : nullptr;
      const size_t sys_stride = (size_t)a.n_cg * a.B;
      float r[TS];
      // ---- RHS_x (ADMM.py:552-559): Ldr_T(gamma + rho phi)/2 + (rho_u zu + rho_d zd)/2 - (gu+gd)/2 + H^T y
      {
        float v[TS], f[TS];
        {
          float tmp[TS];
          ld_state(ST_GAM, v);
          ld_state(ST_PHI, tmp);
#pragma unroll
          for (int k = 0; k < TS; ++k) v[k] = v[k] + a.rho * tmp[k];
        }
        c.put(pbuf, v);
        __syncthreads();
        if (BANDM) {
          c.band_ldrt(v, pbuf, r);
        } else {
          {
            float vs[TS];
            const float vnext = c.has_next ? pbuf[c.own + TS] : 0.f;
#pragma unroll
            for (int k = 0; k < TS; ++k) vs[k] = (t0 + k + 1 < T) ? ((k < TS - 1) ? v[k + 1] : vnext) : 0.f;
            c.put(qbuf, vs);
          }
          __syncthreads();
          c.father_sum(f);
          {
            float vs[TS];
            c.get(qbuf, vs);
#pragma unroll
            for (int k = 0; k < TS; ++k) f[k] += c.wself * vs[k];
          }
#pragma unroll
          for (int k = 0; k < TS; ++k) {
            const int t = t0 + k;
            // rows of apply_op_Ldr_T: t = T-1 keeps v; t = 0 keeps the identity term only under Q1
            r[k] = (t == T - 1) ? v[k] : ((t == 0 && !a.q1) ? -f[k] : v[k] - f[k]);
          }
        }
        ld_state(ST_ZU, v); ld_state(ST_ZD, f);
#pragma unroll
        for (int k = 0; k < TS; ++k) v[k] = (a.rho_u * v[k] + a.rho_d * f[k]) / 2.f;
#pragma unroll
        for (int k = 0; k < TS; ++k) r[k] = r[k] / 2.f + v[k];
        ld_state(ST_GU, v); ld_state(ST_GD, f);
#pragma unroll
        for (int k = 0; k < TS; ++k) {
          const int t = t0 + k;
          float o = 0.f;
          if (active && t < T) {
            const float hty = t < y_rows ? ldy<PIPE>(yw + (size_t)t * N) : 0.f;
            o = r[k] - (v[k] + f[k]) / 2.f + hty;
          }
          r[k] = o;
        }
      }
      // ---- x solve (ADMM.py:571), warm start x_old
      {
        float x[TS];
        ld_state(ST_X, x);
        if (MASKM) {
          float m0[TS];
#pragma unroll
          for (int k = 0; k < TS; ++k) m0[k] = (active && t0 + k < T) ? mw[(size_t)(t0 + k) * N] : 0.f;
          c.template cg<MGA_SYS_X, true, false>(x, r, a.ax, a.cx, a.n_cg, al, be, a.B, m0);
        } else {
          c.template cg<MGA_SYS_X, false, BANDM>(x, r, a.ax, a.cx, a.n_cg, al, be, a.B, r);
        }
        if (a.want_diag) {
          float xo[TS];
          ld_state(ST_X, xo);
          float s2 = 0.f;
#pragma unroll
          for (int k = 0; k < TS; ++k) {
            if (active && t0 + k < T) {
              const float dx = x[k] - xo[k];
              s2 += dx * dx;
              if (a.dx_sum) atomicAdd(a.dx_sum + ((size_t)it * T + t0 + k) * N + orig, (double)dx);
            }
          }
          diag_put(MGA_DIAG_DX2, s2);
        }
        st_state(ST_X, x);
        // right-hand side of the z_u system (ADMM.py:579)
        ld_state(ST_GU, r);
#pragma unroll
        for (int k = 0; k < TS; ++k) r[k] = r[k] / 2.f + a.azu * x[k];
      }
      // ---- z_u solve (ADMM.py:579-580) + its dual ascent (ADMM.py:595)
      {
        float z[TS];
        ld_state(ST_ZU, z);
        c.template cg<MGA_SYS_ZU, false>(z, r, a.azu, a.czu, a.n_cg, al ? al + sys_stride : nullptr,
                                         be ? be + sys_stride : nullptr, a.B, r);
        float x[TS], g[TS];
        ld_state(ST_X, x);
        ld_state(ST_GU, g);
        float s0 = 0.f, s1 = 0.f;
        {
          float zo[TS];
          ld_state(ST_ZU, zo);
#pragma unroll
          for (int k = 0; k < TS; ++k) {
            const float d0 = x[k] - z[k], d1 = z[k] - zo[k];
            s0 += d0 * d0;
            s1 += d1 * d1;
            g[k] = g[k] + a.rho_u * d0;
          }
        }
        st_state(ST_GU, g);
        st_state(ST_ZU, z);
        if (a.want_diag) { diag_put(MGA_DIAG_X_ZU2, s0); diag_put(MGA_DIAG_DZU2, s1); }
        // right-hand side of the z_d system (ADMM.py:587)
        ld_state(ST_GD, r);
#pragma unroll
        for (int k = 0; k < TS; ++k) r[k] = r[k] / 2.f + a.azd * x[k];
      }
      // ---- z_d solve (ADMM.py:587-588) + its dual ascent (ADMM.py:597)
      {
        float z[TS];
        ld_state(ST_ZD, z);
        c.template cg<MGA_SYS_ZD, false, BANDM>(z, r, a.azd, a.czd, a.n_cg, al ? al + 2 * sys_stride : nullptr,
                                         be ? be + 2 * sys_stride : nullptr, a.B, r);
        float x[TS], g[TS];
        ld_state(ST_X, x);
        ld_state(ST_GD, g);
        float s0 = 0.f, s1 = 0.f;
        {
          float zo[TS];
          ld_state(ST_ZD, zo);
#pragma unroll
          for (int k = 0; k < TS; ++k) {
            const float d0 = x[k] - z[k], d1 = z[k] - zo[k];
            s0 += d0 * d0;
            s1 += d1 * d1;
            g[k] = g[k] + a.rho_d * d0;
          }
        }
        st_state(ST_GD, g);
        st_state(ST_ZD, z);
        if (a.want_diag) { diag_put(MGA_DIAG_X_ZD2, s0); diag_put(MGA_DIAG_DZD2, s1); }
      }
      // ---- phi prox + gamma ascent (ADMM.py:600-605) and the remaining diagnostics (ADMM.py:612-637)
      {
        float x[TS], qs[TS];
        ld_state(ST_X, x);
        c.put(pbuf, x);
        __syncthreads();
        if (BANDM) {
          c.band_ldr(x, pbuf, qs);
        } else {
          c.shifted_ldr(x, qs);
          c.put(qbuf, qs);
          __syncthreads();
        }
        if (a.want_diag) {     // GLR = x . L_u x, recover = ||Hx - y||^2
          float lux[TS];
#pragma unroll
          for (int k = 0; k < TS; ++k) lux[k] = 0.f;
          c.fwd_u(lux);
          float sg = 0.f, sr = 0.f;
#pragma unroll
          for (int k = 0; k < TS; ++k) {
            const int t = t0 + k;
            if (active && t < T) {
              sg += x[k] * (x[k] - lux[k]);
              if (MASKM) {             // ||x * mask - y|| (ADMM.py:620-621)
                const float h = x[k] * mw[(size_t)t * N] - ldy<PIPE>(yw + (size_t)t * N);
                sr += h * h;
              } else if (t < t_in) {
                const float h = x[k] - ldy<PIPE>(yw + (size_t)t * N);
                sr += h * h;
              }
            }
          }
          diag_put(MGA_DIAG_GLR, sg);
          diag_put(MGA_DIAG_RECOVER2, sr);
        }
        int bad = 0;
#pragma unroll
        for (int k = 0; k < TS; ++k) bad |= !isfinite(x[k]);
        // x is dead from here: reuse its registers for (L_d x), gamma and phi
        if (!BANDM) {
          const float qprev = c.has_prev ? qbuf[c.own - 1] : 0.f;
#pragma unroll
          for (int k = TS - 1; k > 0; --k) qs[k] = qs[k - 1];        // qs[k] = (L_d x)[t0 + k]
          qs[0] = qprev;
        }
        float gam[TS], phi[TS];
        ld_state(ST_GAM, gam);
        ld_state(ST_PHI, phi);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k = 0; k < TS; ++k) {
          const int t = t0 + k;
          const float q = qs[k];
          const float ph = soft_thr(q - gam[k] / a.rho, a.thr);
          const float gn = gam[k] + a.rho * (ph - q);
          if (active && t < T) {
            bad |= !isfinite(ph) || !isfinite(gn);
            const float e = ph - q, f = ph - phi[k];
            s0 += e * e;
            s1 += f * f;
            s2 += fabsf(q);
            s3 += q * q;
            phi[k] = ph;
            gam[k] = gn;
          }
        }
        st_state(ST_PHI, phi);
        st_state(ST_GAM, gam);
        if (a.want_diag) {
          diag_put(MGA_DIAG_PHI_LDX2, s0);
          diag_put(MGA_DIAG_DPHI2, s1);
          diag_put(MGA_DIAG_DGTV, s2);
          diag_put(MGA_DIAG_DGLR, s3);
          diag_put(MGA_DIAG_NONFINITE, (float)bad);
        }
        __syncthreads();
        if (a.want_diag && threadIdx.x < MGA_DIAG_COLS && a.diag) {
          const int nw = (blockDim.x + 31) >> 5;
          float tot = 0.f;
          for (int k = 0; k < nw; ++k) tot += dred[threadIdx.x * 32 + k];
          if (tot != 0.f) atomicAdd(a.diag + (size_t)it * MGA_DIAG_COLS + threadIdx.x, (double)tot);
        }
        // the next writer of dred / pbuf is separated from these reads by the barriers of the next solve
      }
    }
    // ---- results
    {
      float o[TS];
      ld_state(ST_X, o);
      if (active) {
#pragma unroll
        for (int k = 0; k < TS; ++k)
          if (t0 + k < T) a.x_out[((size_t)b * T + t0 + k) * N + orig] = o[k];
      }
      for (int v = ST_ZU; v < ST_COUNT; ++v) {
        if (a.out[v]) {
          ld_state(v, o);
          if (active) {
#pragma unroll
            for (int k = 0; k < TS; ++k)
              if (t0 + k < T) a.out[v][((size_t)b * T + t0 + k) * N + orig] = o[k];
          }
        }
      }
    }
    if (PIPE) {
      // hand the finished window to the download stream: stores -> device-scope fence -> CTA barrier -> one count per window;
      // the CTA that completes a chunk publishes it to the host (system-scope fence, then the flag in mapped host memory)
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) pipe_release(a.done, a.host_done, a.chunk, a.epoch, b, a.B);
    }
    if (threadIdx.x == 0) s_next = atomicAdd(a.next_window, 1);
    __syncthreads();
    b = (int64_t)gridDim.x + s_next;
  }
}

// ---- launch geometry --------------------------------------------------------------------------
struct ResGeom {
  int CH, S, NT, TP, threads, Kt;
  size_t core_bytes, state_bytes;
};

inline int res_tp(int min_len) {          // smallest 4 * odd >= min_len
  int chunks = (min_len + 3) / 4;
  if ((chunks & 1) == 0) ++chunks;
  return 4 * chunks;
}

// CH = chunks of 4 time steps per thread.  Default: the largest CH <= 3 (a node's table rows and
// in-list entries are then read once per 12 time steps; measured 252k vs 210k windows/s against
// CH = 1 at PEMS04 shape on B200); MGA_RES_CH overrides for experiments.
inline int res_kt(int kd_eff, int ku_eff) {            // the K (slots per forward table) of the kernel instantiation
  const int kk = std::max(kd_eff, ku_eff);
  return kk <= 4 ? 4 : (kk <= 6 ? 6 : (kk <= 8 ? 8 : 10));
}

inline bool res_geometry(const GraphDev& g, int kd_eff, int ku_eff, int ell_total, int force_ch, ResGeom* out) {
  const int Kt = res_kt(kd_eff, ku_eff);
  const int NT = ((g.N + 31) / 32) * 32;
  const int chunks = (g.T + 3) / 4;
  for (int ch = std::min(3, chunks); ch >= 1; --ch) {
    if (force_ch > 0 && ch != force_ch) continue;
    const int S = (chunks + ch - 1) / ch;
    if ((int64_t)S * NT > 1024) continue;
    ResGeom r;
    r.CH = ch; r.S = S; r.NT = NT; r.threads = S * NT;
    r.TP = res_tp(S * 4 * ch);
    r.Kt = Kt;
    const size_t rows = (size_t)(g.N + 8) * r.TP;
    r.core_bytes = 2 * rows * 4 + 68 * 4 + MGA_DIAG_COLS * 32 * 4 + (size_t)(ell_total + 64) * 8;
#if MGA_RES_TAB_SMEM
    r.core_bytes += (size_t)(2 * Kt) * r.threads * 8;
#endif
    if (g.temporal == MGA_TEMPORAL_BAND) r.core_bytes += (size_t)((g.T * g.skip + 3) / 4 * 4) * 4;
    r.state_bytes = (size_t)ST_COUNT * g.N * r.TP * 4;
    *out = r;
    return true;
  }
  return false;
}

inline int res_forced_ch() {
  const char* e = std::getenv("MGA_RES_CH");
  return e ? std::atoi(e) : 0;
}

template <int CH, int K, int MAXT, int MINB, int VAR>
inline int launch_res(mga_plan* p, ResArgs& a, const ResGeom& geo, cudaStream_t st) {
  if (a.ready && VAR != 0) { set_error("resident: the pipelined host launch is built for the forecasting kernel only"); return MGA_ERR_UNSUPPORTED; }
  auto kern = (a.ready && VAR == 0) ? k_admm_resident<CH, K, MAXT, MINB, VAR, VAR == 0> : k_admm_resident<CH, K, MAXT, MINB, VAR, false>;
  const size_t core = geo.core_bytes;
  const size_t with_state = core + geo.state_bytes;
  // State in shared memory only if it does not cost residency: compare CTAs/SM both ways.
  int occ_core = 0, occ_state = 0;
  MGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem_optin));
  MGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_core, kern, geo.threads, core));
  if (with_state <= (size_t)p->max_smem_optin)
    MGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_state, kern, geo.threads, with_state));
  if (occ_core < 1) { set_error("resident kernel does not fit on an SM"); return MGA_ERR_UNSUPPORTED; }
  a.state_in_smem = (occ_state >= occ_core) ? 1 : 0;
  const int occ = a.state_in_smem ? occ_state : occ_core;
  const size_t smem = a.state_in_smem ? with_state : core;
  int64_t grid = std::min<int64_t>(a.B, (int64_t)occ * p->sm_count);
  if (!a.state_in_smem) {
    // two launches may be in flight (res_slot): each owns a launch-independent half of the scratch, sized for a full
    // grid - a stride that followed THIS launch's grid let a short last chunk park its state inside its predecessor's
    const size_t per_slot = (size_t)occ * p->sm_count * geo.state_bytes;
    int rc = ensure_workspace(p, p->ws, 2 * per_slot);
    if (rc) return rc;
    a.scratch = reinterpret_cast<float*>(static_cast<char*>(p->ws.base) + (size_t)(p->res_slot & 1) * per_slot);
  }
  a.next_window = p->r_counters + (p->r_counter_next++ % mga_plan::kCounters) * 32;
  MGA_CUDA(cudaMemsetAsync(a.next_window, 0, sizeof(int), st));
  if (std::getenv("MGA_RES_VERBOSE"))
    std::fprintf(stderr, "[mga] resident<%d,%d,%d,%d>: threads %d, smem %zu B (state %s), %d CTA/SM, grid %lld\n", CH, K,
                 MAXT, MINB, geo.threads, smem, a.state_in_smem ? "smem" : "L2 scratch", occ, (long long)grid);
  kern<<<(unsigned)grid, geo.threads, smem, st>>>(a);
  MGA_LAUNCH_CHECK("k_admm_resident");
  return MGA_OK;
}

template <int CH, int K, int MAXT, int MINB>
inline int launch_res_cg(mga_plan* p, ResArgs& a, const CgArgs& g, const ResGeom& geo, cudaStream_t st) {
  auto kern = k_cg_resident<CH, K, MAXT, MINB>;
  int occ = 0;
  MGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem_optin));
  MGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, geo.threads, geo.core_bytes));
  if (occ < 1) { set_error("resident CG kernel does not fit on an SM"); return MGA_ERR_UNSUPPORTED; }
  const int64_t grid = std::min<int64_t>(a.B, (int64_t)occ * p->sm_count);
  a.state_in_smem = 1;
  a.next_window = p->r_counters + (p->r_counter_next++ % mga_plan::kCounters) * 32;
  MGA_CUDA(cudaMemsetAsync(a.next_window, 0, sizeof(int), st));
  kern<<<(unsigned)grid, geo.threads, geo.core_bytes, st>>>(a, g);
  MGA_LAUNCH_CHECK("k_cg_resident");
  return MGA_OK;
}

template <int CH, int K>
inline int pick_threads_cg(mga_plan* p, ResArgs& a, const CgArgs& g, const ResGeom& geo, cudaStream_t st) {
  if (geo.threads <= 192) return launch_res_cg<CH, K, 192, (MGA_RES_MINB > 1 ? 3 : 1)>(p, a, g, geo, st);
  if (geo.threads <= 320) return launch_res_cg<CH, K, 320, MGA_RES_MINB>(p, a, g, geo, st);
  if (geo.threads <= 512) return launch_res_cg<CH, K, 512, 1>(p, a, g, geo, st);
  if (geo.threads <= 640) return launch_res_cg<CH, K, 640, 1>(p, a, g, geo, st);      // 2 slabs x 320: 96 registers, no spills
  return launch_res_cg<CH, K, 1024, 1>(p, a, g, geo, st);
}

template <int CH, int K, int MASKM>
inline int pick_threads_m(mga_plan* p, ResArgs& a, const ResGeom& geo, cudaStream_t st) {
  if (geo.threads <= 192) return launch_res<CH, K, 192, (MGA_RES_MINB > 1 ? 3 : 1), MASKM>(p, a, geo, st);
  if (geo.threads <= 320) return launch_res<CH, K, 320, MGA_RES_MINB, MASKM>(p, a, geo, st);
  if (geo.threads <= 512) return launch_res<CH, K, 512, 1, MASKM>(p, a, geo, st);
  if (geo.threads <= 640) return launch_res<CH, K, 640, 1, MASKM>(p, a, geo, st);      // 2 slabs x 320: 96 registers, no spills
  return launch_res<CH, K, 1024, 1, MASKM>(p, a, geo, st);
}

template <int CH, int K>
inline int pick_threads(mga_plan* p, ResArgs& a, const ResGeom& geo, cudaStream_t st) {
  if (a.band_w) return pick_threads_m<CH, K, 2>(p, a, geo, st);
  return a.mask ? pick_threads_m<CH, K, 1>(p, a, geo, st) : pick_threads_m<CH, K, 0>(p, a, geo, st);
}

}  // namespace mga
