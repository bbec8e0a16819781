// Streaming mode, fast path ("v2"): for windows too large for one CTA (T = 288, 20 000 nodes) in the
// unrolled (fixed iteration count) forecasting mode.  mga_stream.cu stays the general path (fp64, masks,
// ablations, banded line graph, time-varying weights, tolerance mode).
//
// What differs from mga_stream.cu
//  * Layout.  All vectors the solver owns are NODE-major per window, v[b][n][TP4] (TP4 = T rounded up to
//    4, pads kept at zero), nodes in reverse-Cuthill-McKee order.  One thread owns one 16-byte chunk =
//    4 consecutive time steps of one node; consecutive threads own consecutive chunks, so every own
//    access is a coalesced 128-bit load/store, and a gather of "neighbour n', same 4 time steps" is a
//    128-bit load whose 32-byte sector is fully used by the lanes next to it (they want the next chunks
//    of the same neighbour row).  The reference's (B, T, N) layout is converted once on the way in
//    (k2_init / k2_import) and once on the way out (k2_export).
//  * The shift (as in the resident kernel).  L_d reads p[t-1] and L_d^T reads q[t+1] (ADMM.py:171,
//    200-208); q is stored shifted, qs[t] = q[t+1], so both gathers are chunk-aligned; only the thread's
//    own row is touched off-chunk (one scalar each).
//  * No recomputation at the gather: p' = r + beta p gets its own elementwise kernel (12 B/pt), because
//    gathering r AND p at every neighbour doubles the L2 traffic, which is what bounds these kernels.
//
// Per CG iteration of the x / z_d systems: k2_pupdate (12 B/pt), k2_ldr_shift (p' -> qs, 8 B/pt),
// k2_ldrt_lhs (p', qs -> Ap, <p',Ap>, 12 B/pt), k2_xr (24 B/pt) = 56 B/pt of HBM traffic against the
// 48 B/pt the algorithm needs; z_u: 12 + 8 + 24 = 44 against 40.  Gathers are served by L1/L2: a window
// (<= 1.9 MB per vector) is far smaller than the 126 MB L2.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "mga_common.cuh"

namespace mga {

#ifndef MGA_S2_BLOCK
#define MGA_S2_BLOCK 256
#endif
constexpr int kB2 = MGA_S2_BLOCK;

__device__ __forceinline__ float4 ld4(const float* v, size_t chunk) { return reinterpret_cast<const float4*>(v)[chunk]; }
__device__ __forceinline__ void st4(float* v, size_t chunk, float4 a) { reinterpret_cast<float4*>(v)[chunk] = a; }
// streamed (evict-first) accesses of the time-tiled kernels: the vectors pass through once, the L1 lines are kept
// for the graph tables (ncu: with plain loads every table read of phase 2 was an L2 round trip - long_scoreboard)
#ifndef MGA_S3_STREAM_HINTS
#define MGA_S3_STREAM_HINTS 1
#endif
__device__ __forceinline__ float4 ld4s(const float* v, size_t chunk) {
#if MGA_S3_STREAM_HINTS
  return __ldcs(reinterpret_cast<const float4*>(v) + chunk);
#else
  return reinterpret_cast<const float4*>(v)[chunk];
#endif
}
__device__ __forceinline__ float ld1s(const float* v, size_t k) {
#if MGA_S3_STREAM_HINTS
  return __ldcs(v + k);
#else
  return v[k];
#endif
}
__device__ __forceinline__ void st4s(float* v, size_t chunk, float4 a) {
#if MGA_S3_STREAM_HINTS
  __stcs(reinterpret_cast<float4*>(v) + chunk, a);
#else
  reinterpret_cast<float4*>(v)[chunk] = a;
#endif
}

struct Chunk {
  int b;           // window
  int n, c;        // node (internal), chunk of 4 time steps
  int q;           // chunk index inside the window: n * C4 + c
  size_t g;        // global chunk index
  bool ok;
};

// A CTA owns a 2-D tile of one window: NB consecutive nodes x CB consecutive chunks; block = (CB, NB),
// grid = (B, tilesN, tilesC), so no thread ever divides.  Lanes run along the chunks (a quarter-warp reads
// >= 96 contiguous bytes of one row); RCM-consecutive nodes share most of their neighbours, so the rows a
// tile gathers are re-used from L1.
__device__ __forceinline__ Chunk locate2(const Graph2& g) {
  Chunk k;
  k.b = blockIdx.x;
  k.n = blockIdx.y * g.NB + threadIdx.y;
  k.c = blockIdx.z * g.CB + threadIdx.x;
  k.ok = threadIdx.y < g.NB && k.n < g.N && k.c < g.C4;
  if (!k.ok) { k.n = 0; k.c = 0; }
  k.q = k.n * g.C4 + k.c;
  k.g = (size_t)k.b * (size_t)(g.N * g.C4) + k.q;
  return k;
}
__device__ __forceinline__ const float4* win4(const float* v, const Graph2& g, int b) {
  return reinterpret_cast<const float4*>(v) + (size_t)b * (size_t)(g.N * g.C4);
}
__device__ __forceinline__ int tid2() { return threadIdx.y * blockDim.x + threadIdx.x; }

// elementwise kernels: consecutive threads own consecutive chunks of one window; grid = (B, ceil(Q / kFlat))
constexpr int kFlat = 256;
__device__ __forceinline__ Chunk locate_flat(const Graph2& g) {
  Chunk k;
  const int Q = g.N * g.C4;
  k.b = blockIdx.x;
  k.q = blockIdx.y * kFlat + threadIdx.x;
  k.ok = k.q < Q;
  if (!k.ok) k.q = 0;
  k.n = 0; k.c = 0;
  k.g = (size_t)k.b * (size_t)Q + k.q;
  return k;
}

// block sum -> one double atomicAdd per CTA; only warp 0 adds up the warp partials
__device__ __forceinline__ void block_add(float v, double* slot, int tid, int nthreads) {
  __shared__ float red[32];
  const int lane = tid & 31, w = tid >> 5, nw = (nthreads + 31) >> 5;
  v = warp_sum<float>(v);
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum<float>(t);
    if (lane == 0) atomicAdd(slot, (double)t);
  }
}

// ---- layout conversion ----------------------------------------------------------------------------
// (B, T, N) caller order -> node-major internal
__global__ void __launch_bounds__(kB2) k2_import(Graph2 g, const float* __restrict__ src, float* __restrict__ dst) {
  const Chunk k = locate2(g);
  if (!k.ok) return;
  const int o = g.perm[k.n];
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int t = 4 * k.c + j;
    v[j] = t < g.T ? src[((size_t)k.b * g.T + t) * g.N + o] : 0.f;
  }
  st4(dst, k.g, make_float4(v[0], v[1], v[2], v[3]));
}

// node-major internal -> (B, T, N) caller order
__global__ void __launch_bounds__(kB2) k2_export(Graph2 g, const float* __restrict__ src, float* __restrict__ dst) {
  const Chunk k = locate2(g);
  if (!k.ok) return;
  const int o = g.perm[k.n];
  const float4 a = ld4(src, k.g);
  const float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int t = 4 * k.c + j;
    if (t < g.T) dst[((size_t)k.b * g.T + t) * g.N + o] = v[j];
  }
}

// initial_guess (ADMM.py:766-781) + initial state (ADMM.py:537-544); one WARP per (window, node): the lanes share the
// regression sums (shuffle) and then write the node's row chunk by chunk - 512 contiguous bytes per store instruction
// (one thread per row wrote 16 bytes every 1152: 1.6 TB/s at T = 288).  The sums run over t in the reference's order
// per lane and are combined across lanes; float32 rounding of the regression is at the 1e-7 level either way.
__global__ void __launch_bounds__(256) k2_init(Graph2 g, int64_t B, const float* __restrict__ y, float* __restrict__ x,
                                               float* __restrict__ zu, float* __restrict__ zd, float* __restrict__ gu,
                                               float* __restrict__ gd, float* __restrict__ gam, float t_mean, float t_var) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (idx >= B * g.N) return;
  const int64_t b = idx / g.N;
  const int n = (int)(idx - b * g.N);
  const float* yw = y + b * (int64_t)g.t_in * g.N + g.perm[n];
  float sy = 0.f, sty = 0.f;
  for (int t = lane; t < g.t_in; t += 32) {
    const float v = yw[(size_t)t * g.N];
    sy += v;
    sty += (float)t * v;
  }
  sy = warp_sum<float>(sy);
  sty = warp_sum<float>(sty);
  const float my = sy / (float)g.t_in, mty = sty / (float)g.t_in;
  const float w = (mty - t_mean * my) / t_var;
  const float c = my - w * t_mean;
  const size_t row = (size_t)idx * g.C4;
  for (int cc = lane; cc < g.C4; cc += 32) {
    float v[4], tenth[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = 4 * cc + j;
      v[j] = t < g.T ? (t < g.t_in ? yw[(size_t)t * g.N] : w * (float)t + c) : 0.f;
      tenth[j] = t < g.T ? 0.1f : 0.f;
    }
    const float4 a = make_float4(v[0], v[1], v[2], v[3]), d = make_float4(tenth[0], tenth[1], tenth[2], tenth[3]);
    st4(x, row + cc, a); st4(zu, row + cc, a); st4(zd, row + cc, a);
    st4(gu, row + cc, d); st4(gd, row + cc, d); st4(gam, row + cc, d);
  }
}

// ---- operator pieces ------------------------------------------------------------------------------
// sum_j w_j v[nbr_j][chunk c] over a forward ELL table (rows of `stride` slots, -1 = no neighbour)
// Gathers index the window with 32-bit arithmetic (vw = the vector's window base as float4*); four table
// entries per trip, their indices / weights first, then the four 128-bit loads, then the FMAs.
__device__ __forceinline__ float4 fwd_gather(const int* __restrict__ nbr, const float* __restrict__ w, int slots,
                                             const float4* __restrict__ vw, int n, int c, int C4) {
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int* nb = nbr + n * slots;
  const float* ww = w + n * slots;
  int j = 0;
  for (; j + 4 <= slots; j += 4) {
    int m[4];
    float wj[4];
    float4 a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { m[q] = nb[j + q]; wj[q] = ww[j + q]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = vw[max(m[q], 0) * C4 + c];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (m[q] < 0) continue;       // "-1 = no neighbour" contributes nothing (quirk Q6)
      acc.x += wj[q] * a[q].x; acc.y += wj[q] * a[q].y; acc.z += wj[q] * a[q].z; acc.w += wj[q] * a[q].w;
    }
  }
  for (; j < slots; ++j) {
    const int m = nb[j];
    if (m < 0) continue;
    const float wj = ww[j];
    const float4 a = vw[m * C4 + c];
    acc.x += wj * a.x; acc.y += wj * a.y; acc.z += wj * a.z; acc.w += wj * a.w;
  }
  return acc;
}

// qs = shifted L_d v:  qs[t] = q[t+1] = v[t+1] - sum_j w_j v_nbr[t]  (0 for t+1 >= T)   ADMM.py:166-177
__global__ void __launch_bounds__(kB2) k2_ldr_shift(Graph2 g, const float* __restrict__ v, float* __restrict__ qs) {
  const Chunk k = locate2(g);
  if (!k.ok) return;
  const float4* vw = win4(v, g, k.b);
  const float4 own = vw[k.q];
  const float nxt = (k.c + 1 < g.C4) ? reinterpret_cast<const float*>(vw + k.q + 1)[0] : 0.f;
  const float4 acc = fwd_gather(g.nbr_d, g.w_d, g.kd, vw, k.n, k.c, g.C4);
  const int t = 4 * k.c;
  float4 o;
  o.x = (t + 1 < g.T) ? own.y - acc.x : 0.f;
  o.y = (t + 2 < g.T) ? own.z - acc.y : 0.f;
  o.z = (t + 3 < g.T) ? own.w - acc.z : 0.f;
  o.w = (t + 4 < g.T) ? nxt - acc.w : 0.f;
  st4(qs, k.g, o);
}

// father sum over the in-list: f[t] = sum w qs_src[t]  (ADMM.py:200-209 as a gather, scatter order kept)
__device__ __forceinline__ float4 inlist_gather(const Graph2& g, const float4* __restrict__ qw, int n, int c) {
  float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
  int e = g.in_ptr[n];
  const int e1 = g.in_ptr[n + 1];
  for (; e + 4 <= e1; e += 4) {
    int m[4];
    float w[4];
    float4 a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { m[q] = g.in_src[e + q]; w[q] = g.in_w[e + q]; }
#pragma unroll
    for (int q = 0; q < 4; ++q) a[q] = qw[m[q] * g.C4 + c];
#pragma unroll
    for (int q = 0; q < 4; ++q) { f.x += w[q] * a[q].x; f.y += w[q] * a[q].y; f.z += w[q] * a[q].z; f.w += w[q] * a[q].w; }
  }
  for (; e < e1; ++e) {
    const float w = g.in_w[e];
    const float4 a = qw[g.in_src[e] * g.C4 + c];
    f.x += w * a.x; f.y += w * a.y; f.z += w * a.z; f.w += w * a.w;
  }
  return f;
}

// MODE 0: ap = A v, dot <v, ap> -> slot      (CG iteration, phase 1)
// MODE 1: r = rhs - A v, dot <r, r> -> slot  (initial residual)
// A = diag + c L_d^T L_d with qs = shifted L_d v already computed.  xsys: H^T H term + LHS_x's evaluation order.
template <int MODE>
__global__ void __launch_bounds__(kB2) k2_ldrt_lhs(Graph2 g, const float* __restrict__ v, const float* __restrict__ qs,
                                                   const float* __restrict__ rhs, float* __restrict__ out, double* __restrict__ slot,
                                                   float a, float cc, int xsys) {
  const Chunk k = locate2(g);
  float dot = 0.f;
  if (k.ok) {
    const float4* qw = win4(qs, g, k.b);
    const float4 pv = ld4(v, k.g);
    const float4 q1 = qw[k.q];
    const float qprev = k.c > 0 ? reinterpret_cast<const float*>(qw + k.q)[-1] : 0.f;   // q[4c] = qs[4c-1]; q[0] = 0 (ADMM.py:176)
    const float4 f = inlist_gather(g, qw, k.n, k.c);
    const float p[4] = {pv.x, pv.y, pv.z, pv.w};
    const float q[4] = {qprev, q1.x, q1.y, q1.z};
    const float ff[4] = {f.x, f.y, f.z, f.w};
    float o[4];
    const int t0 = 4 * k.c;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + j;
      const float l = q[j] - ff[j];      // row T-1: f = 0 because qs[T-1] = 0; Q1 is moot as q[0] = 0
      float val;
      if (xsys) val = ((t < g.t_in ? p[j] : 0.f) + a * p[j]) + cc * l;      // ADMM.py:372-379
      else val = cc * l + a * p[j];                                          // ADMM.py:394
      o[j] = t < g.T ? val : 0.f;
    }
    if (MODE == 0) {
      st4(out, k.g, make_float4(o[0], o[1], o[2], o[3]));
      dot = (p[0] * o[0] + p[1] * o[1]) + (p[2] * o[2] + p[3] * o[3]);
    } else {
      const float4 rh = ld4(rhs, k.g);
      const float r0 = rh.x - o[0], r1 = rh.y - o[1], r2 = rh.z - o[2], r3 = rh.w - o[3];
      st4(out, k.g, make_float4(r0, r1, r2, r3));
      dot = (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
    }
  }
  block_add(dot, slot + k.b, tid2(), blockDim.x * blockDim.y);
}

// z_u system: A = c L_u + a I  (ADMM.py:389-390), same two modes
template <int MODE>
__global__ void __launch_bounds__(kB2) k2_lu_lhs(Graph2 g, const float* __restrict__ v, const float* __restrict__ rhs,
                                                 float* __restrict__ out, double* __restrict__ slot, float a, float cc) {
  const Chunk k = locate2(g);
  float dot = 0.f;
  if (k.ok) {
    const float4* vw = win4(v, g, k.b);
    const float4 pv = vw[k.q];
    const float4 acc = fwd_gather(g.nbr_u, g.w_u, g.ku, vw, k.n, k.c, g.C4);
    float4 o;
    o.x = cc * (pv.x - acc.x) + a * pv.x;
    o.y = cc * (pv.y - acc.y) + a * pv.y;
    o.z = cc * (pv.z - acc.z) + a * pv.z;
    o.w = cc * (pv.w - acc.w) + a * pv.w;    // pads: v = 0 and every gathered pad is 0
    if (MODE == 0) {
      st4(out, k.g, o);
      dot = (pv.x * o.x + pv.y * o.y) + (pv.z * o.z + pv.w * o.w);
    } else {
      const float4 rh = ld4(rhs, k.g);
      const float4 r = make_float4(rh.x - o.x, rh.y - o.y, rh.z - o.z, rh.w - o.w);
      st4(out, k.g, r);
      dot = (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
    }
  }
  block_add(dot, slot + k.b, tid2(), blockDim.x * blockDim.y);
}

// p' = r + beta p   (first iteration: p' = r).  dots: RR(k) = dots[2k], PAP(k) = dots[2k+1], each (B)
__global__ void __launch_bounds__(kFlat) k2_pupdate(Graph2 g, int64_t B, int it, const float* __restrict__ r,
                                                   float* __restrict__ p, const double* __restrict__ dots) {
  const Chunk k = locate_flat(g);
  if (!k.ok) return;
  const float4 rv = ld4(r, k.g);
  if (it == 0) { st4(p, k.g, rv); return; }
  const float beta = (float)dots[(size_t)(2 * it) * B + k.b] / (float)dots[(size_t)(2 * it - 2) * B + k.b];   // ADMM.py:356
  const float4 pv = ld4(p, k.g);
  st4(p, k.g, make_float4(rv.x + beta * pv.x, rv.y + beta * pv.y, rv.z + beta * pv.z, rv.w + beta * pv.w));
}

// x += alpha p ; r -= alpha Ap ; RR(k+1) += r.r   (ADMM.py:350-355)
// x_in: the iterate this step starts from - the warm start x0 in the first iteration (read in place, no copy of
// the warm start is made), x itself afterwards
// XUP = false: r only - the x update of this iteration is applied by the NEXT iteration's k4_cg, which stages p anyway
// (K4Args::xd_in / xd_out): 12 instead of 24 B/pt here, 8 B/pt more in a kernel that is not bound by HBM
template <bool XUP = true>
__global__ void __launch_bounds__(kFlat) k2_xr(Graph2 g, int64_t B, int it, const float* x_in, float* x, float* __restrict__ r,
                                              const float* __restrict__ p, const float* __restrict__ ap, double* __restrict__ dots) {
  const Chunk k = locate_flat(g);
  float dot = 0.f;
  if (k.ok) {
    const float alpha = (float)dots[(size_t)(2 * it) * B + k.b] / (float)dots[(size_t)(2 * it + 1) * B + k.b];
    const float4 rv = ld4(r, k.g), av = ld4(ap, k.g);
    if (XUP) {
      const float4 xv = ld4(x_in, k.g), pv = ld4(p, k.g);
      st4(x, k.g, make_float4(xv.x + alpha * pv.x, xv.y + alpha * pv.y, xv.z + alpha * pv.z, xv.w + alpha * pv.w));
    }
    const float4 rn = make_float4(rv.x - alpha * av.x, rv.y - alpha * av.y, rv.z - alpha * av.z, rv.w - alpha * av.w);
    st4(r, k.g, rn);
    dot = (rn.x * rn.x + rn.y * rn.y) + (rn.z * rn.z + rn.w * rn.w);
  }
  block_add(dot, dots + (size_t)(2 * it + 2) * B + k.b, threadIdx.x, kFlat);
}

// ---- time-tiled kernels: the gathered vector AND the graph table staged in shared memory -------------------
// For graphs whose node set fits one CTA's shared memory (the PEMS graphs at any T) a CTA owns ALL nodes of one
// window over a tile of CB3 chunks (4 * CB3 time steps).  Phase 1 streams the tile in with coalesced 128-bit
// loads (and applies the p update on the way: p' = r + beta p is written once to HBM and once to shared memory),
// phase 2 gathers the neighbours' chunks from shared memory - the gathers that bound k2_ldr_shift / k2_ldrt_lhs /
// k2_lu_lhs no longer leave the SM.  Because the gathers are chunk-aligned (shifted q) a tile needs no halo
// columns, only one halo SCALAR per node (p'[first step of the next tile] / qs[last step of the previous tile]).
// The p update moves into the operator kernel, so per CG iteration: x / z_d systems (r,p -> p',qs : 16) +
// (p',qs -> Ap : 12) + k2_xr 24 = 52 B/pt in 3 launches (k2: 56 in 4); z_u (r,p -> p',Ap : 16) + 24 = 40 B/pt = the
// algorithmic minimum, in 2 launches (k2: 44 in 3).  p is ping-ponged (a tile's halo reads the next tile's OLD p).
// First version (tables read through L1, one CTA per tile): no faster than k2 - ncu showed 231 warp instructions per
// 32 chunks and every table read of phase 2 a long-scoreboard stall.  Hence: CTAs are persistent (grid = resident
// CTAs, tiles handed out round-robin), the table is staged ONCE per CTA in shared memory as (byte offset of the
// neighbour's row in the tile, weight) - entry -> LDS.64, gather -> LDS.128 at [thread base + offset], no integer
// math in between -, the self link leaves the tables (the owner holds the value), rows are unrolled over the
// table width, and all per-thread predicates / pointers are hoisted out of the row loop.

struct Smem3 {
  float4* tile;    // (nb, R3, CB) chunks: the gathered vector
  float4* tile2;   // (nb, R3, CB) chunks: second operand of the tile (p of p' = r + beta p; the vector A is applied to)
  float* halo;     // (nb, R3)
  float* halo2;    // (R3)
  int* ext;        // (R3) global rows of the tile's external rows (node-tiled plans)
  float* wself;    // (NT3)
  int* ptr;        // (NT3 + 1) in-list offsets (k3_ldrt_lhs only)
  int* ord;        // (NT3) row order (k3_ldrt_lhs only)
  int2* tab;       // the tile's table entries
  int stride_t;    // float4 between the two tile buffers of the double-buffered mode (0: single buffer)
  int stride_h;
};
__device__ __forceinline__ Smem3 carve3(const Graph2& g, float4* base, int CB, bool with_ptr) {
  Smem3 s;
  const int nb = g.db3 ? 2 : 1;
  s.stride_t = g.db3 ? g.R3 * CB : 0;
  s.stride_h = g.db3 ? g.R3 : 0;
  s.tile = base;
  s.tile2 = base + nb * g.R3 * CB;
  s.halo = reinterpret_cast<float*>(base + 2 * nb * g.R3 * CB);
  s.halo2 = s.halo + nb * g.R3;
  s.ext = reinterpret_cast<int*>(s.halo2 + g.R3);
  s.wself = reinterpret_cast<float*>(s.ext + g.R3);
  s.ptr = reinterpret_cast<int*>(s.wself + g.NT3 + (((nb + 2) * g.R3 + g.NT3) & 1));      // keeps the entries 8-byte aligned
  const int np = with_ptr ? ((g.NT3 + 2) & ~1) : 0;
  s.ord = s.ptr + np;
  s.tab = reinterpret_cast<int2*>(s.ord + (with_ptr ? ((g.NT3 + 1) & ~1) : 0));
  return s;
}

// One unit of work: window b, time tile starting at chunk c0, node tile [n0, n0 + nt) with nh external rows.
// Single-tile plans: id -> (window, time tile), all nodes.  Node-tiled plans: id -> (node tile, window, time tile).
struct Item3 {
  int b, c0, j, n0, nt, nh;
};
__device__ __forceinline__ Item3 item3(const Graph2& g, int64_t B, int id, const int* __restrict__ extp) {
  Item3 t;
  if (g.ntile3 > 1) {
    const int per = (int)B * g.tiles3;          // node-tile major: (node tile, window, time tile)
    t.j = id / per;
    const int rem = id - t.j * per;
    t.b = rem / g.tiles3;
    t.c0 = (rem - t.b * g.tiles3) * (int)blockDim.x;
    t.n0 = t.j * g.NT3;
    t.nt = min(g.NT3, g.N - t.n0);
    t.nh = extp[t.j + 1] - extp[t.j];
  } else {
    t.b = id / g.tiles3;
    t.c0 = (id - t.b * g.tiles3) * (int)blockDim.x;
    t.j = 0; t.n0 = 0; t.nt = g.N; t.nh = 0;
  }
  return t;
}
// the work of a CTA: single-tile plans walk the items round-robin (neighbouring time tiles run side by side), node-tiled
// plans give every CTA a contiguous range, so that its node tile (and the staged table) changes at most once or twice
__device__ __forceinline__ void range3(const Graph2& g, int total, int* first, int* last, int* stride) {
  if (g.ntile3 > 1) {
    const int per = (total + gridDim.x - 1) / gridDim.x;
    *first = blockIdx.x * per; *last = min(total, *first + per); *stride = 1;
  } else {
    *first = blockIdx.x; *last = total; *stride = gridDim.x;
  }
}
// global row of local row n of the tile
__device__ __forceinline__ int grow3(const Item3& t, const Smem3& s, int n) { return n < t.nt ? t.n0 + n : s.ext[n - t.nt]; }

// asynchronous global -> shared copies (LDGSTS): the whole tile is in flight at once and costs no registers
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// SRC 0: v = r + beta p   SRC 1: v = r (first iteration)   SRC 2: v = r (no store: r is x0 of the initial residual)
// ncu on the register-staged version: every dependent round of global loads costs the loaded HBM latency (~2 us);
// a tile took 3 such rounds.  Here ALL chunks of the tile (and the halo scalars) are issued as cp.async before
// anything waits: one round per tile.  Each thread then finishes its OWN chunks (p' = r + beta p in place, streamed
// to HBM for the tile's own nodes), so no barrier is needed between the copy and that pass.
// Double-buffered mode (g.db3, one CTA per SM): the copies of the CTA's NEXT tile are issued before the gathers of
// the current one, so the HBM stream never stops; single-buffered mode (two CTAs per SM): issue, wait, gather.
template <int SRC>
__device__ __forceinline__ void tile_issue(const Graph2& g, const Item3& t, int cur, const float* __restrict__ r,
                                           const float* __restrict__ p_old, const Smem3& s, bool want_halo) {
  const int CB = blockDim.x, NBt = blockDim.y, tx = threadIdx.x;
  const int c = t.c0 + tx, cn = t.c0 + CB;
  if (c < g.C4) {
    const bool last = want_halo && tx == CB - 1 && cn < g.C4;
    const size_t w0 = (size_t)t.b * (size_t)(g.N * g.C4) + c;
    const float4* rw = reinterpret_cast<const float4*>(r) + w0;
    const float4* pw = reinterpret_cast<const float4*>(p_old) + w0;
    float4* t1 = s.tile + cur * s.stride_t + tx;
    float4* t2 = s.tile2 + cur * s.stride_t + tx;
    float* h1 = s.halo + cur * s.stride_h;
    for (int n = threadIdx.y; n < t.nt + t.nh; n += NBt) {
      const int gr = grow3(t, s, n) * g.C4;
      cp_async16(t1 + n * CB, rw + gr);
      if (SRC == 0) cp_async16(t2 + n * CB, pw + gr);
      if (last) {                 // first element of the next time tile's first chunk
        cp_async4(h1 + n, reinterpret_cast<const float*>(rw + gr + 1));
        if (SRC == 0) cp_async4(s.halo2 + n, reinterpret_cast<const float*>(pw + gr + 1));
      }
    }
  }
  cp_async_commit();
}

// xd_in / xd_out (SRC 0): the deferred x update of the previous iteration, x_out = x_in + alpha(it - 1) p_old for the tile's OWN
// rows - p_old is staged here anyway, and k2_xr then only updates r (see k4_cg, mga_stream4.cuh).  The thread's chunks of x are
// requested before the wait for the tile's copies, so both round trips overlap; at most kXR3 own rows per thread (host check).
constexpr int kXR3 = 8;
template <int SRC>
__device__ __forceinline__ void tile_finish(const Graph2& g, int64_t B, int it, const Item3& t, int cur, float* __restrict__ p_new,
                                            const double* __restrict__ dots, const Smem3& s, bool want_halo,
                                            const float* xd_in = nullptr, float* xd_out = nullptr) {
  const int CB = blockDim.x, NBt = blockDim.y, tx = threadIdx.x;
  float4 xv[kXR3];
  const bool xdef = SRC == 0 && xd_out != nullptr && t.c0 + tx < g.C4;
  if (xdef) {
    const float4* xi = reinterpret_cast<const float4*>(xd_in) + (size_t)t.b * (size_t)(g.N * g.C4) + (size_t)t.n0 * g.C4 + t.c0 + tx;
#pragma unroll
    for (int j = 0; j < kXR3; ++j) {
      const int n = threadIdx.y + j * NBt;
      if (n < t.nt) xv[j] = __ldcs(xi + (size_t)n * g.C4);
    }
  }
  const int c = t.c0 + tx, cn = t.c0 + CB;
  const bool cok = c < g.C4, last = want_halo && tx == CB - 1, hok = cn < g.C4;
  float4* ow = reinterpret_cast<float4*>(p_new) + (size_t)t.b * (size_t)(g.N * g.C4) + (cok ? c : 0);
  float4* t1 = s.tile + cur * s.stride_t + tx;
  const float4* t2 = s.tile2 + cur * s.stride_t + tx;
  float* h1 = s.halo + cur * s.stride_h;
  float beta = 0.f;
  if (SRC == 0) beta = (float)dots[(size_t)(2 * it) * B + t.b] / (float)dots[(size_t)(2 * it - 2) * B + t.b];   // ADMM.py:356
  cp_async_wait_all();
  for (int n = threadIdx.y; n < t.nt + t.nh; n += NBt) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cok) {
      v = t1[n * CB];
      if (SRC == 0) {
        const float4 q = t2[n * CB];
        v = make_float4(v.x + beta * q.x, v.y + beta * q.y, v.z + beta * q.z, v.w + beta * q.w);
      }
      if (SRC != 2 && n < t.nt) __stcs(ow + (t.n0 + n) * g.C4, v);      // external rows are stored by their own tile
    }
    if (SRC == 0 || !cok) t1[n * CB] = v;
    if (last) h1[n] = hok ? (SRC == 0 ? h1[n] + beta * s.halo2[n] : h1[n]) : 0.f;
  }
  if (xdef) {
    const float al = (float)dots[(size_t)(2 * it - 2) * B + t.b] / (float)dots[(size_t)(2 * it - 1) * B + t.b];   // as k2_xr forms it (ADMM.py:351)
    float4* xo = reinterpret_cast<float4*>(xd_out) + (size_t)t.b * (size_t)(g.N * g.C4) + (size_t)t.n0 * g.C4 + c;
#pragma unroll
    for (int j = 0; j < kXR3; ++j) {
      const int n = threadIdx.y + j * NBt;
      if (n < t.nt) {
        const float4 q = t2[n * CB];
        float4 u = xv[j];
        u.x += al * q.x; u.y += al * q.y; u.z += al * q.z; u.w += al * q.w;
        __stcs(xo + (size_t)n * g.C4, u);
      }
    }
  }
}

// acc += sum_j w_j tile[nbr_j][tx]: entries (byte offset of the neighbour's row in the tile, weight), `mine` = the
// thread's chunk column in the tile; K > 0: compile-time table width
template <int K>
__device__ __forceinline__ float4 gather3(const int2* row, int slots, const char* mine, float4 acc) {
#ifdef MGA_S3_NOGATHER      // timing experiment only: what the tile traffic alone costs
  return acc;
#endif
  if (K > 0) {
    int2 e[K > 0 ? K : 1];
#pragma unroll
    for (int j = 0; j < K; ++j) e[j] = row[j];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float wj = __int_as_float(e[j].y);
      const float4 a = *reinterpret_cast<const float4*>(mine + e[j].x);
      acc.x += wj * a.x; acc.y += wj * a.y; acc.z += wj * a.z; acc.w += wj * a.w;
    }
  } else {
#pragma unroll 2
    for (int j = 0; j < slots; ++j) {
      const int2 e = row[j];
      const float wj = __int_as_float(e.y);
      const float4 a = *reinterpret_cast<const float4*>(mine + e.x);
      acc.x += wj * a.x; acc.y += wj * a.y; acc.z += wj * a.z; acc.w += wj * a.w;
    }
  }
  return acc;
}

// the tile's slice of a forward table, its external-row list and (optionally) the self weights -> shared memory
__device__ __forceinline__ void stage3(const Graph2& g, const Item3& t, const Smem3& s, const int2* __restrict__ tab, int K,
                                       const int* __restrict__ extp, const int* __restrict__ ext, bool want_self) {
  const int tid = tid2(), nt = blockDim.x * blockDim.y;
  const int2* src = tab + (size_t)t.n0 * K;
  for (int k = tid; k < t.nt * K; k += nt) s.tab[k] = src[k];
  if (t.nh > 0) {
    const int* e = ext + extp[t.j];
    for (int k = tid; k < t.nh; k += nt) s.ext[k] = e[k];
  }
  if (want_self)
    for (int k = tid; k < t.nt; k += nt) s.wself[k] = g.wself_d[t.n0 + k];
}

}  // namespace mga
#include "mga_stream4.cuh"
namespace mga {

// (r, p) -> p', qs = shifted L_d p'   [SRC 2: x0 -> qs]
template <int SRC, int K>
__global__ void __launch_bounds__(1024, 1) k3_p_ldr(Graph2 g, int64_t B, int it, const float* __restrict__ r,
                                                   const float* __restrict__ p_old, float* __restrict__ p_new,
                                                   float* __restrict__ qs, const double* __restrict__ dots,
                                                   const float* xd_in, float* xd_out) {
  extern __shared__ float4 s3[];
  const int CB = blockDim.x, NBt = blockDim.y, tx = threadIdx.x;
  const Smem3 s = carve3(g, s3, CB, false);
  const int total = (int)B * g.tiles3 * g.ntile3;
  const bool db = g.db3 != 0;
  int first, last, stride;
  range3(g, total, &first, &last, &stride);
  int cur = 0, staged = -1;
  if (first < last) {
    const Item3 t0 = item3(g, B, first, g.extp_d);
    stage3(g, t0, s, g.tab_d, g.kd3, g.extp_d, g.ext_d, true);
    staged = t0.j;
    if (t0.nh > 0) __syncthreads();                      // the external-row list is read by the copies
    if (db) tile_issue<SRC>(g, t0, 0, r, p_old, s, true);
  }
  for (int tl = first; tl < last; tl += stride, cur ^= (db ? 1 : 0)) {
    const Item3 t = item3(g, B, tl, g.extp_d);
    const int c = t.c0 + tx;
    if (!db) {
      if (tl != first) __syncthreads();                  // the previous tile's gathers are done
      if (t.j != staged) {
        stage3(g, t, s, g.tab_d, g.kd3, g.extp_d, g.ext_d, true);
        staged = t.j;
        __syncthreads();
      }
      tile_issue<SRC>(g, t, 0, r, p_old, s, true);
    }
    tile_finish<SRC>(g, B, it, t, cur, p_new, dots, s, true, xd_in, xd_out);
    __syncthreads();
    if (db && tl + stride < last) tile_issue<SRC>(g, item3(g, B, tl + stride, g.extp_d), cur ^ 1, r, p_old, s, true);
    if (c >= g.C4) continue;
    const float4* tile = s.tile + cur * s.stride_t;
    const char* mine = reinterpret_cast<const char*>(tile + tx);
    // the element after the thread's chunk: next chunk of the row, or the halo scalar for the tile's last chunk
    const float* nxt_p = tx + 1 < CB ? reinterpret_cast<const float*>(tile + tx + 1) : s.halo + cur * s.stride_h;
    const int nxt_stride = tx + 1 < CB ? CB * 4 : 1;
    float4* qw = reinterpret_cast<float4*>(qs) + (size_t)t.b * (size_t)(g.N * g.C4) + (size_t)t.n0 * g.C4 + c;
    const int tt = 4 * c;
    const bool v1 = tt + 1 < g.T, v2 = tt + 2 < g.T, v3 = tt + 3 < g.T, v4 = tt + 4 < g.T;
    for (int n = threadIdx.y; n < t.nt; n += NBt) {
      const float4 own = tile[n * CB + tx];
      const float nxt = nxt_p[n * nxt_stride];
      const float ws = s.wself[n];
      const float4 acc = gather3<K>(s.tab + n * g.kd3, g.kd3, mine,
                                    make_float4(ws * own.x, ws * own.y, ws * own.z, ws * own.w));
      float4 o;
      o.x = v1 ? own.y - acc.x : 0.f;
      o.y = v2 ? own.z - acc.y : 0.f;
      o.z = v3 ? own.w - acc.z : 0.f;
      o.w = v4 ? nxt - acc.w : 0.f;
      __stcs(qw + n * g.C4, o);
    }
  }
}

// (v, qs) -> Ap, <v, Ap>  [MODE 1: r = rhs - A v, <r, r>];  A = diag + c L_d^T L_d; qs tile in shared memory
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k3_ldrt_lhs(Graph2 g, int64_t B, const float* __restrict__ v,
                                                      const float* __restrict__ qs, const float* __restrict__ rhs,
                                                      float* __restrict__ out, double* __restrict__ slot, float a, float cc,
                                                      int xsys) {
  extern __shared__ float4 s3[];
  const int CB = blockDim.x, NBt = blockDim.y, tx = threadIdx.x;
  const Smem3 s = carve3(g, s3, CB, true);
  const int total = (int)B * g.tiles3 * g.ntile3;
  const bool db = g.db3 != 0;
  const bool self_in = g.in_self3 != 0;
  // the tile's in-list (offsets rebased to the tile), row order, self weights and external rows -> shared memory
  auto stage = [&](const Item3& t) {
    const int tid = tid2(), nthr = CB * NBt;
    const int e0 = g.in_ptr3[t.n0];
    for (int k = tid; k <= t.nt; k += nthr) s.ptr[k] = g.in_ptr3[t.n0 + k] - e0;
    for (int k = tid; k < t.nt; k += nthr) { s.ord[k] = g.ord3[t.n0 + k] - t.n0; s.wself[k] = g.wself_d[t.n0 + k]; }
    const int ne = g.in_ptr3[t.n0 + t.nt] - e0;
    for (int k = tid; k < ne; k += nthr) s.tab[k] = g.tab_in3[e0 + k];
    const int* ex = g.ext_in + g.extp_in[t.j];
    for (int k = tid; k < t.nh; k += nthr) s.ext[k] = ex[k];
  };
  // all copies of a tile into buffer `k`: qs -> tile (own + external rows), v -> tile2 (own rows),
  // q[4 c0] = qs[4 c0 - 1] -> halo (q[0] = 0, ADMM.py:176)
  auto issue = [&](const Item3& t, int k) {
    const int ccx = t.c0 + tx;
    float4* t1 = s.tile + k * s.stride_t + tx;
    float4* t2 = s.tile2 + k * s.stride_t + tx;
    float* h1 = s.halo + k * s.stride_h;
    if (ccx < g.C4) {
      const size_t wq = (size_t)t.b * (size_t)(g.N * g.C4) + ccx;
      const float4* qw = reinterpret_cast<const float4*>(qs) + wq;
      const float4* vw = reinterpret_cast<const float4*>(v) + wq;
      for (int n = threadIdx.y; n < t.nt + t.nh; n += NBt) {
        const int gr = grow3(t, s, n) * g.C4;
        cp_async16(t1 + n * CB, qw + gr);
        if (n < t.nt) {
          cp_async16(t2 + n * CB, vw + gr);
          if (tx == 0 && t.c0 > 0) cp_async4(h1 + n, reinterpret_cast<const float*>(qw + gr) - 1);
        }
      }
    } else {
      for (int n = threadIdx.y; n < t.nt + t.nh; n += NBt) t1[n * CB] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tx == 0 && t.c0 == 0)
      for (int n = threadIdx.y; n < t.nt; n += NBt) h1[n] = 0.f;
    cp_async_commit();
  };
  int first, last, stride;
  range3(g, total, &first, &last, &stride);
  int cur = 0, staged = -1;
  if (first < last) {
    const Item3 t0 = item3(g, B, first, g.extp_in);
    stage(t0);
    staged = t0.j;
    if (t0.nh > 0) __syncthreads();
    if (db) issue(t0, 0);
  }
  for (int tl = first; tl < last; tl += stride, cur ^= (db ? 1 : 0)) {
    const Item3 t = item3(g, B, tl, g.extp_in);
    const int c = t.c0 + tx;
    const bool cok = c < g.C4;
    const size_t w0 = (size_t)t.b * (size_t)(g.N * g.C4) + (size_t)t.n0 * g.C4;
    if (!db) {
      if (tl != first) __syncthreads();
      if (t.j != staged) {
        stage(t);
        staged = t.j;
        __syncthreads();
      }
      issue(t, 0);
    }
    cp_async_wait_all();
    __syncthreads();
    if (db && tl + stride < last) issue(item3(g, B, tl + stride, g.extp_in), cur ^ 1);
    const float4* tile = s.tile + cur * s.stride_t;
    const float4* tile2 = s.tile2 + cur * s.stride_t;
    const char* mine = reinterpret_cast<const char*>(tile + tx);
    // the element before the thread's chunk: q[4c] = qs[4c - 1]
    const float* prv_p = tx > 0 ? reinterpret_cast<const float*>(tile + tx) - 1 : s.halo + cur * s.stride_h;
    const int prv_stride = tx > 0 ? CB * 4 : 1;
    float dot = 0.f;
    if (cok) {
      const int t0 = 4 * c;
      const float4* rw = reinterpret_cast<const float4*>(rhs) + w0 + c;
      float4* ow = reinterpret_cast<float4*>(out) + w0 + c;
      float hx[4], tv[4];      // H^T H keeps rows t < t_in (ADMM.py:372-374); pads (t >= T) stay 0
#pragma unroll
      for (int j = 0; j < 4; ++j) { hx[j] = (xsys && t0 + j < g.t_in) ? 1.f : 0.f; tv[j] = t0 + j < g.T ? 1.f : 0.f; }
      for (int k = threadIdx.y; k < t.nt; k += NBt) {
        const int n = s.ord[k];
        const float4 pv = tile2[n * CB + tx];
        float4 rh;
        if (MODE == 1) rh = __ldcs(rw + n * g.C4);
        const float4 q1 = tile[n * CB + tx];
        const float qprev = prv_p[n * prv_stride];
        const float ws = self_in ? s.wself[n] : 0.f;
        float4 f = make_float4(ws * q1.x, ws * q1.y, ws * q1.z, ws * q1.w);     // self link: w_self qs_own
        int e = s.ptr[n];
        const int e1 = s.ptr[n + 1];
        for (; e + 2 <= e1; e += 2) f = gather3<2>(s.tab + e, 2, mine, f);
        if (e < e1) f = gather3<1>(s.tab + e, 1, mine, f);
        const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
        const float q[4] = {qprev, q1.x, q1.y, q1.z};
        const float ff[4] = {f.x, f.y, f.z, f.w};
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float l = q[j] - ff[j];
          float val;
          if (xsys) val = (hx[j] * pp[j] + a * pp[j]) + cc * l;      // ADMM.py:372-379
          else val = cc * l + a * pp[j];                            // ADMM.py:394
          o[j] = tv[j] * val;
        }
        if (MODE == 0) {
          __stcs(ow + n * g.C4, make_float4(o[0], o[1], o[2], o[3]));
          dot += (pp[0] * o[0] + pp[1] * o[1]) + (pp[2] * o[2] + pp[3] * o[3]);
        } else {
          const float r0 = rh.x - o[0], r1 = rh.y - o[1], r2 = rh.z - o[2], r3 = rh.w - o[3];
          __stcs(ow + n * g.C4, make_float4(r0, r1, r2, r3));
          dot += (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
        }
      }
    }
    block_add(dot, slot + t.b, tid2(), blockDim.x * blockDim.y);
  }
}

// z_u system: (r, p) -> p', Ap = (c L_u + a I) p', <p', Ap>   [SRC 2 / MODE 1: r = rhs - A x0, <r, r>]
template <int SRC, int MODE, int K>
__global__ void __launch_bounds__(1024, 1) k3_lu(Graph2 g, int64_t B, int it, const float* __restrict__ r,
                                                const float* __restrict__ p_old, float* __restrict__ p_new,
                                                const float* __restrict__ rhs, float* __restrict__ out,
                                                const double* __restrict__ dots, double* __restrict__ slot, float a, float cc,
                                                const float* xd_in, float* xd_out) {
  extern __shared__ float4 s3[];
  const int CB = blockDim.x, NBt = blockDim.y, tx = threadIdx.x;
  const Smem3 s = carve3(g, s3, CB, false);
  const int total = (int)B * g.tiles3 * g.ntile3;
  const bool db = g.db3 != 0;
  int first, last, stride;
  range3(g, total, &first, &last, &stride);
  int cur = 0, staged = -1;
  if (first < last) {
    const Item3 t0 = item3(g, B, first, g.extp_u);
    stage3(g, t0, s, g.tab_u, g.ku3, g.extp_u, g.ext_u, false);
    staged = t0.j;
    if (t0.nh > 0) __syncthreads();
    if (db) tile_issue<SRC>(g, t0, 0, r, p_old, s, false);
  }
  for (int tl = first; tl < last; tl += stride, cur ^= (db ? 1 : 0)) {
    const Item3 t = item3(g, B, tl, g.extp_u);
    const int c = t.c0 + tx;
    if (!db) {
      if (tl != first) __syncthreads();
      if (t.j != staged) {
        stage3(g, t, s, g.tab_u, g.ku3, g.extp_u, g.ext_u, false);
        staged = t.j;
        __syncthreads();
      }
      tile_issue<SRC>(g, t, 0, r, p_old, s, false);
    }
    tile_finish<SRC>(g, B, it, t, cur, p_new, dots, s, false, xd_in, xd_out);
    __syncthreads();
    if (db && tl + stride < last) tile_issue<SRC>(g, item3(g, B, tl + stride, g.extp_u), cur ^ 1, r, p_old, s, false);
    const float4* tile = s.tile + cur * s.stride_t;
    const char* mine = reinterpret_cast<const char*>(tile + tx);
    float dot = 0.f;
    if (c < g.C4) {
      const size_t w0 = (size_t)t.b * (size_t)(g.N * g.C4) + (size_t)t.n0 * g.C4;
      const float4* rw = reinterpret_cast<const float4*>(rhs) + w0 + c;
      float4* ow = reinterpret_cast<float4*>(out) + w0 + c;
      for (int n = threadIdx.y; n < t.nt; n += NBt) {
        float4 rh;
        if (MODE == 1) rh = __ldcs(rw + n * g.C4);
        const float4 pv = tile[n * CB + tx];
        const float4 acc = gather3<K>(s.tab + n * g.ku3, g.ku3, mine, make_float4(0.f, 0.f, 0.f, 0.f));
        float4 o;
        o.x = cc * (pv.x - acc.x) + a * pv.x;
        o.y = cc * (pv.y - acc.y) + a * pv.y;
        o.z = cc * (pv.z - acc.z) + a * pv.z;
        o.w = cc * (pv.w - acc.w) + a * pv.w;    // pads: v = 0 and every gathered pad is 0
        if (MODE == 0) {
          __stcs(ow + n * g.C4, o);
          dot += (pv.x * o.x + pv.y * o.y) + (pv.z * o.z + pv.w * o.w);
        } else {
          const float4 rr = make_float4(rh.x - o.x, rh.y - o.y, rh.z - o.z, rh.w - o.w);
          __stcs(ow + n * g.C4, rr);
          dot += (rr.x * rr.x + rr.y * rr.y) + (rr.z * rr.z + rr.w * rr.w);
        }
      }
    }
    block_add(dot, slot + t.b, tid2(), blockDim.x * blockDim.y);
  }
}

// alpha / beta live in arrays of B_out windows per iteration (this call's B windows are a slice of them)
__global__ void k2_coeffs(int64_t B, int64_t B_out, int iters, const double* __restrict__ dots, float* __restrict__ alpha,
                          float* __restrict__ beta) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * iters) return;
  const int k = (int)(idx / B);
  const int64_t b = idx - (int64_t)k * B;
  idx = (int64_t)k * B_out + b;
  const float rr = (float)dots[(size_t)(2 * k) * B + b], pap = (float)dots[(size_t)(2 * k + 1) * B + b];
  const float rrn = (float)dots[(size_t)(2 * k + 2) * B + b];
  if (alpha) alpha[idx] = rr / pap;
  if (beta) beta[idx] = rrn / rr;
}

// ---- elementwise steps of combined_loop ---------------------------------------------------------------
// (L_d v)[t] for the 4 time steps of chunk (n, c): unaligned in time, scalar gathers (once per outer iteration)
__device__ __forceinline__ void ldr_chunk(const Graph2& g, const float* __restrict__ v, size_t w0, int n, int c, const float4 own,
                                          float (&out)[4]) {
  const float o[4] = {own.x, own.y, own.z, own.w};
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int* nb = g.nbr_d + (size_t)n * g.kd;
  const float* ww = g.w_d + (size_t)n * g.kd;
  const int t0 = 4 * c;
  const float4* vw = reinterpret_cast<const float4*>(v) + w0 + c;
  for (int j = 0; j < g.kd; ++j) {
    const int m = nb[j];
    if (m < 0) continue;
    const float wj = ww[j];
    // the neighbour's steps t0-1 .. t0+2: its chunk c (one 128-bit load) and the last element of chunk c-1; pads are 0
    const float4 a = vw[(size_t)m * g.C4];
    const float prev = c > 0 ? reinterpret_cast<const float*>(vw + (size_t)m * g.C4)[-1] : 0.f;
    acc[0] += wj * prev; acc[1] += wj * a.x; acc[2] += wj * a.y; acc[3] += wj * a.z;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int t = t0 + q;
    out[q] = (t >= 1 && t < g.T) ? o[q] - acc[q] : 0.f;
  }
}

__global__ void __launch_bounds__(kB2) k2_ldr(Graph2 g, const float* __restrict__ v, float* __restrict__ out) {
  const Chunk k = locate2(g);
  if (!k.ok) return;
  float o[4];
  ldr_chunk(g, v, (size_t)k.b * g.N * g.C4, k.n, k.c, ld4(v, k.g), o);
  st4(out, k.g, make_float4(o[0], o[1], o[2], o[3]));
}

// RHS_x (ADMM.py:552-559): Ldr_T(gamma + rho phi)/2 + (rho_u zu + rho_d zd)/2 - (gu + gd)/2 + H^T y
__global__ void __launch_bounds__(kB2) k2_rhs_x(Graph2 g, const float* __restrict__ gam, const float* __restrict__ phi,
                                                const float* __restrict__ zu, const float* __restrict__ zd,
                                                const float* __restrict__ gu, const float* __restrict__ gd,
                                                const float* __restrict__ y, float* __restrict__ rhs, float rho, float rho_u,
                                                float rho_d) {
  const Chunk k = locate2(g);
  if (!k.ok) return;
  const size_t w0 = (size_t)k.b * g.N * g.C4;
  const float4 ga = ld4(gam, k.g), ph = ld4(phi, k.g);
  const float v[4] = {ga.x + rho * ph.x, ga.y + rho * ph.y, ga.z + rho * ph.z, ga.w + rho * ph.w};
  float f[4] = {0.f, 0.f, 0.f, 0.f};
  const int t0 = 4 * k.c;
  const float4* gw = reinterpret_cast<const float4*>(gam) + w0 + k.c;
  const float4* pw = reinterpret_cast<const float4*>(phi) + w0 + k.c;
  const bool more = k.c + 1 < g.C4;
  for (int e = g.in_ptr[k.n]; e < g.in_ptr[k.n + 1]; ++e) {
    const float w = g.in_w[e];
    // the source's steps t0+1 .. t0+4: its chunk c and the first element of chunk c+1 (pads of both vectors are 0)
    const size_t row = (size_t)g.in_src[e] * g.C4;
    const float4 ga2 = gw[row], ph2 = pw[row];
    const float gn = more ? reinterpret_cast<const float*>(gw + row + 1)[0] : 0.f;
    const float pn = more ? reinterpret_cast<const float*>(pw + row + 1)[0] : 0.f;
    f[0] += w * (ga2.y + rho * ph2.y);
    f[1] += w * (ga2.z + rho * ph2.z);
    f[2] += w * (ga2.w + rho * ph2.w);
    f[3] += w * (gn + rho * pn);
  }
  const float4 a = ld4(zu, k.g), b = ld4(zd, k.g), c = ld4(gu, k.g), d = ld4(gd, k.g);
  const float zuv[4] = {a.x, a.y, a.z, a.w}, zdv[4] = {b.x, b.y, b.z, b.w};
  const float guv[4] = {c.x, c.y, c.z, c.w}, gdv[4] = {d.x, d.y, d.z, d.w};
  const int o = g.perm[k.n];
  float out[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int t = t0 + q;
    // rows of apply_op_Ldr_T: t = T-1 keeps v; t = 0 keeps the identity term only under Q1 (ADMM.py:217-223)
    const float l = (t == g.T - 1) ? v[q] : ((t == 0 && !g.q1) ? -f[q] : v[q] - f[q]);
    const float hty = t < g.t_in ? y[((size_t)k.b * g.t_in + t) * g.N + o] : 0.f;
    out[q] = t < g.T ? l / 2.f + (rho_u * zuv[q] + rho_d * zdv[q]) / 2.f - (guv[q] + gdv[q]) / 2.f + hty : 0.f;
  }
  st4(rhs, k.g, make_float4(out[0], out[1], out[2], out[3]));
}

// RHS_zu / RHS_zd (ADMM.py:579, 587)
__global__ void __launch_bounds__(kFlat) k2_rhs_z(size_t chunks, const float* __restrict__ gz, const float* __restrict__ x,
                                                float* __restrict__ rhs, float half_rho) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= chunks) return;
  const float4 g = ld4(gz, k), xv = ld4(x, k);
  st4(rhs, k, make_float4(g.x / 2.f + half_rho * xv.x, g.y / 2.f + half_rho * xv.y, g.z / 2.f + half_rho * xv.z,
                          g.w / 2.f + half_rho * xv.w));
}

__device__ __forceinline__ float soft2(float s, float d) {
  const float u = fabsf(s) - d;
  const float sg = (float)((s > 0.f) - (s < 0.f));
  return sg * u * (float)(u > 0.f);   // ADMM.py:407-408
}

// The tail of one outer iteration in a single pass (ADMM.py:595-637): both dual ascents, phi prox, gamma
// ascent and all diagnostics.
__global__ void __launch_bounds__(kB2) k2_tail(Graph2 g, int want_diag, const float* __restrict__ x,
                                               const float* __restrict__ x_old, const float* __restrict__ zu,
                                               const float* __restrict__ zu_old, const float* __restrict__ zd,
                                               const float* __restrict__ zd_old, float* __restrict__ gu, float* __restrict__ gd,
                                               float* __restrict__ gam, float* __restrict__ phi, const float* __restrict__ y,
                                               float rho, float rho_u, float rho_d, float thr, double* __restrict__ diag,
                                               double* __restrict__ dx_sum) {
  const Chunk k = locate2(g);
  float d[MGA_DIAG_COLS];
#pragma unroll
  for (int c = 0; c < MGA_DIAG_COLS; ++c) d[c] = 0.f;
  if (k.ok) {
    const size_t w0 = (size_t)k.b * g.N * g.C4;
    const float4 xv4 = ld4(x, k.g), zu4 = ld4(zu, k.g), zd4 = ld4(zd, k.g), gu4 = ld4(gu, k.g), gd4 = ld4(gd, k.g);
    const float4 ga4 = ld4(gam, k.g), ph4 = ld4(phi, k.g);
    const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w}, zuv[4] = {zu4.x, zu4.y, zu4.z, zu4.w};
    const float zdv[4] = {zd4.x, zd4.y, zd4.z, zd4.w};
    float guv[4] = {gu4.x, gu4.y, gu4.z, gu4.w}, gdv[4] = {gd4.x, gd4.y, gd4.z, gd4.w};
    float gav[4] = {ga4.x, ga4.y, ga4.z, ga4.w}, phv[4] = {ph4.x, ph4.y, ph4.z, ph4.w};
    float ldx[4];
    ldr_chunk(g, x, w0, k.n, k.c, xv4, ldx);
    float lux[4] = {0.f, 0.f, 0.f, 0.f}, xo[4] = {0.f, 0.f, 0.f, 0.f}, zuo[4] = {0.f, 0.f, 0.f, 0.f}, zdo[4] = {0.f, 0.f, 0.f, 0.f};
    if (want_diag) {
      const float4 acc = fwd_gather(g.nbr_u, g.w_u, g.ku, win4(x, g, k.b), k.n, k.c, g.C4);
      lux[0] = xv[0] - acc.x; lux[1] = xv[1] - acc.y; lux[2] = xv[2] - acc.z; lux[3] = xv[3] - acc.w;
      const float4 a = ld4(x_old, k.g), b = ld4(zu_old, k.g), c = ld4(zd_old, k.g);
      xo[0] = a.x; xo[1] = a.y; xo[2] = a.z; xo[3] = a.w;
      zuo[0] = b.x; zuo[1] = b.y; zuo[2] = b.z; zuo[3] = b.w;
      zdo[0] = c.x; zdo[1] = c.y; zdo[2] = c.z; zdo[3] = c.w;
    }
    const int o = g.perm[k.n];
    const int t0 = 4 * k.c;
    int bad = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = t0 + q;
      if (t >= g.T) continue;
      guv[q] = guv[q] + rho_u * (xv[q] - zuv[q]);
      gdv[q] = gdv[q] + rho_d * (xv[q] - zdv[q]);
      const float ph = soft2(ldx[q] - gav[q] / rho, thr);
      const float gn = gav[q] + rho * (ph - ldx[q]);
      bad |= !isfinite(xv[q]) || !isfinite(zuv[q]) || !isfinite(zdv[q]) || !isfinite(ph) || !isfinite(gn);
      if (want_diag) {
        const float dx = xv[q] - xo[q];
        d[MGA_DIAG_DX2] += dx * dx;
        if (dx_sum) atomicAdd(dx_sum + (size_t)t * g.N + o, (double)dx);
        const float a = xv[q] - zuv[q], bz = zuv[q] - zuo[q];
        d[MGA_DIAG_X_ZU2] += a * a;
        d[MGA_DIAG_DZU2] += bz * bz;
        d[MGA_DIAG_GLR] += xv[q] * lux[q];
        if (t < g.t_in) {
          const float h = xv[q] - y[((size_t)k.b * g.t_in + t) * g.N + o];
          d[MGA_DIAG_RECOVER2] += h * h;
        }
        const float e = ph - ldx[q], f = ph - phv[q];
        d[MGA_DIAG_PHI_LDX2] += e * e;
        d[MGA_DIAG_DPHI2] += f * f;
        d[MGA_DIAG_DGTV] += fabsf(ldx[q]);
        const float e2 = xv[q] - zdv[q], f2 = zdv[q] - zdo[q];
        d[MGA_DIAG_X_ZD2] += e2 * e2;
        d[MGA_DIAG_DZD2] += f2 * f2;
        d[MGA_DIAG_DGLR] += ldx[q] * ldx[q];
      }
      phv[q] = ph;
      gav[q] = gn;
    }
    d[MGA_DIAG_NONFINITE] = (float)bad;
    st4(gu, k.g, make_float4(guv[0], guv[1], guv[2], guv[3]));
    st4(gd, k.g, make_float4(gdv[0], gdv[1], gdv[2], gdv[3]));
    st4(gam, k.g, make_float4(gav[0], gav[1], gav[2], gav[3]));
    st4(phi, k.g, make_float4(phv[0], phv[1], phv[2], phv[3]));
  }
  __shared__ float red[MGA_DIAG_COLS][kB2 / 32];
  const int tid = tid2(), lane = tid & 31, w = tid >> 5;
#pragma unroll
  for (int c = 0; c < MGA_DIAG_COLS; ++c) {
    const float v = warp_sum<float>(d[c]);
    if (lane == 0) red[c][w] = v;
  }
  __syncthreads();
  if (tid < MGA_DIAG_COLS) {
    float t = 0.f;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    for (int q = 0; q < nw; ++q) t += red[tid][q];
    if (t != 0.f) atomicAdd(diag + tid, (double)t);
  }
}

// ---- host side -----------------------------------------------------------------------------------------
void stream2_tiling(Graph2* g) {
  g->CB = std::min(g->C4, 8);
  int step = 32;                                        // NBt must make CB * NBt whole warps
  for (int d = 2; d <= g->CB; d *= 2) if (g->CB % d == 0) step = 32 / d;
  g->NBt = std::max(step, (kB2 / g->CB) / step * step); // thread rows per CTA, CB * NBt <= kB2
  g->tilesN = (g->N + g->NBt - 1) / g->NBt;
  g->NB = (g->N + g->tilesN - 1) / g->tilesN;           // balanced node tiles, NB <= NBt (extra rows are masked off)
  g->tilesC = (g->C4 + g->CB - 1) / g->CB;
  // time-tiled shared-memory kernels (k3_*): all nodes x CB3 chunks + halo + self weights + the graph table per CTA;
  g->CB3 = 0;
  int force = 0;
  if (const char* e = std::getenv("MGA_S3_CB")) force = std::atoi(e);     // 0 = automatic, < 0 = off, > 0 = chunks per tile
  if (force >= 0) {
    // tiles of >= 4 chunks (64-byte row segments) or the whole row.  Single-buffered (default): 2 tile buffers, two CTAs
    // per SM (N <= ~360 with 8-chunk tiles, ~590 with 4-chunk tiles), or one CTA per SM (8-chunk tiles, N <= ~690).  Double-buffered (MGA_S3_DB=1): 4 buffers, one CTA
    // per SM, the next tile's copies in flight during the gathers - measured no faster (T = 288: 40.0 vs 39.1 ms per
    // step, T = 24: 56.0 vs 52.9): two CTAs per SM already overlap copy and gather.  Larger graphs stay on the k2 kernels.
    g->db3 = 0;
    if (const char* e = std::getenv("MGA_S3_DB")) g->db3 = std::atoi(e) != 0;
    const int nb = g->db3 ? 2 : 1;
    const int cands[3] = {force > 0 ? std::min(force, g->C4) : std::min(g->C4, 8), std::min(g->C4, 4), 0};
    const size_t table = (size_t)g->N * std::max(g->kd, g->ku) * 8 + (size_t)(g->N + 2) * 8;
    // two CTAs per SM if a tile allows it, else one CTA of 1024 threads with the 8-chunk tile (MGA_S3_ONE=0 turns the
    // second pass off).  Measured: N = 600, T = 96 (one CTA, 8 chunks) 52.6 vs 60.2 ms per step on the k2 kernels;
    // N = 883, T = 288 (one CTA, 4 chunks) 50.2 vs 48.7 - so 4-chunk tiles are only used two per SM.
    bool one = !g->db3;
    if (const char* e = std::getenv("MGA_S3_ONE")) one = one && std::atoi(e) != 0;
    const size_t limits[2] = {g->db3 ? (size_t)226 * 1024 : (size_t)(228 * 1024) / 2 - 1024, (size_t)226 * 1024};
    g->one3 = 0;
    // MGA_S3_SINGLE: 0 = node tiles only, 1 (default) = one tile of all nodes only with 8-chunk (or whole-row) tiles at two
    // CTAs per SM, 2 = also 4-chunk tiles and the one-CTA-per-SM mode.  Measured against node tiles x 8-chunk time tiles:
    // N = 450, T = 96: 4-chunk single tile 42.5 ms per step, two node tiles of 225: 40.0; N = 600: one CTA per SM 53.0,
    // three node tiles of 200: 52.7 - so everything beyond N ~ 360 goes to node tiles.
    int single = 1;
    if (const char* e = std::getenv("MGA_S3_SINGLE")) single = std::atoi(e);
    for (int l = 0; l < (one && single == 2 ? 2 : 1) && single > 0 && g->CB3 == 0; ++l)
      for (int k = 0; cands[k] > 0 && k <= 1 - l && k < single && g->CB3 == 0; ++k)
        if ((size_t)g->N * (2 * nb * (size_t)cands[k] * 16 + (nb + 2) * 4) + 8 + table <= limits[l]) { g->CB3 = cands[k]; g->one3 = l; }
  }
  g->NT3 = 0;          // 0: one tile holds all nodes
  if (force >= 0 && g->CB3 == 0) {
    // node tiles for graphs beyond one CTA's shared memory: NT3 RCM-consecutive nodes per tile plus the external rows
    // they reference, over the whole row or 8-chunk time tiles; the plan checks that two CTAs fit an SM once the
    // external rows are known
    int nt = 256;
    if (const char* e = std::getenv("MGA_S3_NT")) nt = std::atoi(e);
    if (nt > 0) { g->CB3 = std::min(g->C4, force > 0 ? force : 8); g->NT3 = nt; g->db3 = 0; g->one3 = 0; }
  }
  stream2_threads3(g);
}

// block shape of the time-tiled kernels for the chosen tile (CB3, one3)
void stream2_threads3(Graph2* g) {
  if (g->CB3 <= 0) return;
  const int cb = g->CB3;
  // measured on B200 (PEMS04 graph): 512 threads (2 CTAs/SM) win for tiles of >= 4 chunks (T = 24: 52.3 vs 54.6 ms per
  // step, T = 288: 37.9 vs 38.4), 256 threads for the 3-chunk tile of T = 12 (55.3 vs 60.3)
  int threads = g->one3 ? 1024 : cb >= 4 ? 512 : 256;
  if (const char* e = std::getenv("MGA_S3_THREADS")) threads = std::atoi(e) >= 1024 ? 1024 : std::atoi(e) >= 512 ? 512 : 256;
  g->NB3t = (cb & (cb - 1)) == 0 ? threads / cb : 32 * std::max(1, (threads / 32) / cb);     // cb * NB3t whole warps
  g->tiles3 = (g->C4 + cb - 1) / cb;
}

bool stream2_eligible(const mga_plan* p, int dtype) {
  return p->has_s2 && dtype == MGA_F32;
}

struct Bufs2 {
  float *r, *p, *ap, *qs, *p2;     // p2: second p buffer of the time-tiled kernels (p is ping-ponged)
  double* dots;
};
constexpr int kCgVecs = 5;

static size_t vec_bytes2(const Graph2& g, int64_t B) { return (((size_t)B * g.N * g.C4 * 16) + 255) & ~(size_t)255; }
static size_t dots_bytes2(int64_t B, int max_iter) { return (((size_t)(2 * max_iter + 1) * B * sizeof(double)) + 255) & ~(size_t)255; }

// Resident CTAs per SM of a time-tiled kernel at this block size / dynamic shared memory (opt-in above 48 KB);
// asked once per (kernel, configuration, device).
static int k3_ctas_per_sm(const void* kern, int threads, size_t smem, int* rc) {
  static std::mutex mu;
  static std::map<std::tuple<const void*, int, size_t, int>, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(kern, threads, smem, dev);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  // the opt-in limit only ever grows: lowering it would break the cached larger configurations
  static std::map<std::pair<const void*, int>, size_t> limit;
  // (default limit: 48 KB minus the kernel's static shared memory - start below it)
  size_t& lim = limit.emplace(std::make_pair(kern, dev), (size_t)40 * 1024).first->second;
  cudaError_t e = cudaSuccess;
  if (smem > lim) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) lim = smem;
  }
  int n = 0;
  if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
  if (e != cudaSuccess || n < 1) { *rc = cuda_fail(e == cudaSuccess ? cudaErrorLaunchOutOfResources : e, "time-tiled kernel occupancy"); return 0; }
  cache[key] = n;
  return n;
}

// ---- fused TMA-staged kernels (mga_stream4.cuh): eligibility, tensor maps, launch -------------------------------
struct K4Plan {
  bool ok;
  int nstage, rows_box, nbox, rows_tile;
  int cons;              // consumer threads of the instantiation (512 / 768)
  size_t smem[2];        // by SYS
};

static size_t k4_smem_bytes(const Graph2& g, int nstage, int rows_tile, int sys) {
  const size_t N = g.N, kf = sys == 0 ? g.kd3 : g.ku3, n_in = sys == 0 ? g.in_ptr3_total : 0;
  return (size_t)nstage * 2 * rows_tile * kCB4 * 16 + 4 * (size_t)rows_tile * 16 + 5 * N * 4 + (2 * N + 2) * 4 + N * kf * 8 +
         n_in * 8 + (2 * (size_t)nstage + 2) * 8 + 2 * (size_t)nstage * 4 + 16;
}

static K4Plan k4_plan(const mga_plan* p) {
  const Graph2& g = p->g2;
  K4Plan k{};
  int want = 1;
  if (const char* e = std::getenv("MGA_S4")) want = std::atoi(e);
  k4_env(&k.cons);
  // one all-node tile of 8-chunk rows (the PEMS-sized graphs at T >= 29); node-tiled plans keep the k3 kernels
  if (!want || g.CB3 != kCB4 || g.ntile3 != 1 || g.db3 || g.C4 < kCB4 || g.N > k.cons) return k;
  k.nbox = (g.N + 255) / 256;
  k.rows_box = ((g.N + k.nbox - 1) / k.nbox + 7) & ~7;     // whole 128-byte lines in the halo buffer (TMA destinations)
  if (k.rows_box > 256) return k;
  k.rows_tile = k.rows_box * k.nbox;
  k.nstage = 2;
  for (int sys = 0; sys < 2; ++sys) {
    k.smem[sys] = k4_smem_bytes(g, k.nstage, k.rows_tile, sys);
    if (k.smem[sys] > (size_t)p->max_smem_optin) return k;
  }
  k.ok = true;
  return k;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// (4 C4 floats, N rows, B windows) view of a node-major workspace vector; box = {box_w floats, rows_box, 1}
static int k4_map(mga_plan* p, const float* v, int64_t B, int box_w, int rows_box, CUtensorMap* out) {
  static EncodeTiledFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
    return reinterpret_cast<EncodeTiledFn>(fn);
  }();
  if (!encode) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return MGA_ERR_CUDA; }
  const auto key = std::make_pair(static_cast<const void*>(v), B * 64 + box_w);
  auto it = p->tmaps.find(key);
  if (it == p->tmaps.end()) {
    if (p->tmaps.size() > 64) p->tmaps.clear();
    const Graph2& g = p->g2;
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)g.C4 * 4, (cuuint64_t)g.N, (cuuint64_t)B};
    const cuuint64_t strides[2] = {(cuuint64_t)g.C4 * 16, (cuuint64_t)g.N * g.C4 * 16};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)rows_box, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(v), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"); return MGA_ERR_CUDA; }
    it = p->tmaps.emplace(key, m).first;
  }
  *out = it->second;
  return MGA_OK;
}

template <int SYS, int SRC, int K, int NC>
static int k4_launch(mga_plan* p, const K4Plan& k4, const K4Maps& maps, const K4Args& a, cudaStream_t st) {
  auto kern = k4_cg<SYS, SRC, K, NC>;
  static std::mutex mu;
  static std::map<int, size_t> limit;            // per device: the opt-in limit only ever grows
  {
    std::lock_guard<std::mutex> lock(mu);
    size_t& lim = limit[p->device];
    if (k4.smem[SYS] > lim) {
      MGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4.smem[SYS]));
      lim = k4.smem[SYS];
    }
  }
  kern<<<std::min(a.total, p->sm_count), NC + 32, k4.smem[SYS], st>>>(maps, p->g2, a);
  MGA_LAUNCH_CHECK("k4_cg");
  return MGA_OK;
}

template <int SYS, int SRC, int K>
static int k4_launch_nc(mga_plan* p, const K4Plan& k4, const K4Maps& maps, const K4Args& a, cudaStream_t st) {
  if (k4.cons == 768) return k4_launch<SYS, SRC, K, 768>(p, k4, maps, a, st);
  return k4_launch<SYS, SRC, K, 512>(p, k4, maps, a, st);
}

template <int SYS, int SRC>
static int k4_launch_k(mga_plan* p, const K4Plan& k4, const K4Maps& maps, const K4Args& a, cudaStream_t st) {
  const int K = SYS == 0 ? p->g2.kd3 : p->g2.ku3;
  if (K == 4) return k4_launch_nc<SYS, SRC, 4>(p, k4, maps, a, st);
  if (K == 6) return k4_launch_nc<SYS, SRC, 6>(p, k4, maps, a, st);
  if (K == 8) return k4_launch_nc<SYS, SRC, 8>(p, k4, maps, a, st);
  return k4_launch_nc<SYS, SRC, 0>(p, k4, maps, a, st);
}

// one CG solve with the fused kernels: per iteration k4_cg (r, p -> p', Ap, <p', Ap>) + k2_xr
static int cg4(mga_plan* p, const K4Plan& k4, int system, const float* rhs, const float* x0, float* x, int64_t B, int n_cg,
               float a, float c, const Bufs2& w, cudaStream_t st) {
  const Graph2& g = p->g2;
  const dim3 fgrid((unsigned)B, (g.N * g.C4 + kFlat - 1) / kFlat);
  K4Args ka{};
  ka.B = B; ka.nstage = k4.nstage; ka.rows_box = k4.rows_box; ka.nbox = k4.nbox; ka.rows_tile = k4.rows_tile;
  ka.tiles = (g.C4 + kCB4 - 1) / kCB4;
  ka.total = (int)B * ka.tiles;
  ka.a = a; ka.cc = c; ka.xsys = system == MGA_SYS_X;
  ka.dots = w.dots;
  int rc;
  // tensor maps (cached per vector): tile boxes {32 floats, rows_box, 1}, halo boxes {4 floats, rows_box, 1}
  K4Maps m_init{}, m_it[2]{};
  const float* pbuf_c[2] = {w.p, w.p2};
  if ((rc = k4_map(p, x0, B, kCB4 * 4, k4.rows_box, &m_init.r)) || (rc = k4_map(p, x0, B, 4, k4.rows_box, &m_init.r_halo))) return rc;
  m_init.p = m_init.pnew = m_init.r;
  m_init.p_halo = m_init.r_halo;
  for (int cur = 0; cur < 2; ++cur) {            // p_old = pbuf[cur], p_new = pbuf[cur ^ 1]
    if ((rc = k4_map(p, w.r, B, kCB4 * 4, k4.rows_box, &m_it[cur].r)) || (rc = k4_map(p, w.r, B, 4, k4.rows_box, &m_it[cur].r_halo)) ||
        (rc = k4_map(p, pbuf_c[cur], B, kCB4 * 4, k4.rows_box, &m_it[cur].p)) ||
        (rc = k4_map(p, pbuf_c[cur], B, 4, k4.rows_box, &m_it[cur].p_halo)) ||
        (rc = k4_map(p, pbuf_c[cur ^ 1], B, kCB4 * 4, k4.rows_box, &m_it[cur].pnew)))
      return rc;
  }
  // r = rhs - A x0, RR(0)
  ka.it = 0; ka.rhs = rhs; ka.out = w.r; ka.slot = w.dots;
  rc = system == MGA_SYS_ZU ? k4_launch_k<1, 2>(p, k4, m_init, ka, st) : k4_launch_k<0, 2>(p, k4, m_init, ka, st);
  if (rc) return rc;
  int cur = 0;                                   // p_old = w.p (never read in the first iteration), p_new = w.p2
  float* pbuf[2] = {w.p, w.p2};
  // measured at T = 288, B = 256 (profiles/r02_defer_x.txt, same box): x / z_d iteration 214.4 -> 194.2 us, z_u 151.6 -> 147.4 us,
  // step 34.04 -> 32.04 ms.  MGA_S4_DEFER_X = 0 off, 1 two-hop systems only, 2 (default) all systems
  static const int defer_mode = [] { const char* e = std::getenv("MGA_S4_DEFER_X"); return e ? std::atoi(e) : 2; }();
  const bool defer_x = defer_mode == 2 || (defer_mode == 1 && system != MGA_SYS_ZU);
  for (int it = 0; it < n_cg; ++it) {
    ka.it = it; ka.rhs = nullptr; ka.out = w.ap;
    ka.slot = w.dots + (size_t)(2 * it + 1) * B;
    // x += alpha p of iteration it - 1 rides on this launch (it reads that p as its p_old); the last iteration's is k2_xr's
    ka.xd_in = (defer_x && it > 0) ? (it == 1 ? x0 : x) : nullptr;
    ka.xd_out = (defer_x && it > 0) ? x : nullptr;
    if (system == MGA_SYS_ZU)
      rc = it == 0 ? k4_launch_k<1, 1>(p, k4, m_it[cur], ka, st) : k4_launch_k<1, 0>(p, k4, m_it[cur], ka, st);
    else
      rc = it == 0 ? k4_launch_k<0, 1>(p, k4, m_it[cur], ka, st) : k4_launch_k<0, 0>(p, k4, m_it[cur], ka, st);
    if (rc) return rc;
    if (defer_x && it + 1 < n_cg) k2_xr<false><<<fgrid, kFlat, 0, st>>>(g, B, it, nullptr, nullptr, w.r, nullptr, w.ap, w.dots);
    else k2_xr<true><<<fgrid, kFlat, 0, st>>>(g, B, it, (it == 0 || (defer_x && n_cg == 1)) ? x0 : x, x, w.r, pbuf[cur ^ 1], w.ap, w.dots);
    MGA_LAUNCH_CHECK("k2_xr");
    cur ^= 1;
  }
  return MGA_OK;
}

// the tail of an outer iteration with the x tile staged by TMA (k5_tail); same arguments as k2_tail
template <int K>
static int k5_launch(mga_plan* p, const K4Plan& k4, const K5Maps& maps, K5Args a, cudaStream_t st) {
  constexpr int NC = 512;
  const Graph2& g = p->g2;
  a.nstage = 3;
  const size_t smem = (size_t)a.nstage * ((size_t)k4.rows_tile * kCB4 * 16 + (size_t)k4.rows_tile * 16) + ((size_t)g.N + 1) * 4 +
                      (size_t)g.N * (g.kd3 + g.ku3) * 8 + 2 * (size_t)a.nstage * 8 + (size_t)MGA_DIAG_COLS * 32 * 4 + 16;
  if (smem > (size_t)p->max_smem_optin) { a.nstage = 2; }
  const size_t smem2 = a.nstage == 3 ? smem : smem - ((size_t)k4.rows_tile * kCB4 * 16 + (size_t)k4.rows_tile * 16) - 16;
  auto kern = k5_tail<K, NC>;
  static std::mutex mu;
  static std::map<int, size_t> limit;
  {
    std::lock_guard<std::mutex> lock(mu);
    size_t& lim = limit[p->device];
    if (smem2 > lim) {
      MGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      lim = smem2;
    }
  }
  kern<<<std::min(a.total, p->sm_count), NC + 32, smem2, st>>>(maps, g, a);
  MGA_LAUNCH_CHECK("k5_tail");
  return MGA_OK;
}

static bool k5_ok(const mga_plan* p, const K4Plan& k4) {
  static const bool on = [] { const char* e = std::getenv("MGA_S5"); return !e || std::atoi(e) != 0; }();
  return on && k4.ok && p->g2.N * kCB4 <= kK5Slots * 512;
}

static int k5_tail_launch(mga_plan* p, const K4Plan& k4, int want_diag, const float* x, const float* x_old, const float* zu,
                          const float* zu_old, const float* zd, const float* zd_old, float* gu, float* gd, float* gam, float* phi,
                          const float* y, float rho, float rho_u, float rho_d, float thr, double* diag, double* dx_sum, int64_t B,
                          cudaStream_t st) {
  const Graph2& g = p->g2;
  K5Maps maps{};
  int rc;
  if ((rc = k4_map(p, x, B, kCB4 * 4, k4.rows_box, &maps.x)) || (rc = k4_map(p, x, B, 4, k4.rows_box, &maps.x_halo))) return rc;
  K5Args a{};
  a.B = B; a.rows_box = k4.rows_box; a.nbox = k4.nbox; a.rows_tile = k4.rows_tile;
  a.tiles = (g.C4 + kCB4 - 1) / kCB4;
  a.total = (int)B * a.tiles;
  a.want_diag = want_diag;
  a.x = x; a.x_old = x_old; a.zu = zu; a.zu_old = zu_old; a.zd = zd; a.zd_old = zd_old;
  a.gu = gu; a.gd = gd; a.gam = gam; a.phi = phi; a.y = y;
  a.rho = rho; a.rho_u = rho_u; a.rho_d = rho_d; a.thr = thr;
  a.diag = diag; a.dx_sum = dx_sum;
  const int K = g.kd3 == g.ku3 ? g.kd3 : 0;
  if (K == 4) return k5_launch<4>(p, k4, maps, a, st);
  if (K == 6) return k5_launch<6>(p, k4, maps, a, st);
  if (K == 8) return k5_launch<8>(p, k4, maps, a, st);
  return k5_launch<0>(p, k4, maps, a, st);
}

// CG_solver (ADMM.py:329-368) with a fixed iteration count on internal-layout vectors; x holds x0 / the solution.
// x0: warm start (read only); x: the solution (x0 == x: in place).  With n_cg == 0 the solution IS the warm start.
static int cg2(mga_plan* p, int system, const mga_params* m, const float* rhs, const float* x0, float* x, int64_t B,
               int64_t B_out, int n_cg, float* alpha, float* beta, const Bufs2& w, cudaStream_t st) {
  const Graph2& g = p->g2;
  const dim3 grid((unsigned)B, g.tilesN, g.tilesC), blk(g.CB, g.NBt);
  const dim3 fgrid((unsigned)B, (g.N * g.C4 + kFlat - 1) / kFlat);
  float a, c;
  if (system == MGA_SYS_X) { a = (float)((m->rho_u + m->rho_d) / 2); c = (float)(m->rho / 2); }
  else if (system == MGA_SYS_ZU) { a = (float)(m->rho_u / 2); c = (float)m->mu_u; }
  else { a = (float)(m->rho_d / 2); c = (float)m->mu_d2; }
  const int xsys = system == MGA_SYS_X;
  MGA_CUDA(cudaMemsetAsync(w.dots, 0, (size_t)(2 * n_cg + 1) * B * sizeof(double), st));
  const K4Plan k4 = k4_plan(p);
  if (k4.ok) {
    const int rc4 = cg4(p, k4, system, rhs, x0, x, B, n_cg, a, c, w, st);
    if (rc4) return rc4;
  } else if (g.CB3 > 0) {
    // time-tiled shared-memory kernels: p update fused into the operator kernel, p ping-ponged
    const dim3 blk3(g.CB3, g.NB3t);
    const int total = (int)B * g.tiles3 * g.ntile3;
    float *p_old = w.p, *p_new = w.p2;
    int rc3 = MGA_OK;
#define MGA_K3_LAUNCH(kern, smem, ...)                                                    \
    do {                                                                                   \
      const int ctas = k3_ctas_per_sm((const void*)kern, g.CB3 * g.NB3t, (size_t)(smem), &rc3);      \
      if (rc3) return rc3;                                                                 \
      kern<<<std::min(total, ctas * p->sm_count), blk3, (size_t)(smem), st>>>(__VA_ARGS__); \
      MGA_LAUNCH_CHECK(#kern);                                                             \
    } while (0)
#define MGA_K3_BY_K(K, macro) \
    do { if ((K) == 4) { macro(4); } else if ((K) == 6) { macro(6); } else if ((K) == 8) { macro(8); } else { macro(0); } } while (0)
    if (system == MGA_SYS_ZU) {
#define MGA_K3_LU_INIT(K) MGA_K3_LAUNCH((k3_lu<2, 1, K>), g.smem3_u, g, B, 0, x0, nullptr, nullptr, rhs, w.r, w.dots, w.dots, a, c, nullptr, nullptr)
      MGA_K3_BY_K(g.ku3, MGA_K3_LU_INIT);
    } else {
#define MGA_K3_PLDR_INIT(K) MGA_K3_LAUNCH((k3_p_ldr<2, K>), g.smem3_d, g, B, 0, x0, nullptr, nullptr, w.qs, w.dots, nullptr, nullptr)
      MGA_K3_BY_K(g.kd3, MGA_K3_PLDR_INIT);
      MGA_K3_LAUNCH((k3_ldrt_lhs<1>), g.smem3_in, g, B, x0, w.qs, rhs, w.r, w.dots, a, c, xsys);
    }
    // deferred x update as in cg4: iteration it - 1's x += alpha p rides on this iteration's tile kernel (its p_old), k2_xr<false>
    // updates r only; own rows per thread must fit the kernel's register array.  MGA_S3_DEFER_X=0: off
    static const bool defer_env = [] { const char* e = std::getenv("MGA_S3_DEFER_X"); return !e || std::atoi(e) != 0; }();
    const int own_rows = g.ntile3 > 1 ? g.NT3 : g.N;
    const bool defer_x = defer_env && !g.db3 && (own_rows + g.NB3t - 1) / g.NB3t <= kXR3;
    for (int it = 0; it < n_cg; ++it) {
      double* pap = w.dots + (size_t)(2 * it + 1) * B;
      const float* xd_in = (defer_x && it > 0) ? (it == 1 ? x0 : x) : nullptr;
      float* xd_out = (defer_x && it > 0) ? x : nullptr;
      if (system == MGA_SYS_ZU) {
#define MGA_K3_LU_FIRST(K) MGA_K3_LAUNCH((k3_lu<1, 0, K>), g.smem3_u, g, B, it, w.r, p_old, p_new, nullptr, w.ap, w.dots, pap, a, c, nullptr, nullptr)
#define MGA_K3_LU_NEXT(K) MGA_K3_LAUNCH((k3_lu<0, 0, K>), g.smem3_u, g, B, it, w.r, p_old, p_new, nullptr, w.ap, w.dots, pap, a, c, xd_in, xd_out)
        if (it == 0) MGA_K3_BY_K(g.ku3, MGA_K3_LU_FIRST);
        else MGA_K3_BY_K(g.ku3, MGA_K3_LU_NEXT);
      } else {
#define MGA_K3_PLDR_FIRST(K) MGA_K3_LAUNCH((k3_p_ldr<1, K>), g.smem3_d, g, B, it, w.r, p_old, p_new, w.qs, w.dots, nullptr, nullptr)
#define MGA_K3_PLDR_NEXT(K) MGA_K3_LAUNCH((k3_p_ldr<0, K>), g.smem3_d, g, B, it, w.r, p_old, p_new, w.qs, w.dots, xd_in, xd_out)
        if (it == 0) MGA_K3_BY_K(g.kd3, MGA_K3_PLDR_FIRST);
        else MGA_K3_BY_K(g.kd3, MGA_K3_PLDR_NEXT);
        MGA_K3_LAUNCH((k3_ldrt_lhs<0>), g.smem3_in, g, B, p_new, w.qs, nullptr, w.ap, pap, a, c, xsys);
      }
      if (defer_x && it + 1 < n_cg) k2_xr<false><<<fgrid, kFlat, 0, st>>>(g, B, it, nullptr, nullptr, w.r, nullptr, w.ap, w.dots);
      else k2_xr<true><<<fgrid, kFlat, 0, st>>>(g, B, it, it == 0 ? x0 : x, x, w.r, p_new, w.ap, w.dots);
      MGA_LAUNCH_CHECK("k2_xr");
      std::swap(p_old, p_new);
    }
  } else {
  // r = rhs - A x0 ; RR(0)
  if (system == MGA_SYS_ZU) {
    k2_lu_lhs<1><<<grid, blk, 0, st>>>(g, x0, rhs, w.r, w.dots, a, c);
    MGA_LAUNCH_CHECK("k2_lu_lhs");
  } else {
    k2_ldr_shift<<<grid, blk, 0, st>>>(g, x0, w.qs);
    MGA_LAUNCH_CHECK("k2_ldr_shift");
    k2_ldrt_lhs<1><<<grid, blk, 0, st>>>(g, x0, w.qs, rhs, w.r, w.dots, a, c, xsys);
    MGA_LAUNCH_CHECK("k2_ldrt_lhs");
  }
  for (int it = 0; it < n_cg; ++it) {
    k2_pupdate<<<fgrid, kFlat, 0, st>>>(g, B, it, w.r, w.p, w.dots);
    MGA_LAUNCH_CHECK("k2_pupdate");
    double* pap = w.dots + (size_t)(2 * it + 1) * B;
    if (system == MGA_SYS_ZU) {
      k2_lu_lhs<0><<<grid, blk, 0, st>>>(g, w.p, nullptr, w.ap, pap, a, c);
      MGA_LAUNCH_CHECK("k2_lu_lhs");
    } else {
      k2_ldr_shift<<<grid, blk, 0, st>>>(g, w.p, w.qs);
      MGA_LAUNCH_CHECK("k2_ldr_shift");
      k2_ldrt_lhs<0><<<grid, blk, 0, st>>>(g, w.p, w.qs, nullptr, w.ap, pap, a, c, xsys);
      MGA_LAUNCH_CHECK("k2_ldrt_lhs");
    }
    k2_xr<true><<<fgrid, kFlat, 0, st>>>(g, B, it, it == 0 ? x0 : x, x, w.r, w.p, w.ap, w.dots);
    MGA_LAUNCH_CHECK("k2_xr");
  }
  }
  if (n_cg == 0 && x != x0)
    MGA_CUDA(cudaMemcpyAsync(x, x0, (size_t)B * g.N * g.C4 * 16, cudaMemcpyDeviceToDevice, st));
  if ((alpha || beta) && n_cg > 0) {
    const int64_t tot = B * n_cg;
    k2_coeffs<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(B, B_out, n_cg, w.dots, alpha, beta);
    MGA_LAUNCH_CHECK("k2_coeffs");
  }
  return MGA_OK;
}

static Bufs2 carve2(char* base, const Graph2& g, int64_t B, int max_iter) {
  const size_t vec = vec_bytes2(g, B);
  Bufs2 w;
  w.r = reinterpret_cast<float*>(base);
  w.p = reinterpret_cast<float*>(base + vec);
  w.ap = reinterpret_cast<float*>(base + 2 * vec);
  w.qs = reinterpret_cast<float*>(base + 3 * vec);
  w.p2 = reinterpret_cast<float*>(base + 4 * vec);
  w.dots = reinterpret_cast<double*>(base + kCgVecs * vec);
  (void)max_iter;
  return w;
}

// Windows per group.  Windows are independent, so a batch can be walked in groups; measured on B200, keeping
// a group's CG vectors L2-resident buys nothing (the gather kernels are bound by L1 wavefronts, not by where
// the rows come from: T=288 5.57k windows/s with groups of 64 or with all 256 at once), while small groups
// become launch-bound.  Groups therefore only cap the workspace (16 vectors per window) at 32 GB.
static int64_t group_windows(const mga_plan* p, int64_t B) {
  if (const char* e = std::getenv("MGA_S2_GROUP")) { const long v = std::atol(e); if (v > 0) return std::min<int64_t>(B, v); }
  const Graph2& g = p->g2;
  const double per_window = (12.0 + kCgVecs) * (double)g.N * g.C4 * 16.0;
  const int64_t G = (int64_t)(32e9 / per_window);
  return std::max<int64_t>(1, std::min<int64_t>(B, G));
}

static int stream2_cg_group(mga_plan* p, int system, const mga_params* m, const void* rhs, void* x, int64_t B, int64_t B_out,
                            int n_cg, void* alpha, void* beta, cudaStream_t st);

// mga_cg_solve, fixed iteration count, caller-layout rhs / x
int stream2_cg(mga_plan* p, int system, const mga_params* m, const void* rhs, void* x, int64_t B, int n_cg, void* alpha,
               void* beta, cudaStream_t st) {
  const int64_t G = group_windows(p, B);
  const size_t win = (size_t)p->g2.T * p->g2.N;
  for (int64_t b0 = 0; b0 < B; b0 += G) {
    const int64_t nb = std::min(G, B - b0);
    int rc = stream2_cg_group(p, system, m, static_cast<const float*>(rhs) + b0 * win, static_cast<float*>(x) + b0 * win, nb, B,
                              n_cg, alpha ? static_cast<float*>(alpha) + b0 : nullptr,
                              beta ? static_cast<float*>(beta) + b0 : nullptr, st);
    if (rc) return rc;
  }
  return MGA_OK;
}

static int stream2_cg_group(mga_plan* p, int system, const mga_params* m, const void* rhs, void* x, int64_t B, int64_t B_out,
                            int n_cg, void* alpha, void* beta, cudaStream_t st) {
  const Graph2& g = p->g2;
  const size_t vec = vec_bytes2(g, B);
  int rc = ensure_workspace(p, p->ws, (2 + kCgVecs) * vec + dots_bytes2(B, n_cg) + 256);
  if (rc) return rc;
  char* base = static_cast<char*>(p->ws.base);
  float* rhs_i = reinterpret_cast<float*>(base);
  float* x_i = reinterpret_cast<float*>(base + vec);
  Bufs2 w = carve2(base + 2 * vec, g, B, n_cg);
  const dim3 grid((unsigned)B, g.tilesN, g.tilesC), blk(g.CB, g.NBt);
  k2_import<<<grid, blk, 0, st>>>(g, static_cast<const float*>(rhs), rhs_i);
  MGA_LAUNCH_CHECK("k2_import");
  k2_import<<<grid, blk, 0, st>>>(g, static_cast<const float*>(x), x_i);
  MGA_LAUNCH_CHECK("k2_import");
  rc = cg2(p, system, m, rhs_i, x_i, x_i, B, B_out, n_cg, static_cast<float*>(alpha), static_cast<float*>(beta), w, st);
  if (rc) return rc;
  k2_export<<<grid, blk, 0, st>>>(g, x_i, static_cast<float*>(x));
  MGA_LAUNCH_CHECK("k2_export");
  return MGA_OK;
}

static int stream2_admm_group(mga_plan* p, const mga_params* prm, const void* y_, void* x_out, int64_t B, int64_t B_out,
                              int n_outer, int max_cg, double t_mean, double t_var, int diag_flags,
                              const mga_admm_outputs* outs, cudaStream_t st);

// combined_loop (ADMM.py:528-648), forecasting mode, ablation None, fixed iteration counts.  Windows are
// independent, so the batch is walked in L2-sized groups (group_windows); diagnostics accumulate over groups.
int stream2_admm(mga_plan* p, const mga_params* prm, const void* y_, void* x_out, int64_t B, int n_outer, int max_cg,
                 double t_mean, double t_var, int diag_flags, const mga_admm_outputs* outs, cudaStream_t st) {
  const Graph2& g = p->g2;
  const int64_t G = group_windows(p, B);
  const size_t win = (size_t)g.T * g.N, ywin = (size_t)g.t_in * g.N;
  for (int64_t b0 = 0; b0 < B; b0 += G) {
    const int64_t nb = std::min(G, B - b0);
    mga_admm_outputs o = *outs;
    auto off = [&](void* ptr, size_t n) -> void* { return ptr ? static_cast<float*>(ptr) + n : nullptr; };
    o.zu = off(outs->zu, b0 * win); o.zd = off(outs->zd, b0 * win); o.phi = off(outs->phi, b0 * win);
    o.gamma = off(outs->gamma, b0 * win); o.gamma_u = off(outs->gamma_u, b0 * win); o.gamma_d = off(outs->gamma_d, b0 * win);
    o.alpha = off(outs->alpha, b0); o.beta = off(outs->beta, b0);
    const int flags = b0 == 0 ? diag_flags : (diag_flags | 2);       // later groups add to the first one's sums
    int rc = stream2_admm_group(p, prm, static_cast<const float*>(y_) + b0 * ywin, static_cast<float*>(x_out) + b0 * win, nb, B,
                                n_outer, max_cg, t_mean, t_var, flags, &o, st);
    if (rc) return rc;
  }
  return MGA_OK;
}

static int stream2_admm_group(mga_plan* p, const mga_params* prm, const void* y_, void* x_out, int64_t B, int64_t B_out,
                              int n_outer, int max_cg, double t_mean, double t_var, int diag_flags,
                              const mga_admm_outputs* outs, cudaStream_t st) {
  const Graph2& g = p->g2;
  const float* y = static_cast<const float*>(y_);
  const bool want_diag = (diag_flags & 1) != 0, accumulate = (diag_flags & 2) != 0;
  const size_t vec = vec_bytes2(g, B);
  const int n_state = 12;   // x0, x1, zu0, zu1, zd0, zd1, gu, gd, gam, phi, rhs, spare
  const size_t need = n_state * vec + kCgVecs * vec + dots_bytes2(B, max_cg) + 512 +
                      (size_t)std::max(n_outer, 1) * MGA_DIAG_COLS * sizeof(double);
  int rc = ensure_workspace(p, p->ws, need);
  if (rc) return rc;
  char* base = static_cast<char*>(p->ws.base);
  auto V = [&](int k) { return reinterpret_cast<float*>(base + (size_t)k * vec); };
  float *x_cur = V(0), *x_nxt = V(1), *zu_cur = V(2), *zu_nxt = V(3), *zd_cur = V(4), *zd_nxt = V(5);
  float *gu = V(6), *gd = V(7), *gam = V(8), *phi = V(9), *rhs = V(10);
  Bufs2 w = carve2(base + n_state * vec, g, B, max_cg);
  double* own_diag = reinterpret_cast<double*>(base + n_state * vec + kCgVecs * vec + dots_bytes2(B, max_cg) + 256);
  double* diag = outs->diag ? outs->diag : own_diag;
  double* dx_sum = want_diag ? outs->dx_sum : nullptr;
  if (want_diag && !accumulate) {
    MGA_CUDA(cudaMemsetAsync(diag, 0, (size_t)n_outer * MGA_DIAG_COLS * sizeof(double), st));
    if (dx_sum) MGA_CUDA(cudaMemsetAsync(dx_sum, 0, (size_t)n_outer * g.T * g.N * sizeof(double), st));
  }
  double* nf_row = own_diag;     // non-finite flag row when diagnostics are off
  if (!want_diag) MGA_CUDA(cudaMemsetAsync(nf_row, 0, MGA_DIAG_COLS * sizeof(double), st));

  const dim3 grid((unsigned)B, g.tilesN, g.tilesC), blk(g.CB, g.NBt);
  const size_t chunks = (size_t)B * g.N * g.C4;
  const unsigned grid_c = (unsigned)((chunks + kFlat - 1) / kFlat);
  const unsigned grid_bn = (unsigned)((B * g.N * 32 + 255) / 256);      // one warp per (window, node)
  k2_init<<<grid_bn, 256, 0, st>>>(g, B, y, x_cur, zu_cur, zd_cur, gu, gd, gam, (float)t_mean, (float)t_var);
  MGA_LAUNCH_CHECK("k2_init");
  k2_ldr<<<grid, blk, 0, st>>>(g, x_cur, phi);      // phi = L_d x (ADMM.py:541)
  MGA_LAUNCH_CHECK("k2_ldr");
  const float rho = (float)prm->rho, rho_u = (float)prm->rho_u, rho_d = (float)prm->rho_d;
  const float thr = (float)(prm->mu_d1 / prm->rho);
  const size_t coef_stride = (size_t)max_cg * B_out;
  const K4Plan k4t = k4_plan(p);
  for (int it = 0; it < n_outer; ++it) {
    auto coef = [&](void* basep, int s) -> float* {
      return basep ? static_cast<float*>(basep) + ((size_t)it * 3 + s) * coef_stride : nullptr;
    };
    k2_rhs_x<<<grid, blk, 0, st>>>(g, gam, phi, zu_cur, zd_cur, gu, gd, y, rhs, rho, rho_u, rho_d);
    MGA_LAUNCH_CHECK("k2_rhs_x");
    // warm starts (ADMM.py:571, 580, 588) are read in place: the first x update of a solve writes the other buffer
    if ((rc = cg2(p, MGA_SYS_X, prm, rhs, x_cur, x_nxt, B, B_out, max_cg, coef(outs->alpha, 0), coef(outs->beta, 0), w, st))) return rc;
    k2_rhs_z<<<grid_c, kFlat, 0, st>>>(chunks, gu, x_nxt, rhs, (float)(prm->rho_u / 2));
    MGA_LAUNCH_CHECK("k2_rhs_z");
    if ((rc = cg2(p, MGA_SYS_ZU, prm, rhs, zu_cur, zu_nxt, B, B_out, max_cg, coef(outs->alpha, 1), coef(outs->beta, 1), w, st))) return rc;
    k2_rhs_z<<<grid_c, kFlat, 0, st>>>(chunks, gd, x_nxt, rhs, (float)(prm->rho_d / 2));
    MGA_LAUNCH_CHECK("k2_rhs_z");
    if ((rc = cg2(p, MGA_SYS_ZD, prm, rhs, zd_cur, zd_nxt, B, B_out, max_cg, coef(outs->alpha, 2), coef(outs->beta, 2), w, st))) return rc;
    double* drow = want_diag ? diag + (size_t)it * MGA_DIAG_COLS : nf_row;
    if (k5_ok(p, k4t)) {
      if ((rc = k5_tail_launch(p, k4t, want_diag ? 1 : 0, x_nxt, x_cur, zu_nxt, zu_cur, zd_nxt, zd_cur, gu, gd, gam, phi, y, rho, rho_u,
                               rho_d, thr, drow, dx_sum ? dx_sum + (size_t)it * g.T * g.N : nullptr, B, st)))
        return rc;
    } else {
      k2_tail<<<grid, blk, 0, st>>>(g, want_diag ? 1 : 0, x_nxt, x_cur, zu_nxt, zu_cur, zd_nxt, zd_cur, gu, gd, gam, phi,
                                   y, rho, rho_u, rho_d, thr, drow, dx_sum ? dx_sum + (size_t)it * g.T * g.N : nullptr);
      MGA_LAUNCH_CHECK("k2_tail");
    }
    std::swap(x_cur, x_nxt);
    std::swap(zu_cur, zu_nxt);
    std::swap(zd_cur, zd_nxt);
    if (outs->cg_iters) { outs->cg_iters[it * 3 + 0] = -1; outs->cg_iters[it * 3 + 1] = -1; outs->cg_iters[it * 3 + 2] = -1; }
  }
  auto give = [&](void* dst, const float* src) -> int {
    if (!dst) return MGA_OK;
    k2_export<<<grid, blk, 0, st>>>(g, src, static_cast<float*>(dst));
    MGA_LAUNCH_CHECK("k2_export");
    return MGA_OK;
  };
  if ((rc = give(x_out, x_cur))) return rc;
  if ((rc = give(outs->zu, zu_cur))) return rc;
  if ((rc = give(outs->zd, zd_cur))) return rc;
  if ((rc = give(outs->gamma_u, gu))) return rc;
  if ((rc = give(outs->gamma_d, gd))) return rc;
  if ((rc = give(outs->phi, phi))) return rc;
  if ((rc = give(outs->gamma, gam))) return rc;
  if (outs->outer_done) *outs->outer_done = n_outer;
  return MGA_OK;
}

}  // namespace mga
