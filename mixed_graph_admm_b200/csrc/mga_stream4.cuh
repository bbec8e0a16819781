// Streaming mode, fused CG phase 1 with TMA-staged tiles ("k4"): included by mga_stream2.cu.
//
// One kernel per CG iteration does what k3_p_ldr + k3_ldrt_lhs (x / z_d systems) or k3_lu (z_u system) did:
//     (r, p) -> p' = r + beta p,  Ap = A p',  <p', Ap>                      (ADMM.py:348-358, operators 138-228)
// so the intermediate qs = shifted L_d p' never goes to HBM: 16 B per lattice point (+ halo) instead of 28, and with
// k2_xr (24 B/pt) a CG iteration of the 2-hop systems moves 40 B/pt in 2 launches (k3: 52 in 3; algorithmic: 48).
//
// A persistent CTA per SM walks (window, time tile) items.  A tile is ALL nodes x 8 chunks (32 time steps) of r and p:
//   * a producer warp (one elected lane) stages the next tile while the consumer warps work on the current one:
//     cp.async.bulk.tensor (TMA) 3-D boxes {32 floats, <= 256 node rows, 1 window} of a (T, N, B) tensor map, completion
//     on an mbarrier (expect-tx); out-of-range rows / columns are zero-filled by the TMA unit, so partial tiles and the
//     rounding of the node rows to whole boxes need no code.  Two stages (full / empty barrier pairs).  The same lane
//     computes beta of the tile's window (it can afford the latency of the two loads) and leaves it with the stage.
//   * L_d^T L_d couples t-1, t, t+1, so a tile needs p' one step beyond each edge, for every node.  Those halo columns
//     come through TMA as well: boxes {4 floats, node rows, 1} left and right of the tile into one small buffer with its
//     own barrier pair (free again as soon as phase A has consumed it).  Fetching them with per-thread loads - 4 scalars
//     per node and tile, whatever the mechanism: registers or 4-byte cp.async - cost 45 of 160 us per launch: 32 distinct
//     sectors per warp instruction sit in the same LSU queue as the shared-memory gathers.
//   * phase A: p' = r + beta p in place in the r tile; one thread then hands the tile to a TMA store (p' -> HBM),
//   * phase B: qs = shifted L_d p' (gathers from the p' tile, table entries = (row byte offset, weight)) overwrites the
//     dead p tile; q at the tile's first step is one scalar gather per node from the halo column,
//   * phase C: Ap = D p' + c (q - in-list gather of qs), streamed to HBM, <p', Ap> added per warp.
// Thread = one 16-byte chunk (4 time steps of one node); the 8 lanes of a quarter-warp own the 8 chunks of one row, so a
// neighbour's chunk gather is a conflict-free 128-byte shared-memory wavefront, and the element next to a chunk comes
// from the neighbouring lane by shuffle.
#pragma once
#include <cuda.h>

#ifndef MGA_K4_X            // timing experiments only (bit mask): 1 no phase-C gather, 2 no phase-B gather, 4 no p' TMA store,
#define MGA_K4_X 0          // 8 no Ap store, 16 no halo fetch, 32 no phase A, 64 no first-step q gather
#endif

namespace mga {

constexpr int kCB4 = 8;                    // chunks per tile row

struct K4Args {
  int64_t B;
  int it;
  int nstage, rows_box, nbox, rows_tile;   // rows_tile = rows_box * nbox >= N
  int tiles, total;                        // time tiles per window, B * tiles
  const float* rhs;      // SRC 2
  float* out;            // Ap (SRC 0 / 1) or r = rhs - A x0 (SRC 2)
  const double* dots;    // RR(k) = dots[2k], PAP(k) = dots[2k+1], each (B)
  double* slot;          // (B): where this launch's dot product goes
  float a, cc;
  int xsys;
  // deferred x update of the PREVIOUS iteration (SRC 0): x_out = x_in + alpha(it-1) p_old, with the p_old tile this launch
  // stages anyway - k2_xr then only updates r (12 instead of 24 B/pt).  NULL: off.
  const float* xd_in;
  float* xd_out;
};

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(b)) : "memory");
}
// bounded wait: a lost completion traps (error to the host) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t addr = s2u(b);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load3(void* dst, const CUtensorMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(s2u(dst)), "l"(m), "r"(s2u(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store3(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(s2u(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NC>
__device__ __forceinline__ void bar_cons() { asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory"); }

struct K4Maps {          // tile boxes {32 floats, rows_box, 1} and halo boxes {4 floats, rows_box, 1} of the same vectors
  CUtensorMap r, p, pnew, r_halo, p_halo;
};

// SYS 0: A = diag + c L_d^T L_d (x / z_d systems), SYS 1: A = c L_u + a I (z_u system)
// SRC 0: v = r + beta p (stored as the new p)   SRC 1: v = r (first iteration; stored as p)   SRC 2: v = x0, out = rhs - A v
// K: compile-time width of the forward table (0: run-time)
// NC: consumer threads (a multiple of 32 and of 8, >= N); the block has NC + 32 threads (one producer warp)
template <int SYS, int SRC, int K, int NC>
__global__ void __launch_bounds__(NC + 32, 1)
k4_cg(const __grid_constant__ K4Maps maps, const Graph2 g, const K4Args a) {
  extern __shared__ __align__(128) unsigned char smem4[];
  const int N = g.N, C4 = g.C4, T = g.T;
  const int tid = threadIdx.x;
  const size_t tile_f4 = (size_t)a.rows_tile * kCB4;                 // float4 per tile buffer
  float4* tiles = reinterpret_cast<float4*>(smem4);                  // [stage][R | P][rows_tile * 8]
  float4* hbuf = tiles + 2 * tile_f4 * a.nstage;                     // [r left, r right, p left, p right][rows_tile]: halo chunks
  float* hl = reinterpret_cast<float*>(hbuf + 4 * (size_t)a.rows_tile);   // p' one step before the tile, per node
  float* hr = hl + N;                                                // p' one step after the tile
  float* qsl = hr + N;                                               // q at the tile's first step
  float* p0 = qsl + N;                                               // p' at the tile's first step
  float* wself = p0 + N;
  int* ptr = reinterpret_cast<int*>(wself + N);                      // (N + 1) words + (N) words: the phase-C slots  [SYS 0]
  int* ord = ptr + N + 1;                                            //   int2 sdesc[k] = (row | list length << 16, first in-list entry), k-th row in ord4 order
  const int kf = SYS == 0 ? g.kd3 : g.ku3;
  int2* sdesc = reinterpret_cast<int2*>(ptr + ((5 * N) & 1));        // 8-byte aligned inside the (2 N + 2)-word block
  int2* tab = reinterpret_cast<int2*>(ord + N + ((N & 1) ^ 1));      // forward table (N, kf), 8-byte aligned: 7 N + 1 words precede it (odd iff N is even)
  int2* tab_in = tab + (size_t)N * kf;                               // in-list entries                  [SYS 0]
  const int n_in = SYS == 0 ? g.in_ptr3_total : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab_in + n_in);       // full[nstage], empty[nstage]
  uint64_t* full = bars;
  uint64_t* empty = bars + a.nstage;
  uint64_t* hfull = bars + 2 * a.nstage;                             // the halo buffer: one barrier pair
  uint64_t* hempty = hfull + 1;
  float* beta_s = reinterpret_cast<float*>(hempty + 1);              // (nstage) beta of the window of the staged tile, then (nstage) alpha of the previous iteration

  // ---- barrier init, then the producer starts fetching while the consumers stage the graph tables
  if (tid == 0) {
    for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NC / 32); }
    mbar_init(hfull, 1);
    mbar_init(hempty, NC / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  __syncthreads();
  const uint32_t stage_bytes = (uint32_t)(tile_f4 * 16) * (SRC == 0 ? 2u : 1u);
  const uint32_t halo_bytes = (uint32_t)(a.rows_tile * 16) * (SRC == 0 ? 4u : 2u);
  const bool want_halo = SYS == 0 && !(MGA_K4_X & 16);

  if (tid >= NC) {
    // ================= producer warp: one lane keeps the stages full =================
    if (tid == NC) {
      int k = 0;
      for (int tl = blockIdx.x; tl < a.total; tl += gridDim.x, ++k) {
        const int s = k % a.nstage, ph = (k / a.nstage) & 1;
        const int b = tl / a.tiles, c0 = (tl - b * a.tiles) * kCB4;
        mbar_wait(empty + s, ph ^ 1);                                // first pass over the stages: free at once
        // beta of the tile's window travels with the stage: this lane can afford the load latency, the consumers cannot
        // (written before the arrive on `full`, whose release / acquire pair publishes it)
        if (SRC == 0) {
          beta_s[s] = (float)a.dots[(size_t)(2 * a.it) * a.B + b] / (float)a.dots[(size_t)(2 * a.it - 2) * a.B + b];   // ADMM.py:356
          if (a.xd_out)      // alpha of the previous iteration, formed as k2_xr forms it (ADMM.py:351)
            beta_s[a.nstage + s] = (float)a.dots[(size_t)(2 * a.it - 2) * a.B + b] / (float)a.dots[(size_t)(2 * a.it - 1) * a.B + b];
        }
        mbar_expect_tx(full + s, stage_bytes);
        float4* tR = tiles + (size_t)s * 2 * tile_f4;
        float4* tP = tR + tile_f4;
        for (int j = 0; j < a.nbox; ++j) {
          tma_load3(tR + (size_t)j * a.rows_box * kCB4, &maps.r, 4 * c0, j * a.rows_box, b, full + s);
          if (SRC == 0) tma_load3(tP + (size_t)j * a.rows_box * kCB4, &maps.p, 4 * c0, j * a.rows_box, b, full + s);
        }
        if (want_halo) {
          // the halo buffer is free once phase A of the previous tile has read it; columns outside the window are zero-filled
          mbar_wait(hempty, (k & 1) ^ 1);
          mbar_expect_tx(hfull, halo_bytes);
          for (int j = 0; j < a.nbox; ++j) {
            const int r0 = j * a.rows_box;
            tma_load3(hbuf + r0, &maps.r_halo, 4 * c0 - 4, r0, b, hfull);
            tma_load3(hbuf + a.rows_tile + r0, &maps.r_halo, 4 * (c0 + kCB4), r0, b, hfull);
            if (SRC == 0) {
              tma_load3(hbuf + 2 * a.rows_tile + r0, &maps.p_halo, 4 * c0 - 4, r0, b, hfull);
              tma_load3(hbuf + 3 * a.rows_tile + r0, &maps.p_halo, 4 * (c0 + kCB4), r0, b, hfull);
            }
          }
        }
      }
    }
    return;
  }

  // ================= consumers =================
  {
    const int2* src = SYS == 0 ? g.tab_d : g.tab_u;
    for (int k = tid; k < N * kf; k += NC) tab[k] = src[k];
    if (SYS == 0) {
      for (int k = tid; k < n_in; k += NC) tab_in[k] = g.tab_in3[k];
      for (int k = tid; k < N; k += NC) {
        const int n = g.ord4[k], e0 = g.in_ptr3[n];
        sdesc[k] = make_int2(n | ((g.in_ptr3[n + 1] - e0) << 16), e0);
        wself[k] = g.wself_d[k];
      }
    }
  }
  // (published to the other consumers by the barrier that ends phase A of the first tile)
  const int lane = tid & 31;
  const bool self_in = g.in_self3 != 0;
  const int col = tid & 7;                   // items advance by NC (a multiple of 8): a thread keeps its column
  // the node of this thread in the per-node steps: spread over all warps (every (NC / N)-th thread)
  const int nstep = NC / N > 0 ? NC / N : 1;
  const int node = (tid % nstep == 0 && tid / nstep < N) ? tid / nstep : -1;
  int k = 0;
  for (int tl = blockIdx.x; tl < a.total; tl += gridDim.x, ++k) {
    const int s = k % a.nstage, ph = (k / a.nstage) & 1;
    const int b = tl / a.tiles, c0 = (tl - b * a.tiles) * kCB4;
    const int c = c0 + col;                  // this thread's chunk column in the window
    const bool cok = c < C4;
    float4* tR = tiles + (size_t)s * 2 * tile_f4;
    float4* tP = tR + tile_f4;
    // deferred x update: this thread's chunks of x are requested before the wait for the tile (N <= NC: at most 8 items)
    constexpr int kXI = 8;
    float4 xv[kXI];
    const bool xdef = SRC == 0 && a.xd_out != nullptr;
    const size_t xw0 = (size_t)b * (size_t)(N * C4) + c;
    if (xdef && cok) {
      const float4* xi = reinterpret_cast<const float4*>(a.xd_in) + xw0;
#pragma unroll
      for (int j = 0; j < kXI; ++j) {
        const int item = tid + j * NC;
        if (item < N * kCB4) xv[j] = __ldcs(xi + (size_t)(item >> 3) * C4);
      }
    }
    // ... and the chunks of the CTA's NEXT tile are pulled into L2 now: the loads above are consumed right behind the wait
    // (the tile itself is staged a whole item ahead), so an HBM round trip per tile would be exposed - measured +34 us per launch
    if (xdef && tl + (int)gridDim.x < a.total) {
      const int tn = tl + (int)gridDim.x, bn = tn / a.tiles, cn = (tn - bn * a.tiles) * kCB4 + col;
      if (cn < C4) {
        const float4* xn = reinterpret_cast<const float4*>(a.xd_in) + (size_t)bn * (size_t)(N * C4) + cn;
#pragma unroll
        for (int j = 0; j < kXI; ++j) {
          const int item = tid + j * NC;
          if (item < N * kCB4) asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + (size_t)(item >> 3) * C4));
        }
      }
    }
    mbar_wait(full + s, ph);
    const float beta = SRC == 0 ? beta_s[s] : 0.f;

    // ---- phase A: p' = r + beta p, in place (and x += alpha_prev p with the old p, before phase B overwrites its tile)
    if (SRC == 0 && !(MGA_K4_X & 32)) {
      if (xdef) {
        const float al = beta_s[a.nstage + s];
        float4* xo = reinterpret_cast<float4*>(a.xd_out) + xw0;
#pragma unroll
        for (int j = 0; j < kXI; ++j) {
          const int item = tid + j * NC;
          if (item < N * kCB4) {
            float4 v = tR[item];
            const float4 q = tP[item];
            v.x += beta * q.x; v.y += beta * q.y; v.z += beta * q.z; v.w += beta * q.w;
            tR[item] = v;
            if (SYS == 0 && col == 0) p0[item >> 3] = v.x;
            if (cok) {
              float4 u = xv[j];
              u.x += al * q.x; u.y += al * q.y; u.z += al * q.z; u.w += al * q.w;
              __stcs(xo + (size_t)(item >> 3) * C4, u);
            }
          }
        }
      } else {
        for (int item = tid; item < N * kCB4; item += NC) {
          float4 v = tR[item];
          const float4 q = tP[item];
          v.x += beta * q.x; v.y += beta * q.y; v.z += beta * q.z; v.w += beta * q.w;
          tR[item] = v;
          if (SYS == 0 && col == 0) p0[item >> 3] = v.x;
        }
      }
    } else if (SYS == 0 && col == 0) {
      for (int item = tid; item < N * kCB4; item += NC) p0[item >> 3] = tR[item].x;
    }
    if (want_halo) {
      mbar_wait(hfull, k & 1);
      if (node >= 0) {
        const float* hb = reinterpret_cast<const float*>(hbuf);
        const float rl = hb[node * 4 + 3], rr = hb[((size_t)a.rows_tile + node) * 4];      // last of the chunk before, first of the chunk after
        if (SRC == 0) {
          hl[node] = rl + beta * hb[((size_t)2 * a.rows_tile + node) * 4 + 3];
          hr[node] = rr + beta * hb[((size_t)3 * a.rows_tile + node) * 4];
        } else {
          hl[node] = rl;
          hr[node] = rr;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(hempty);
    } else if (SYS == 0 && node >= 0) {
      hl[node] = 0.f;
      hr[node] = 0.f;
    }
    fence_async_smem();                      // generic-proxy writes of p' -> visible to the TMA store
    bar_cons<NC>();
    if (SRC != 2 && tid == 0 && !(MGA_K4_X & 4)) {   // p' -> HBM (rows beyond N / columns beyond the row are clipped by the map)
      for (int j = 0; j < a.nbox; ++j) tma_store3(&maps.pnew, tR + (size_t)j * a.rows_box * kCB4, 4 * c0, j * a.rows_box, b);
      tma_commit();
    }

    const size_t w0 = (size_t)b * (size_t)(N * C4);
    const int tt = 4 * c;
    float dot = 0.f;
    if (SYS == 0) {
      // ---- phase B: qs[t] = q[t+1] = p'[t+1] - sum_j w_j p'_nbr[t]  (ADMM.py:166-177), into the dead p tile
      {
        const char* mine = reinterpret_cast<const char*>(tR + col);
        const bool v1 = tt + 1 < T, v2 = tt + 2 < T, v3 = tt + 3 < T, v4 = tt + 4 < T;
        for (int it0 = tid - lane; it0 < N * kCB4; it0 += NC) {              // warp-uniform trips: the shuffle needs every lane
          const bool on = it0 + lane < N * kCB4;
          const int item = on ? it0 + lane : N * kCB4 - 1;
          const int n = item >> 3;
          const float4 own = tR[item];
          const float up = __shfl_down_sync(0xffffffffu, own.x, 1);          // the next chunk of the row is the next lane's
          const float nxt = col < kCB4 - 1 ? up : hr[n];
          const float ws = wself[n];
          float4 acc = make_float4(ws * own.x, ws * own.y, ws * own.z, ws * own.w);
          if (!(MGA_K4_X & 2)) acc = gather3<K>(tab + n * kf, kf, mine, acc);
          float4 o;
          o.x = v1 ? own.y - acc.x : 0.f;
          o.y = v2 ? own.z - acc.y : 0.f;
          o.z = v3 ? own.w - acc.z : 0.f;
          o.w = v4 ? nxt - acc.w : 0.f;
          if (on) tP[item] = o;
        }
        // q at the tile's first step, one scalar gather per node: q[4 c0] = p'[4 c0] - sum_j w_j p'_nbr[4 c0 - 1]; q[0] = 0 (ADMM.py:176)
        if (node >= 0) {
          float qv = 0.f;
          if (c0 > 0 && 4 * c0 < T && !(MGA_K4_X & 64)) {
            float acc = wself[node] * hl[node];
            const int2* row = tab + node * kf;
            for (int j = 0; j < kf; ++j) {
              const int2 e = row[j];
              acc += __int_as_float(e.y) * hl[e.x >> 7];             // entry = (row * 128 bytes, weight)
            }
            qv = p0[node] - acc;
          }
          qsl[node] = qv;
        }
      }
      bar_cons<NC>();
      // ---- phase C: Ap = D p' + c (q - f), f = in-list gather of qs (ADMM.py:200-209 as a gather, scatter order kept)
      {
        const char* mine = reinterpret_cast<const char*>(tP + col);
        const float4* rw = reinterpret_cast<const float4*>(a.rhs) + w0 + c;
        float4* ow = reinterpret_cast<float4*>(a.out) + w0 + c;
        float hx[4], tv[4];                  // H^T H keeps rows t < t_in (ADMM.py:372-374); pads (t >= T) stay 0
#pragma unroll
        for (int j = 0; j < 4; ++j) { hx[j] = (a.xsys && tt + j < g.t_in) ? 1.f : 0.f; tv[j] = tt + j < T ? 1.f : 0.f; }
        for (int it0 = tid - lane; it0 < N * kCB4; it0 += NC) {
          const bool on = cok && it0 + lane < N * kCB4;                     // (columns beyond the row: zero-filled tile, nothing to store)
          const int item = it0 + lane < N * kCB4 ? it0 + lane : N * kCB4 - 1;
          const int2 sd = sdesc[item >> 3];
          const int n = sd.x & 0xffff;
          const int idx = n * kCB4 + col;
          const float4 pv = tR[idx];
          float4 rh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (SRC == 2 && on) rh = __ldcs(rw + (size_t)n * C4);
          const float4 q1 = tP[idx];
          const float dn = __shfl_up_sync(0xffffffffu, q1.w, 1);
          const float qprev = col > 0 ? dn : qsl[n];
          const float ws = self_in ? wself[n] : 0.f;
          float4 f = make_float4(ws * q1.x, ws * q1.y, ws * q1.z, ws * q1.w);
          int e = sd.y;
          const int e1 = e + (sd.x >> 16);
          if (!(MGA_K4_X & 1)) {
            for (; e + 4 <= e1; e += 4) f = gather3<4>(tab_in + e, 4, mine, f);
            if (e + 2 <= e1) { f = gather3<2>(tab_in + e, 2, mine, f); e += 2; }
            if (e < e1) f = gather3<1>(tab_in + e, 1, mine, f);
          }
          const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
          const float q[4] = {qprev, q1.x, q1.y, q1.z};
          const float ff[4] = {f.x, f.y, f.z, f.w};
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float l = q[j] - ff[j];    // row T-1: f = 0 because qs[T-1] = 0; Q1 is moot as q[0] = 0
            float val;
            if (a.xsys) val = (hx[j] * pp[j] + a.a * pp[j]) + a.cc * l;      // ADMM.py:372-379
            else val = a.cc * l + a.a * pp[j];                               // ADMM.py:394
            o[j] = tv[j] * val;
          }
          if (!on) continue;
          if (SRC != 2) {
            if (!(MGA_K4_X & 8)) __stcs(ow + (size_t)n * C4, make_float4(o[0], o[1], o[2], o[3]));
            dot += (pp[0] * o[0] + pp[1] * o[1]) + (pp[2] * o[2] + pp[3] * o[3]);
          } else {
            const float r0 = rh.x - o[0], r1 = rh.y - o[1], r2 = rh.z - o[2], r3 = rh.w - o[3];
            __stcs(ow + (size_t)n * C4, make_float4(r0, r1, r2, r3));
            dot += (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
          }
        }
      }
    } else {
      // ---- z_u system: Ap = c (p' - sum_j w_j p'_nbr) + a p'   (ADMM.py:138-148, 389-390)
      if (cok) {
        const char* mine = reinterpret_cast<const char*>(tR + col);
        const float4* rw = reinterpret_cast<const float4*>(a.rhs) + w0 + c;
        float4* ow = reinterpret_cast<float4*>(a.out) + w0 + c;
        for (int item = tid; item < N * kCB4; item += NC) {
          const int n = item >> 3;
          float4 rh;
          if (SRC == 2) rh = __ldcs(rw + (size_t)n * C4);
          const float4 pv = tR[item];
          const float4 acc = gather3<K>(tab + n * kf, kf, mine, make_float4(0.f, 0.f, 0.f, 0.f));
          float4 o;
          o.x = a.cc * (pv.x - acc.x) + a.a * pv.x;
          o.y = a.cc * (pv.y - acc.y) + a.a * pv.y;
          o.z = a.cc * (pv.z - acc.z) + a.a * pv.z;
          o.w = a.cc * (pv.w - acc.w) + a.a * pv.w;    // pads: v = 0 and every gathered pad is 0
          if (SRC != 2) {
            __stcs(ow + (size_t)n * C4, o);
            dot += (pv.x * o.x + pv.y * o.y) + (pv.z * o.z + pv.w * o.w);
          } else {
            const float4 rr = make_float4(rh.x - o.x, rh.y - o.y, rh.z - o.z, rh.w - o.w);
            __stcs(ow + (size_t)n * C4, rr);
            dot += (rr.x * rr.x + rr.y * rr.y) + (rr.z * rr.z + rr.w * rr.w);
          }
        }
      }
    }
    // ---- per-warp: dot product of the tile and release of the stage (no CTA-wide barrier: the next tile's phase A only
    // touches its own stage and hl / hr / p0, which nobody reads after the barrier that ended phase B)
    dot = warp_sum<float>(dot);
    if (SRC != 2 && tid == 0) tma_wait_read();       // the p' store has read its tile
    __syncwarp();
    if (lane == 0) {
      atomicAdd(a.slot + b, (double)dot);
      mbar_arrive(empty + s);
    }
  }
  if (SRC != 2 && tid == 0) tma_wait_all();          // stores complete before the grid ends
}



// ---------------------------------------------------------------------------------------------------------------------
// k5_tail: the tail of one outer iteration (ADMM.py:595-637: both dual ascents, phi prox, gamma ascent, every
// diagnostics sum) with the x tile staged by TMA - what k2_tail does with L1 / L2 gathers (2.4 TB/s of its 1.36 GB).
// Only x is gathered (L_d x needs the neighbours one step back, L_u x the neighbours at the same step), so only x goes
// through shared memory: tile {32 floats, node rows} + the chunk left of it (halo), two stages, producer warp as in
// k4_cg.  The other nine vectors are read and the four results written straight from / to HBM by the thread that owns
// the chunk (coalesced 128-bit accesses).  L_d is unaligned in time (x[t-1] of the neighbour): the element before a
// chunk is the previous lane's last element of the SAME neighbour's row (the 8 lanes of a row share the table row),
// so it comes by shuffle; only the first column reads the halo.
// Items are dealt to the CTAs as contiguous ranges of (time tile, window) with the time tile major: a CTA stays on one
// time tile for ~B * tiles / gridDim windows and adds up x - x_old over them in registers - one double atomic per
// lattice point and CTA for the (T, N) batch sum instead of one per point and window.
struct K5Maps {
  CUtensorMap x, x_halo;
};
struct K5Args {
  int64_t B;
  int nstage, rows_box, nbox, rows_tile, tiles, total;
  int want_diag;
  const float *x, *x_old, *zu, *zu_old, *zd, *zd_old;
  float *gu, *gd, *gam, *phi;
  const float* y;
  float rho, rho_u, rho_d, thr;
  double* diag;
  double* dx_sum;
};

constexpr int kK5Slots = 8;      // items per consumer thread and tile: N * 8 / NC <= 8

__device__ __forceinline__ float soft5(float s, float d) {
  const float u = fabsf(s) - d;
  const float sg = (float)((s > 0.f) - (s < 0.f));
  return sg * u * (float)(u > 0.f);   // ADMM.py:407-408
}

template <int K, int NC>
__global__ void __launch_bounds__(NC + 32, 1) k5_tail(const __grid_constant__ K5Maps maps, const Graph2 g, const K5Args a) {
  extern __shared__ __align__(128) unsigned char smem5[];
  const int N = g.N, C4 = g.C4, T = g.T;
  const int tid = threadIdx.x;
  const size_t tile_f4 = (size_t)a.rows_tile * kCB4;
  float4* tiles = reinterpret_cast<float4*>(smem5);                  // [stage][rows_tile * 8]
  float4* hbuf = tiles + tile_f4 * a.nstage;                         // [stage][rows_tile]: the chunk left of the tile
  float* wself = reinterpret_cast<float*>(hbuf + (size_t)a.rows_tile * a.nstage);
  int2* tab_d = reinterpret_cast<int2*>(wself + N + (N & 1));
  int2* tab_u = tab_d + (size_t)N * g.kd3;
  uint64_t* full = reinterpret_cast<uint64_t*>(tab_u + (size_t)N * g.ku3);
  uint64_t* empty = full + a.nstage;
  float* dred = reinterpret_cast<float*>(empty + a.nstage);          // MGA_DIAG_COLS x 32
  if (tid == 0) {
    for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  __syncthreads();
  // contiguous range of items for this CTA; item id = time tile * B + window
  const int per = (a.total + gridDim.x - 1) / gridDim.x;
  const int first = blockIdx.x * per, last = min(a.total, first + per);
  const uint32_t stage_bytes = (uint32_t)(tile_f4 * 16 + (size_t)a.rows_tile * 16);
  if (tid >= NC) {
    if (tid == NC) {
      int k = 0;
      for (int id = first; id < last; ++id, ++k) {
        const int s = k % a.nstage, ph = (k / a.nstage) & 1;
        const int ct = id / (int)a.B, b = id - ct * (int)a.B, c0 = ct * kCB4;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, stage_bytes);
        for (int j = 0; j < a.nbox; ++j) {
          tma_load3(tiles + (size_t)s * tile_f4 + (size_t)j * a.rows_box * kCB4, &maps.x, 4 * c0, j * a.rows_box, b, full + s);
          tma_load3(hbuf + (size_t)s * a.rows_tile + j * a.rows_box, &maps.x_halo, 4 * c0 - 4, j * a.rows_box, b, full + s);
        }
      }
    }
    return;
  }
  for (int k = tid; k < N * g.kd3; k += NC) tab_d[k] = g.tab_d[k];
  for (int k = tid; k < N * g.ku3; k += NC) tab_u[k] = g.tab_u[k];
  for (int k = tid; k < N; k += NC) wself[k] = g.wself_d[k];
  asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
  const int lane = tid & 31, col = tid & 7;
  float d[MGA_DIAG_COLS];
#pragma unroll
  for (int c = 0; c < MGA_DIAG_COLS; ++c) d[c] = 0.f;
  float4 dxacc[kK5Slots];
#pragma unroll
  for (int j = 0; j < kK5Slots; ++j) dxacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  int ct_acc = -1;
  auto flush_dx = [&](int ct) {          // x - x_old summed over this CTA's windows of time tile ct -> the (T, N) batch sum
    if (ct < 0 || !a.dx_sum) return;
    const int c = ct * kCB4 + col;
#pragma unroll
    for (int j = 0; j < kK5Slots; ++j) {
      const int item = tid + j * NC;
      if (item < N * kCB4 && c < C4) {
        const int o = g.perm[item >> 3];
        const float v[4] = {dxacc[j].x, dxacc[j].y, dxacc[j].z, dxacc[j].w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (4 * c + q < T) atomicAdd(a.dx_sum + (size_t)(4 * c + q) * N + o, (double)v[q]);
      }
      dxacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  int k = 0;
  for (int id = first; id < last; ++id, ++k) {
    const int s = k % a.nstage, ph = (k / a.nstage) & 1;
    const int ct = id / (int)a.B, b = id - ct * (int)a.B, c0 = ct * kCB4;
    if (ct != ct_acc) { flush_dx(ct_acc); ct_acc = ct; }
    const int c = c0 + col;
    const bool cok = c < C4;
    const float4* tX = tiles + (size_t)s * tile_f4;
    const float* hX = reinterpret_cast<const float*>(hbuf + (size_t)s * a.rows_tile);
    const size_t w0 = (size_t)b * (size_t)(N * C4);
    const int t0 = 4 * c;
    mbar_wait(full + s, ph);
    const char* mine = reinterpret_cast<const char*>(tX + col);
#pragma unroll
    for (int j = 0; j < kK5Slots; ++j) {
      const int it0 = tid - lane + j * NC;
      if (it0 >= N * kCB4) break;                                    // warp-uniform
      const bool on = cok && it0 + lane < N * kCB4;
      const int item = it0 + lane < N * kCB4 ? it0 + lane : N * kCB4 - 1;
      const int n = item >> 3;
      const float4 xv4 = tX[item];
      // ---- L_d x at the chunk's four steps: x[t] - sum_j w_j x_nbr[t-1]  (ADMM.py:166-177)
      float acc[4];
      {
        const float ws = wself[n];
        const float up = __shfl_up_sync(0xffffffffu, xv4.w, 1);
        const float prev = col > 0 ? up : hX[n * 4 + 3];
        acc[0] = ws * prev; acc[1] = ws * xv4.x; acc[2] = ws * xv4.y; acc[3] = ws * xv4.z;
        const int2* row = tab_d + n * g.kd3;
        int2 e[K > 0 ? K : 1];
        if (K > 0) {
#pragma unroll
          for (int q = 0; q < K; ++q) e[q] = row[q];
#pragma unroll
          for (int q = 0; q < K; ++q) {
            const float wj = __int_as_float(e[q].y);
            const float4 nb = *reinterpret_cast<const float4*>(mine + e[q].x);
            const float nup = __shfl_up_sync(0xffffffffu, nb.w, 1);
            const float np = col > 0 ? nup : hX[(e[q].x >> 7) * 4 + 3];
            acc[0] += wj * np; acc[1] += wj * nb.x; acc[2] += wj * nb.y; acc[3] += wj * nb.z;
          }
        } else {
          for (int q = 0; q < g.kd3; ++q) {
            const int2 en = row[q];
            const float wj = __int_as_float(en.y);
            const float4 nb = *reinterpret_cast<const float4*>(mine + en.x);
            const float nup = __shfl_up_sync(0xffffffffu, nb.w, 1);
            const float np = col > 0 ? nup : hX[(en.x >> 7) * 4 + 3];
            acc[0] += wj * np; acc[1] += wj * nb.x; acc[2] += wj * nb.y; acc[3] += wj * nb.z;
          }
        }
      }
      const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
      float ldx[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) ldx[q] = (t0 + q >= 1 && t0 + q < T) ? xv[q] - acc[q] : 0.f;
      float lux[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.want_diag) {
        const float4 au = gather3<K>(tab_u + n * g.ku3, g.ku3, mine, make_float4(0.f, 0.f, 0.f, 0.f));
        lux[0] = xv[0] - au.x; lux[1] = xv[1] - au.y; lux[2] = xv[2] - au.z; lux[3] = xv[3] - au.w;
      }
      if (!on) continue;
      // ---- the elementwise part, as k2_tail
      const size_t gq = w0 + (size_t)n * C4 + c;
      const float4 zu4 = ld4s(a.zu, gq), zd4 = ld4s(a.zd, gq), gu4 = ld4s(a.gu, gq), gd4 = ld4s(a.gd, gq);
      const float4 ga4 = ld4s(a.gam, gq), ph4 = ld4s(a.phi, gq);
      const float zuv[4] = {zu4.x, zu4.y, zu4.z, zu4.w}, zdv[4] = {zd4.x, zd4.y, zd4.z, zd4.w};
      float guv[4] = {gu4.x, gu4.y, gu4.z, gu4.w}, gdv[4] = {gd4.x, gd4.y, gd4.z, gd4.w};
      float gav[4] = {ga4.x, ga4.y, ga4.z, ga4.w}, phv[4] = {ph4.x, ph4.y, ph4.z, ph4.w};
      float xo[4] = {0.f, 0.f, 0.f, 0.f}, zuo[4] = {0.f, 0.f, 0.f, 0.f}, zdo[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.want_diag) {
        const float4 p1 = ld4s(a.x_old, gq), p2 = ld4s(a.zu_old, gq), p3 = ld4s(a.zd_old, gq);
        xo[0] = p1.x; xo[1] = p1.y; xo[2] = p1.z; xo[3] = p1.w;
        zuo[0] = p2.x; zuo[1] = p2.y; zuo[2] = p2.z; zuo[3] = p2.w;
        zdo[0] = p3.x; zdo[1] = p3.y; zdo[2] = p3.z; zdo[3] = p3.w;
      }
      const int o = g.perm[n];
      int bad = 0;
      float dxv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int t = t0 + q;
        if (t >= T) continue;
        guv[q] = guv[q] + a.rho_u * (xv[q] - zuv[q]);
        gdv[q] = gdv[q] + a.rho_d * (xv[q] - zdv[q]);
        const float ph = soft5(ldx[q] - gav[q] / a.rho, a.thr);
        const float gn = gav[q] + a.rho * (ph - ldx[q]);
        bad |= !isfinite(xv[q]) || !isfinite(zuv[q]) || !isfinite(zdv[q]) || !isfinite(ph) || !isfinite(gn);
        if (a.want_diag) {
          const float dx = xv[q] - xo[q];
          d[MGA_DIAG_DX2] += dx * dx;
          dxv[q] = dx;
          const float e1 = xv[q] - zuv[q], bz = zuv[q] - zuo[q];
          d[MGA_DIAG_X_ZU2] += e1 * e1;
          d[MGA_DIAG_DZU2] += bz * bz;
          d[MGA_DIAG_GLR] += xv[q] * lux[q];
          if (t < g.t_in) {
            const float h = xv[q] - a.y[((size_t)b * g.t_in + t) * N + o];
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
      d[MGA_DIAG_NONFINITE] += (float)bad;
      dxacc[j].x += dxv[0]; dxacc[j].y += dxv[1]; dxacc[j].z += dxv[2]; dxacc[j].w += dxv[3];
      st4s(a.gu, gq, make_float4(guv[0], guv[1], guv[2], guv[3]));
      st4s(a.gd, gq, make_float4(gdv[0], gdv[1], gdv[2], gdv[3]));
      st4s(a.gam, gq, make_float4(gav[0], gav[1], gav[2], gav[3]));
      st4s(a.phi, gq, make_float4(phv[0], phv[1], phv[2], phv[3]));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
  flush_dx(ct_acc);
  // diagnostics: one double atomic per CTA and column
  {
    const int w = tid >> 5;
#pragma unroll
    for (int c = 0; c < MGA_DIAG_COLS; ++c) {
      const float v = warp_sum<float>(d[c]);
      if (lane == 0) dred[c * 32 + w] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NC) : "memory");
    if (tid < MGA_DIAG_COLS) {
      float t = 0.f;
      for (int q = 0; q < NC / 32; ++q) t += dred[tid * 32 + q];
      if (t != 0.f) atomicAdd(a.diag + tid, (double)t);
    }
  }
}

}  // namespace mga
