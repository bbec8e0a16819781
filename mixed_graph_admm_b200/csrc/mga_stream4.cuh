// Streaming mode, fused CG phase 1 with TMA-staged tiles ("k4"): included by mga_stream2.cu.
//
// One kernel per CG iteration does what k3_p_ldr + k3_ldrt_lhs (x / z_d systems) or k3_lu (z_u system) did:
//     (r, p) -> p' = r + beta p,  Ap = A p',  <p', Ap>                      (ADMM.py:348-358, operators 138-228)
// so the intermediate qs = shifted L_d p' never goes to HBM: 16 B per lattice point (+ halo) instead of 28, and with
// k2_xr (24 B/pt) a CG iteration of the 2-hop systems moves 40 B/pt in 2 launches (k3: 52 in 3; algorithmic: 48).
//
// A persistent CTA per SM walks (window, time tile) items.  A tile is ALL nodes x 8 chunks (32 time steps) of r and p:
//   * a producer warp (one elected lane) stages the next tile while the 16 consumer warps work on the current one:
//     cp.async.bulk.tensor (TMA) 3-D boxes {32 floats, <= 256 node rows, 1 window} of a (T, N, B) tensor map, completion
//     on an mbarrier (expect-tx); out-of-range rows / columns are zero-filled by the TMA unit, so partial tiles and the
//     rounding of the node rows to whole boxes need no code.  Two stages (full / empty barrier pairs).
//   * phase A: p' = r + beta p in place in the r tile; one thread then hands the tile to a TMA store (p' -> HBM),
//   * phase B: qs = shifted L_d p' (gathers from the p' tile, table entries = (row byte offset, weight)) overwrites the
//     dead p tile,
//   * phase C: Ap = D p' + c (q - in-list gather of qs), streamed to HBM, <p', Ap> reduced per tile.
//   L_d^T L_d couples t-1, t, t+1, so a tile needs p' one step beyond each edge, for every node: those two halo
//   scalars per node are fetched by the consumer threads one tile ahead (plain loads, registers), q at the tile's
//   first step is then one scalar gather per node.
// Thread = one 16-byte chunk (4 time steps of one node); the 8 lanes of a quarter-warp own the 8 chunks of one row, so a
// neighbour's chunk gather is a conflict-free 128-byte shared-memory wavefront.
#pragma once
#include <cuda.h>

namespace mga {

constexpr int kCons4 = 512;                // consumer threads (16 warps)
constexpr int kThreads4 = kCons4 + 32;     // + the producer warp
constexpr int kCB4 = 8;                    // chunks per tile row

struct K4Args {
  int64_t B;
  int it;
  int nstage, rows_box, nbox, rows_tile;   // rows_tile = rows_box * nbox >= N
  int tiles, total;                        // time tiles per window, B * tiles
  const float* v_r;      // the vector behind map_r: r (SRC 0 / 1) or x0 (SRC 2) - halo scalars
  const float* v_p;      // p_old (SRC 0)
  const float* rhs;      // SRC 2
  float* out;            // Ap (SRC 0 / 1) or r = rhs - A x0 (SRC 2)
  const double* dots;    // RR(k) = dots[2k], PAP(k) = dots[2k+1], each (B)
  double* slot;          // (B): where this launch's dot product goes
  float a, cc;
  int xsys;
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
__device__ __forceinline__ void bar_cons() { asm volatile("bar.sync 1, %0;" ::"n"(kCons4) : "memory"); }

// halo scalars of tile (b, c0) for node n: the element before the tile's first chunk and the one after its last
struct Halo4 {
  float rl, rr, pl, pr;
};
template <int SRC>
__device__ __forceinline__ Halo4 halo_fetch(const Graph2& g, const K4Args& a, int b, int c0, int n) {
  Halo4 h{0.f, 0.f, 0.f, 0.f};
  const size_t row = ((size_t)b * g.N + n) * (size_t)g.C4;
  if (c0 > 0) {
    const size_t k = (row + c0) * 4 - 1;
    h.rl = __ldg(a.v_r + k);
    if (SRC == 0) h.pl = __ldg(a.v_p + k);
  }
  if (c0 + kCB4 < g.C4) {
    const size_t k = (row + c0 + kCB4) * 4;
    h.rr = __ldg(a.v_r + k);
    if (SRC == 0) h.pr = __ldg(a.v_p + k);
  }
  return h;
}

// SYS 0: A = diag + c L_d^T L_d (x / z_d systems), SYS 1: A = c L_u + a I (z_u system)
// SRC 0: v = r + beta p (stored as the new p)   SRC 1: v = r (first iteration; stored as p)   SRC 2: v = x0, out = rhs - A v
// K: compile-time width of the forward table (0: run-time)
template <int SYS, int SRC, int K>
__global__ void __launch_bounds__(kThreads4, 1)
k4_cg(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_p,
      const __grid_constant__ CUtensorMap map_pnew, const Graph2 g, const K4Args a) {
  extern __shared__ __align__(128) unsigned char smem4[];
  const int N = g.N, C4 = g.C4, T = g.T;
  const int tid = threadIdx.x;
  const size_t tile_f4 = (size_t)a.rows_tile * kCB4;                 // float4 per tile buffer
  float4* tiles = reinterpret_cast<float4*>(smem4);                  // [stage][R | P][rows_tile * 8]
  float* hl = reinterpret_cast<float*>(tiles + 2 * tile_f4 * a.nstage);
  float* hr = hl + N;
  float* qsl = hr + N;
  float* wself = qsl + N;
  int* ptr = reinterpret_cast<int*>(wself + N);                      // (N + 1) in-list offsets        [SYS 0]
  int* ord = ptr + N + 1;                                            // (N) row order of phase C         [SYS 0]
  const int kf = SYS == 0 ? g.kd3 : g.ku3;
  int2* tab = reinterpret_cast<int2*>(ord + N + ((2 * N + 1) & 1));  // forward table (N, kf), 8-byte aligned (4 N floats + 2 N + 1 ints before it)
  int2* tab_in = tab + (size_t)N * kf;                               // in-list entries                  [SYS 0]
  const int n_in = SYS == 0 ? g.in_ptr3_total : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tab_in + n_in);       // full[nstage], empty[nstage]
  float* red = reinterpret_cast<float*>(bars + 2 * a.nstage);
  uint64_t* full = bars;
  uint64_t* empty = bars + a.nstage;

  // ---- one-time staging of the graph tables (all threads), barrier init
  {
    const int2* src = SYS == 0 ? g.tab_d : g.tab_u;
    for (int k = tid; k < N * kf; k += kThreads4) tab[k] = src[k];
    if (SYS == 0) {
      for (int k = tid; k < n_in; k += kThreads4) tab_in[k] = g.tab_in3[k];
      for (int k = tid; k <= N; k += kThreads4) ptr[k] = g.in_ptr3[k];
      for (int k = tid; k < N; k += kThreads4) { ord[k] = g.ord3[k]; wself[k] = g.wself_d[k]; }
    }
    if (tid == 0) {
      for (int s = 0; s < a.nstage; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    __syncthreads();
  }
  const uint32_t stage_bytes = (uint32_t)(tile_f4 * 16) * (SRC == 0 ? 2u : 1u);

  if (tid >= kCons4) {
    // ================= producer warp: one lane keeps the stages full =================
    if (tid == kCons4) {
      int k = 0;
      for (int tl = blockIdx.x; tl < a.total; tl += gridDim.x, ++k) {
        const int s = k % a.nstage, ph = (k / a.nstage) & 1;
        const int b = tl / a.tiles, c0 = (tl - b * a.tiles) * kCB4;
        mbar_wait(empty + s, ph ^ 1);                                // first pass over the stages: free at once
        mbar_expect_tx(full + s, stage_bytes);
        float4* tR = tiles + (size_t)s * 2 * tile_f4;
        float4* tP = tR + tile_f4;
        for (int j = 0; j < a.nbox; ++j) {
          tma_load3(tR + (size_t)j * a.rows_box * kCB4, &map_r, 4 * c0, j * a.rows_box, b, full + s);
          if (SRC == 0) tma_load3(tP + (size_t)j * a.rows_box * kCB4, &map_p, 4 * c0, j * a.rows_box, b, full + s);
        }
      }
    }
    return;
  }

  // ================= consumers =================
  const int lane = tid & 31, wp = tid >> 5;
  const bool self_in = g.in_self3 != 0;
  const int col = tid & 7;                   // items advance by kCons4 (a multiple of 8): a thread keeps its column
  Halo4 hn{0.f, 0.f, 0.f, 0.f};
  float beta_n = 0.f;
  auto prefetch = [&](int tl) {              // halo scalars and beta of tile `tl`, one tile ahead of their use
    if (tl >= a.total) return;
    const int b = tl / a.tiles, c0 = (tl - b * a.tiles) * kCB4;
    if (tid < N) hn = halo_fetch<SRC>(g, a, b, c0, tid);
    if (SRC == 0) beta_n = (float)a.dots[(size_t)(2 * a.it) * a.B + b] / (float)a.dots[(size_t)(2 * a.it - 2) * a.B + b];   // ADMM.py:356
  };
  prefetch(blockIdx.x);
  int k = 0;
  for (int tl = blockIdx.x; tl < a.total; tl += gridDim.x, ++k) {
    const int s = k % a.nstage, ph = (k / a.nstage) & 1;
    const int b = tl / a.tiles, c0 = (tl - b * a.tiles) * kCB4;
    const int c = c0 + col;                  // this thread's chunk column in the window
    const bool cok = c < C4;
    float4* tR = tiles + (size_t)s * 2 * tile_f4;
    float4* tP = tR + tile_f4;
    const Halo4 h = hn;
    const float beta = beta_n;
    prefetch(tl + gridDim.x);
    mbar_wait(full + s, ph);

    // ---- phase A: p' = r + beta p, in place
    if (SRC == 0) {
      for (int item = tid; item < N * kCB4; item += kCons4) {
        float4 v = tR[item];
        const float4 q = tP[item];
        v.x += beta * q.x; v.y += beta * q.y; v.z += beta * q.z; v.w += beta * q.w;
        tR[item] = v;
      }
    }
    if (SYS == 0 && tid < N) {
      hl[tid] = SRC == 0 ? h.rl + beta * h.pl : h.rl;
      hr[tid] = SRC == 0 ? h.rr + beta * h.pr : h.rr;
    }
    fence_async_smem();                      // generic-proxy writes of p' -> visible to the TMA store
    bar_cons();
    if (SRC != 2 && tid == 0) {              // p' -> HBM (rows beyond N / columns beyond the row are clipped by the map)
      for (int j = 0; j < a.nbox; ++j) tma_store3(&map_pnew, tR + (size_t)j * a.rows_box * kCB4, 4 * c0, j * a.rows_box, b);
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
        for (int item = tid; item < N * kCB4; item += kCons4) {
          const int n = item >> 3;
          const float4 own = tR[item];
          const float nxt = col < kCB4 - 1 ? tR[item + 1].x : hr[n];
          const float ws = wself[n];
          const float4 acc = gather3<K>(tab + n * kf, kf, mine, make_float4(ws * own.x, ws * own.y, ws * own.z, ws * own.w));
          float4 o;
          o.x = v1 ? own.y - acc.x : 0.f;
          o.y = v2 ? own.z - acc.y : 0.f;
          o.z = v3 ? own.w - acc.z : 0.f;
          o.w = v4 ? nxt - acc.w : 0.f;
          tP[item] = o;
        }
        // q at the tile's first step, one scalar gather per node: q[4 c0] = p'[4 c0] - sum_j w_j p'_nbr[4 c0 - 1]; q[0] = 0 (ADMM.py:176)
        if (tid < N) {
          float qv = 0.f;
          if (c0 > 0 && 4 * c0 < T) {
            float acc = wself[tid] * hl[tid];
            const int2* row = tab + tid * kf;
            for (int j = 0; j < kf; ++j) {
              const int2 e = row[j];
              acc += __int_as_float(e.y) * hl[e.x >> 7];             // entry = (row * 128 bytes, weight)
            }
            qv = tR[tid * kCB4].x - acc;
          }
          qsl[tid] = qv;
        }
      }
      bar_cons();
      // ---- phase C: Ap = D p' + c (q - f), f = in-list gather of qs (ADMM.py:200-209 as a gather, scatter order kept)
      if (cok) {
        const char* mine = reinterpret_cast<const char*>(tP + col);
        const float4* rw = reinterpret_cast<const float4*>(a.rhs) + w0 + c;
        float4* ow = reinterpret_cast<float4*>(a.out) + w0 + c;
        float hx[4], tv[4];                  // H^T H keeps rows t < t_in (ADMM.py:372-374); pads (t >= T) stay 0
#pragma unroll
        for (int j = 0; j < 4; ++j) { hx[j] = (a.xsys && tt + j < g.t_in) ? 1.f : 0.f; tv[j] = tt + j < T ? 1.f : 0.f; }
        for (int item = tid; item < N * kCB4; item += kCons4) {
          const int n = ord[item >> 3];
          const int idx = n * kCB4 + col;
          const float4 pv = tR[idx];
          float4 rh;
          if (SRC == 2) rh = __ldcs(rw + (size_t)n * C4);
          const float4 q1 = tP[idx];
          const float qprev = col > 0 ? tP[idx - 1].w : qsl[n];
          const float ws = self_in ? wself[n] : 0.f;
          float4 f = make_float4(ws * q1.x, ws * q1.y, ws * q1.z, ws * q1.w);
          int e = ptr[n];
          const int e1 = ptr[n + 1];
          for (; e + 4 <= e1; e += 4) f = gather3<4>(tab_in + e, 4, mine, f);
          if (e + 2 <= e1) { f = gather3<2>(tab_in + e, 2, mine, f); e += 2; }
          if (e < e1) f = gather3<1>(tab_in + e, 1, mine, f);
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
          if (SRC != 2) {
            __stcs(ow + (size_t)n * C4, make_float4(o[0], o[1], o[2], o[3]));
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
        for (int item = tid; item < N * kCB4; item += kCons4) {
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
    // ---- per-tile dot product and release of the stage
    dot = warp_sum<float>(dot);
    if (lane == 0) red[wp] = dot;
    if (SRC != 2 && tid == 0) tma_wait_read();       // the p' store has read its tile
    bar_cons();
    if (tid == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kCons4 / 32; ++w) t += red[w];
      atomicAdd(a.slot + b, (double)t);
      mbar_arrive(empty + s);
    }
  }
  if (SRC != 2 && tid == 0) tma_wait_all();          // stores complete before the grid ends
}

}  // namespace mga
