// Cluster mode: the whole of combined_loop (ADMM.py:528-648) for one window inside one thread-block CLUSTER, in the
// signal's own precision (float64 as the reference's PEMS runs, or float32) and with the reference's stop tests decided
// on the device - the reference's own call pattern (B = 1, float64, CG_tol = 1e-8, ADMM_tol = 1e-6, up to 150 outer
// iterations; ADMM.py:76-80) in ONE launch with no host round trip per CG iteration.
//
// Why a cluster: a float64 window of PEMS size (307 nodes x 24 steps: 59 KB per vector, 12 vectors live) does not fit the
// registers / shared memory of one SM, and with the vectors in L2 one SM executes every gather itself (measured: 28 us
// per CG iteration).  Here the T time steps are dealt to the CTAs of a cluster as slabs of TL consecutive steps
// (T = 24 -> 8 CTAs x 3 steps); a thread owns one node at RPT consecutive steps of the slab (PEMS sizes: one step per
// thread, 3 x 320 threads per CTA - with one thread per node and slab the gathers of a step ran one after the other):
//   * CG vectors of the thread's points in registers, the seven ADMM state vectors in shared memory;
//   * the vector other threads gather from (p, then q = L_d p) is staged per CTA as TL + 1 time rows: the slab plus ONE
//     halo row, pushed by the neighbouring CTA through distributed shared memory (L_d reads step t-1, L_d^T step t+1;
//     ADMM.py:171, 200-208) - 2.4 KB per exchange instead of a whole vector;
//   * dot products: warp shuffle -> CTA -> every CTA writes its partial into a slot of EVERY CTA's shared memory ->
//     cluster barrier -> all threads add the slots in the same order, so alpha, beta and the stop decisions are
//     bit-identical across the cluster and the control flow stays uniform (barrier.cluster needs that);
//   * graph tables (forward ELL, in-list CSR in scatter order) staged once per CTA.
// Tolerance mode is per window here, so it is offered for B = 1 (for B > 1 the reference's tests are batch-global,
// quirk Q3: the general kernels keep those); fixed iteration counts run any B, one cluster per window.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdio>

#include "mga_common.cuh"

namespace cg = cooperative_groups;

namespace mga {

constexpr int kClMaxTL = 4;     // time steps per CTA (T <= 8 * 4 with portable clusters)
constexpr int kClMaxCL = 16;    // largest cluster (8 is the portable size; 9..16 need the non-portable opt-in)
constexpr int kClPortable = 8;

template <typename S>
struct ClArgs {
  int N, NP, T, t_in, CL, TL, kd, ku, q1, nnz;      // TL: time steps per CTA
  int transpose_exact;
  int u_wT, d_wT;          // 1: time-invariant weights; T / T-1: one slice per time step (the caller's learned / time-varying tables)
  const int* csr_slot;     // (nnz) index of every in-list entry into a d_w slice (time-varying weights)
  int n_outer, max_cg, want_diag;
  int64_t B;
  const int* nbr_d; const float* d_w; const int* nbr_u; const float* u_w;
  const int* csr_ptr; const int* csr_src; const float* csr_w;
  const int* perm;         // perm[internal node] = caller's node (tables in reverse-Cuthill-McKee order), or NULL
  const S* y; S* x_out;
  const S* mask;           // mask / interpolation mode (ADMM.py:373-376, 783-811): (B, T, N); y then has T rows
  S* out[6];               // zu, zd, phi, gamma, gamma_u, gamma_d (optional)
  double* diag; double* dx_sum;
  S* alpha; S* beta;       // (n_outer, 3, max_cg, B) or NULL
  int* cg_iters;           // (n_outer, 3) device-visible, tolerance mode; NULL otherwise
  int* outer_done;
  double cg_tol, admm_tol;
  S rho, rho_u, rho_d, thr;
  S ax, cx, azu, czu, azd, czd;
  float t_mean, t_var;
};

enum { CS_X = 0, CS_ZU, CS_ZD, CS_GU, CS_GD, CS_GAM, CS_PHI, CS_COUNT };

template <typename S>
__device__ __forceinline__ S cl_soft(S s, S d) {
  const S u = fabs(s) - d;
  const S sg = (S)((s > (S)0) - (s < (S)0));
  return sg * u * (S)(u > (S)0);   // ADMM.py:407-408
}

// CMP (the two-CTAs-per-SM instantiation): compact tables - 16-bit neighbour / source indices, and the in-list reads its
// weights from the forward table's slices through a 16-bit slot index instead of keeping a permuted copy
template <typename S, int RPT, bool CMP = false>
struct ClCtx {
  using I = typename std::conditional<CMP, short, int>::type;
  int N, NP, T, t_in, t0, i, rank, CL, kd, ku, q1;
  int TL, l0;     // time steps per CTA; this thread's first local step (it owns l0 .. l0 + RPT - 1)
  bool active;
  const I* nbr_d; const float* w_d; const I* nbr_u; const float* w_u;
  const int* cptr; const I* csrc; const float* cw;
  const unsigned short* slot;     // CMP: position of every in-list entry in a d_w slice
  S* pbuf;        // (TL + 1) rows x NP: row 0 = the previous CTA's last step (halo), rows 1..TL = own steps
  S* qbuf;        // (TL + 1) rows x NP: rows 0..TL-1 = own steps, row TL = the next CTA's first step (halo)
  S* red;         // 32 warp partials
  S* slots;       // [2][kClMaxCL] cluster reduction slots (a full copy in every CTA), double-buffered
  int parity;
  bool qdot;      // the in-list is the exact transpose of the forward temporal table
  int sd_stride, su_stride, sc_stride;   // floats between the weight slices of consecutive local steps (0: time-invariant)
  cg::cluster_group cl = cg::this_cluster();

  // Sum over the whole window (all CTAs of the cluster), the same bits in every thread of the cluster.
  __device__ __forceinline__ S csum(S v) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum<S>(v);
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (threadIdx.x < (unsigned)CL) {
      S t = 0;
      for (int k = 0; k < nw; ++k) t += red[k];
      S* remote = cl.map_shared_rank(slots, threadIdx.x);      // thread r delivers this CTA's partial to CTA r
      remote[parity * kClMaxCL + rank] = t;
    }
    cl.sync();
    S tot = 0;
    for (int r = 0; r < CL; ++r) tot += slots[parity * kClMaxCL + r];
    parity ^= 1;
    return tot;
  }

  // the same sum in two halves around a cluster barrier the caller has anyway
  __device__ __forceinline__ void csum_post(S v) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum<S>(v);
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (threadIdx.x < (unsigned)CL) {
      S t = 0;
      for (int k = 0; k < nw; ++k) t += red[k];
      cl.map_shared_rank(slots, threadIdx.x)[parity * kClMaxCL + rank] = t;
    }
  }
  __device__ __forceinline__ S csum_read() {
    S tot = 0;
    for (int r = 0; r < CL; ++r) tot += slots[parity * kClMaxCL + r];
    parity ^= 1;
    return tot;
  }

  // K sums at once (one CTA barrier, one cluster barrier): the outer stop test needs six norms
  template <int K>
  __device__ __forceinline__ void csum_vec(S (&v)[K], S* scratch /* K x 32 local */, S* vs /* K x kClMaxCL, written by every CTA */) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    S* part = scratch;                    // [K][32] warp partials
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const S t = warp_sum<S>(v[k]);
      if (lane == 0) part[k * 32 + w] = t;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < CL * K; e += blockDim.x) {      // (a 32-thread CTA in a 16-CTA cluster has fewer threads than CL * K)
      const int r = e / K, k = e - r * K;
      S t = 0;
      for (int q = 0; q < nw; ++q) t += part[k * 32 + q];
      cl.map_shared_rank(vs, r)[k * kClMaxCL + rank] = t;
    }
    cl.sync();
#pragma unroll
    for (int k = 0; k < K; ++k) {
      S tot = 0;
      for (int r = 0; r < CL; ++r) tot += vs[k * kClMaxCL + r];
      v[k] = tot;
    }
    // (`vs` is written again an outer iteration later: dozens of cluster barriers after these reads)
  }

  // own steps -> pbuf rows 1..TL, last own step -> halo row 0 of the next CTA; cluster barrier
  __device__ __forceinline__ void put_p(const S (&v)[RPT]) {
#pragma unroll
    for (int m = 0; m < RPT; ++m) pbuf[(l0 + m + 1) * NP + i] = v[m];
    if (rank + 1 < CL && l0 + RPT == TL) cl.map_shared_rank(pbuf, rank + 1)[i] = v[RPT - 1];
  }
  __device__ __forceinline__ void publish_p(const S (&v)[RPT]) {
    put_p(v);
    cl.sync();
  }
  // own steps -> qbuf rows 0..TL-1, first own step -> halo row TL of the previous CTA; cluster barrier
  __device__ __forceinline__ void publish_q(const S (&v)[RPT]) {
#pragma unroll
    for (int m = 0; m < RPT; ++m) qbuf[(l0 + m) * NP + i] = v[m];
    if (rank > 0 && l0 == 0) cl.map_shared_rank(qbuf, rank - 1)[TL * NP + i] = v[0];
    cl.sync();
  }
  // q = L_d v with v published in pbuf (ADMM.py:166-177): q[t] = v[t] - sum_j w_j v[t-1, nbr_j], q[0] = 0.
  // Neighbour-major: one table entry per neighbour serves all the thread's steps.
  __device__ __forceinline__ void ldr(const S (&v)[RPT], S (&q)[RPT]) const {
    S acc[RPT];
#pragma unroll
    for (int m = 0; m < RPT; ++m) acc[m] = 0;
    if (active) {
      const S* prev = pbuf + l0 * NP;                  // row l0 + m holds step t - 1 of the thread's m-th step
      for (int j = 0; j < kd; ++j) {
        const int c = nbr_d[i * kd + j];
        if (c < 0) continue;
#pragma unroll
        for (int m = 0; m < RPT; ++m) acc[m] += (S)w_d[(l0 + m) * sd_stride + i * kd + j] * prev[m * NP + c];
      }
    }
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
      const int t = t0 + l0 + m;
      q[m] = (active && t >= 1 && t < T) ? v[m] - acc[m] : (S)0;
    }
  }
  // (L_d^T + Q1) v with v published in qbuf (ADMM.py:196-223): the "father" sum over the in-list at step t+1
  __device__ __forceinline__ void ldrt(const S (&v)[RPT], S (&out)[RPT]) const {
    S f[RPT];
#pragma unroll
    for (int m = 0; m < RPT; ++m) f[m] = 0;
    if (active) {
      const S* next = qbuf + (l0 + 1) * NP;
      for (int e = cptr[i]; e < cptr[i + 1]; ++e) {
        const int src = csrc[e];
#pragma unroll
        for (int m = 0; m < RPT; ++m) {
          // step t reads d_w[t]: one slice after the one L_d uses at the same step (CMP: slices t0 - 1 .. t0 + TL - 1 are staged)
          const float w = CMP ? w_d[(l0 + m + 1) * sd_stride + slot[e]] : cw[(l0 + m) * sc_stride + e];
          f[m] += (S)w * next[m * NP + src];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
      const int t = t0 + l0 + m;
      S o = 0;
      if (active && t < T) o = (t == T - 1) ? v[m] : ((t == 0 && !q1) ? -f[m] : v[m] - f[m]);   // rows beyond T - 1 gather zeros
      out[m] = o;
    }
  }
  // L_u v with v published in pbuf (ADMM.py:138-148)
  __device__ __forceinline__ void lu(const S (&v)[RPT], S (&out)[RPT]) const {
    S acc[RPT];
#pragma unroll
    for (int m = 0; m < RPT; ++m) acc[m] = 0;
    if (active) {
      const S* row = pbuf + (l0 + 1) * NP;
      for (int j = 0; j < ku; ++j) {
        const int c = nbr_u[i * ku + j];
        if (c < 0) continue;
#pragma unroll
        for (int m = 0; m < RPT; ++m) acc[m] += (S)w_u[(l0 + m) * su_stride + i * ku + j] * row[m * NP + c];
      }
    }
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
      const int t = t0 + l0 + m;
      out[m] = (active && t < T) ? v[m] - acc[m] : (S)0;
    }
  }
  // out = A v for system SYS (ADMM.py:371-399), the reference's evaluation order.
  // DOT: also dot = <v, A v>.  For the x / z_d systems with an exact transpose (kNN scatter mode, line graph)
  // A = D + c L_d^T L_d, so <v, A v> = sum D v^2 + c ||L_d v||^2: every term is known once q = L_d v is, and the partial
  // sums ride on the cluster barrier that publishes q - one barrier fewer per CG iteration (the same value up to
  // rounding; the resident float32 kernel does the same).
  // mk != nullptr: H = the caller's elementwise mask (values of this thread's steps) instead of "rows t < t_in" - the
  // reference passes the mask to the FIRST residual of the x solve only (quirk Q4, ADMM.py:344-349)
  template <int SYS, bool DOT>
  __device__ __forceinline__ void apply(const S (&v)[RPT], S (&out)[RPT], S a, S c, S& dot, S (&q)[RPT], const S* mk = nullptr) {
    publish_p(v);
    if (SYS == MGA_SYS_ZU) {
      lu(v, out);
      S loc = 0;
#pragma unroll
      for (int m = 0; m < RPT; ++m) {
        out[m] = c * out[m] + a * v[m];                                                 // ADMM.py:390
        loc += v[m] * out[m];
      }
      if (DOT) dot = csum(loc);
      return;
    }
    ldr(v, q);
    const bool fused = DOT && qdot;
    if (fused) csum_post(quad<SYS>(v, q, a, c));      // partial -> every CTA's slot; completed by the barrier of publish_q
    publish_q(q);
    if (fused) dot = csum_read();
    ldrt(q, out);
    S loc = 0;
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
      if (SYS == MGA_SYS_X) {
        const S hx = mk ? v[m] * mk[m] : ((t0 + l0 + m < t_in) ? v[m] : (S)0);          // H^T H (ADMM.py:372-376)
        out[m] = (hx + a * v[m]) + c * out[m];                                          // ADMM.py:379
      } else {
        out[m] = c * out[m] + a * v[m];                                                 // ADMM.py:394
      }
      loc += v[m] * out[m];
    }
    if (DOT && !fused) dot = csum(loc);
  }
  // this thread's part of <v, A v> = sum D v^2 + c ||L_d v||^2 with q = L_d v
  template <int SYS>
  __device__ __forceinline__ S quad(const S (&v)[RPT], const S (&q)[RPT], S a, S c) const {
    S loc = 0;
#pragma unroll
    for (int m = 0; m < RPT; ++m) {
      const S hx = (SYS == MGA_SYS_X && t0 + l0 + m < t_in) ? v[m] * v[m] : (S)0;
      loc += (hx + a * (v[m] * v[m])) + c * (q[m] * q[m]);
    }
    return loc;
  }

  // CG_solver (ADMM.py:329-368): r holds the right-hand side on entry, x the warm start.  Returns the iteration count
  // (tol > 0 and reached) or -1.  alpha_out / beta_out: this window's column of the (max_cg, B) arrays or NULL.
  //
  // Two cluster barriers per iteration instead of three.  The reference's iteration needs p' = r' + beta p before it can
  // gather A p', and beta needs <r', r'> - a barrier of its own.  A is linear: A p' = A r' + beta A p (and L_d p' =
  // L_d r' + beta L_d p), so the kernel publishes r' TOGETHER with the partial sums of <r', r'> (barrier A), gathers on
  // r', and rebuilds A p' / L_d p' from the previous ones it still holds; <p', A p'> rides on the barrier that publishes
  // L_d r' (barrier B; z_u: its own reduction).  Same iterates up to rounding; alpha, beta, the stop test on
  // sqrt(<r', r'>) (ADMM.py:360) and the iteration counts are those of the reference's recurrence.
  template <int SYS>
  __device__ __forceinline__ int cg_solve(S (&x)[RPT], S (&r)[RPT], S a, S c, int max_cg, double tol, S* alpha_out, S* beta_out,
                                          int64_t B, const S* mask0 = nullptr) {
    S p[RPT], ap[RPT], qp[RPT], t[RPT];
    S pap = 0;
    apply<SYS, false>(x, ap, a, c, pap, qp, mask0);
    S loc = 0;
#pragma unroll
    for (int l = 0; l < RPT; ++l) {
      r[l] = r[l] - ap[l];
      p[l] = r[l];
      loc += r[l] * r[l];
    }
    S rr = csum(loc);
    if (max_cg <= 0) return -1;
    apply<SYS, true>(p, ap, a, c, pap, qp);                    // A p_0, L_d p_0, <p_0, A p_0>
    for (int k = 0; k < max_cg; ++k) {
      const S alpha = rr / pap;
      loc = 0;
#pragma unroll
      for (int l = 0; l < RPT; ++l) {
        x[l] = x[l] + alpha * p[l];
        r[l] = r[l] - alpha * ap[l];
        loc += r[l] * r[l];
      }
      const bool more = k + 1 < max_cg;
      if (more) put_p(r);                                      // barrier A: r' and the partial sums of <r', r'>
      csum_post(loc);
      cl.sync();
      const S rrn = csum_read();
      const S beta = rrn / rr;
      rr = rrn;
      if (alpha_out && rank == 0 && threadIdx.x == 0) {
        alpha_out[(size_t)k * B] = alpha;
        beta_out[(size_t)k * B] = beta;
      }
      if (tol > 0 && sqrt(rr) < (S)tol) return k + 1;          // ADMM.py:360 (B = 1: the max over the batch is this window)
      if (!more) break;
#pragma unroll
      for (int l = 0; l < RPT; ++l) p[l] = r[l] + beta * p[l];
      if (SYS == MGA_SYS_ZU) {
        lu(r, t);
        loc = 0;
#pragma unroll
        for (int l = 0; l < RPT; ++l) {
          ap[l] = (c * t[l] + a * r[l]) + beta * ap[l];
          loc += p[l] * ap[l];
        }
        pap = csum(loc);                                       // barrier B
        continue;
      }
      S qr[RPT];
      ldr(r, qr);
#pragma unroll
      for (int l = 0; l < RPT; ++l) qp[l] = qr[l] + beta * qp[l];
      if (qdot) csum_post(quad<SYS>(p, qp, a, c));
      publish_q(qr);                                           // barrier B: L_d r' and the partial sums of <p', A p'>
      if (qdot) pap = csum_read();
      ldrt(qr, t);
      loc = 0;
#pragma unroll
      for (int l = 0; l < RPT; ++l) {
        S ar;
        if (SYS == MGA_SYS_X) ar = (((t0 + l0 + l < t_in) ? r[l] : (S)0) + a * r[l]) + c * t[l];
        else ar = c * t[l] + a * r[l];
        ap[l] = ar + beta * ap[l];
        loc += p[l] * ap[l];
      }
      if (!qdot) pap = csum(loc);                              // use_kNN=False: the "transpose" is a gather with the forward table
    }
    return -1;
  }
};

// MAXT: launch bound (384: PEMS-sized graphs keep their registers - no spills in float64; 1024: up to 1024 nodes)
// MINB = 2 (batches): two CTAs per SM, i.e. two clusters share a group of SMs - a cluster alone leaves its SMs waiting on barriers
template <typename S, int RPT, int MAXT, int MINB = 1>
__global__ void __launch_bounds__(MAXT, MINB) k_admm_cluster(const ClArgs<S> a) {
  extern __shared__ __align__(16) unsigned char smem_cl[];
  cg::cluster_group cluster = cg::this_cluster();
  const int N = a.N, NP = a.NP, T = a.T, t_in = a.t_in, CL = a.CL, TL = a.TL;
  const int rank = (int)cluster.block_rank();
  const int64_t b = blockIdx.x / CL;
  // ---- carve shared memory
  S* state = reinterpret_cast<S*>(smem_cl);                        // [CS_COUNT][TL][NP]
  S* pbuf = state + (size_t)CS_COUNT * TL * NP;
  S* qbuf = pbuf + (size_t)(TL + 1) * NP;
  S* red = qbuf + (size_t)(TL + 1) * NP;
  S* slots = red + 32;
  // weight slices: one per local time step when the caller's tables vary in time, else one
  constexpr bool CMP = MINB == 2;
  using I = typename ClCtx<S, RPT, CMP>::I;
  const int nsd = a.d_wT > 1 ? (CMP ? TL + 1 : TL) : 1, nsu = a.u_wT > 1 ? TL : 1;
  float* w_d = reinterpret_cast<float*>(slots + 2 * kClMaxCL);      // [nsd][N * kd]: slice l = weights of L_d at step t0 + l (d_w[t - 1])
  float* w_u = w_d + (size_t)nsd * N * a.kd;                         // [nsu][N * ku]: u_w[t]
  float* cw = w_u + (size_t)nsu * N * a.ku;                          // [nsd][nnz]: in-list weights of L_d^T at step t (d_w[t][slot]); not CMP
  int* cptr = reinterpret_cast<int*>(cw + (CMP ? 0 : (size_t)nsd * a.nnz));
  I* nbr_d = reinterpret_cast<I*>(cptr + N + 1);
  I* nbr_u = nbr_d + (size_t)N * a.kd;
  I* csrc = nbr_u + (size_t)N * a.ku;
  unsigned short* slot = reinterpret_cast<unsigned short*>(csrc + a.nnz);     // CMP only
  unsigned char* tail = reinterpret_cast<unsigned char*>(CMP ? (void*)(slot + a.nnz) : (void*)(csrc + a.nnz));
  S* dredS = reinterpret_cast<S*>((reinterpret_cast<size_t>(tail) + 15) & ~(size_t)15);      // 12 x 32 partials of the diagnostics
  for (int k = threadIdx.x; k < N * a.kd; k += blockDim.x) nbr_d[k] = (I)a.nbr_d[k];
  for (int k = threadIdx.x; k < N * a.ku; k += blockDim.x) nbr_u[k] = (I)a.nbr_u[k];
  for (int k = threadIdx.x; k < a.nnz; k += blockDim.x) csrc[k] = (I)a.csr_src[k];
  if (CMP)
    for (int k = threadIdx.x; k < a.nnz; k += blockDim.x) slot[k] = (unsigned short)a.csr_slot[k];
  for (int l = 0; l < nsd; ++l) {
    const int t = rank * TL + l;
    const int sd = a.d_wT > 1 ? min(max(t - 1, 0), a.d_wT - 1) : 0;        // L_d at step t uses d_w[t - 1] (ADMM.py:171)
    const int st = a.d_wT > 1 ? min(t, a.d_wT - 1) : 0;                    // L_d^T at step t uses d_w[t] (ADMM.py:200-208)
    for (int k = threadIdx.x; k < N * a.kd; k += blockDim.x) w_d[(size_t)l * N * a.kd + k] = a.d_w[(size_t)sd * N * a.kd + k];
    if (!CMP)
      for (int k = threadIdx.x; k < a.nnz; k += blockDim.x)
        cw[(size_t)l * a.nnz + k] = a.d_wT > 1 ? a.d_w[(size_t)st * N * a.kd + a.csr_slot[k]] : a.csr_w[k];
  }
  for (int l = 0; l < nsu; ++l) {
    const int su = a.u_wT > 1 ? min(rank * TL + l, a.u_wT - 1) : 0;
    for (int k = threadIdx.x; k < N * a.ku; k += blockDim.x) w_u[(size_t)l * N * a.ku + k] = a.u_w[(size_t)su * N * a.ku + k];
  }
  for (int k = threadIdx.x; k <= N; k += blockDim.x) cptr[k] = a.csr_ptr[k];
  for (int k = threadIdx.x; k < 2 * (TL + 1) * NP; k += blockDim.x) pbuf[k] = (S)0;     // incl. the outermost halo rows
  for (int k = threadIdx.x; k < 2 * kClMaxCL; k += blockDim.x) slots[k] = (S)0;

  ClCtx<S, RPT, CMP> c;
  const int lr = threadIdx.x / NP;                                   // row group of this thread: steps lr * RPT .. + RPT - 1 of the slab
  c.N = N; c.NP = NP; c.T = T; c.t_in = t_in; c.t0 = rank * TL; c.i = threadIdx.x - lr * NP; c.rank = rank; c.CL = CL;
  c.TL = TL; c.l0 = lr * RPT;
  c.kd = a.kd; c.ku = a.ku; c.q1 = a.q1; c.qdot = a.transpose_exact != 0;
  c.sd_stride = a.d_wT > 1 ? N * a.kd : 0; c.su_stride = a.u_wT > 1 ? N * a.ku : 0; c.sc_stride = a.d_wT > 1 ? a.nnz : 0;
  c.active = c.i < N;
  c.nbr_d = nbr_d; c.w_d = w_d; c.nbr_u = nbr_u; c.w_u = w_u; c.cptr = cptr; c.csrc = csrc; c.cw = cw; c.slot = slot;
  c.pbuf = pbuf; c.qbuf = qbuf; c.red = red; c.slots = slots; c.parity = 0;
  cluster.sync();                                                    // tables and zeroed buffers of every CTA are in place
  const int i = c.i, t0 = c.t0 + c.l0;                               // t0: the thread's first step
  const bool active = c.active;
  auto st = [&](int v, int m) -> S& { return state[((size_t)v * TL + c.l0 + m) * NP + i]; };
  // the tables come in reverse-Cuthill-McKee order: the 32 nodes of a warp gather from a few neighbouring rows (fewer
  // shared-memory bank conflicts than with the caller's numbering); global memory keeps the caller's order
  const int orig = active ? (a.perm ? a.perm[i] : i) : 0;
  const bool maskm = a.mask != nullptr;
  const int y_rows = maskm ? T : t_in;
  const S* yw = a.y + (size_t)b * y_rows * N + orig;
  const S* mw = maskm ? a.mask + (size_t)b * T * N + orig : nullptr;

  // ---- initial_guess (ADMM.py:766-781) and initial state (ADMM.py:537-544)
  {
    S w, cc;
    if (maskm) {                                                     // initial_interpolation (ADMM.py:783-811)
      S cnt = 0, st_ = 0, sy = 0, sty = 0, st2 = 0;
      for (int t = 0; t < T; ++t) {
        const S m = active ? mw[(size_t)t * N] : (S)1, v = active ? yw[(size_t)t * N] : (S)0, tt = (S)(float)t;
        cnt += m; st_ += tt * m; sy += v * m; sty += tt * v * m; st2 += tt * tt * m;
      }
      const S tm = st_ / cnt, ym = sy / cnt, tym = sty / cnt, t2m = st2 / cnt;
      w = (tym - tm * ym) / (t2m - tm * tm);
      cc = ym - w * tm;
    } else {
      S sy = 0, sty = 0;
      if (active)
        for (int t = 0; t < t_in; ++t) {
          const S v = yw[(size_t)t * N];
          sy += v;
          sty += (S)(float)t * v;
        }
      const S my = sy / (S)t_in, mty = sty / (S)t_in;
      w = (mty - (S)a.t_mean * my) / (S)a.t_var;
      cc = my - w * (S)a.t_mean;
    }
    S x[RPT], q[RPT];
#pragma unroll
    for (int l = 0; l < RPT; ++l) {
      const int t = t0 + l;
      S v = 0;
      if (active && t < T) {
        if (maskm) v = (w * (S)(float)t + cc) * ((S)1 - mw[(size_t)t * N]) + yw[(size_t)t * N];
        else v = t < t_in ? yw[(size_t)t * N] : w * (S)(float)t + cc;
      }
      x[l] = v;
      const S tenth = (active && t < T) ? (S)0.1 : (S)0;
      st(CS_X, l) = v; st(CS_ZU, l) = v; st(CS_ZD, l) = v;
      st(CS_GU, l) = tenth; st(CS_GD, l) = tenth; st(CS_GAM, l) = tenth;
    }
    c.publish_p(x);
    c.ldr(x, q);                                                     // phi = L_d x (ADMM.py:541)
#pragma unroll
    for (int l = 0; l < RPT; ++l) st(CS_PHI, l) = q[l];
  }

  int outer_done = 0;
  for (int it = 0; it < a.n_outer; ++it) {
    const size_t sys_stride = (size_t)a.max_cg * a.B;
    S* al = a.alpha ? a.alpha + ((size_t)it * 3) * sys_stride + b : nullptr;
    S* be = a.beta ? a.beta + ((size_t)it * 3) * sys_stride + b : nullptr;
    int iters[3] = {-1, -1, -1};
    S d[MGA_DIAG_COLS];
#pragma unroll
    for (int k = 0; k < MGA_DIAG_COLS; ++k) d[k] = 0;
    S r[RPT], x[RPT];
    // ---- RHS_x (ADMM.py:552-559): Ldr_T(gamma + rho phi)/2 + (rho_u zu + rho_d zd)/2 - (gu + gd)/2 + H^T y
    {
      S v[RPT], lt[RPT];
#pragma unroll
      for (int l = 0; l < RPT; ++l) v[l] = st(CS_GAM, l) + a.rho * st(CS_PHI, l);
      c.publish_q(v);
      c.ldrt(v, lt);
#pragma unroll
      for (int l = 0; l < RPT; ++l) {
        const int t = t0 + l;
        S o = 0;
        if (active && t < T) {
          const S hty = t < y_rows ? yw[(size_t)t * N] : (S)0;
          o = lt[l] / (S)2 + (a.rho_u * st(CS_ZU, l) + a.rho_d * st(CS_ZD, l)) / (S)2 - (st(CS_GU, l) + st(CS_GD, l)) / (S)2 + hty;
        }
        r[l] = o;
        x[l] = st(CS_X, l);
      }
    }
    // ---- x solve (ADMM.py:571), warm start x_old
    {
      S m0[RPT];
#pragma unroll
      for (int l = 0; l < RPT; ++l) m0[l] = (maskm && active && t0 + l < T) ? mw[(size_t)(t0 + l) * N] : (S)0;
      iters[0] = c.template cg_solve<MGA_SYS_X>(x, r, a.ax, a.cx, a.max_cg, a.cg_tol, al, be, a.B, maskm ? m0 : nullptr);
    }
#pragma unroll
    for (int l = 0; l < RPT; ++l) {
      const int t = t0 + l;
      if (active && t < T) {
        const S dx = x[l] - st(CS_X, l);
        d[MGA_DIAG_DX2] += dx * dx;
        d[MGA_DIAG_NONFINITE] += (S)(!isfinite(x[l]));
        if (a.want_diag && a.dx_sum) atomicAdd(a.dx_sum + ((size_t)it * T + t) * N + orig, (double)dx);
      }
      st(CS_X, l) = x[l];
    }
    // ---- z_u solve (ADMM.py:579-580)
    {
      S z[RPT];
#pragma unroll
      for (int l = 0; l < RPT; ++l) { r[l] = st(CS_GU, l) / (S)2 + a.azu * x[l]; z[l] = st(CS_ZU, l); }
      iters[1] = c.template cg_solve<MGA_SYS_ZU>(z, r, a.azu, a.czu, a.max_cg, a.cg_tol, al ? al + sys_stride : nullptr,
                                                 be ? be + sys_stride : nullptr, a.B);
#pragma unroll
      for (int l = 0; l < RPT; ++l) {
        if (active && t0 + l < T) {
          const S d0 = x[l] - z[l], d1 = z[l] - st(CS_ZU, l);
          d[MGA_DIAG_X_ZU2] += d0 * d0;
          d[MGA_DIAG_DZU2] += d1 * d1;
          d[MGA_DIAG_NONFINITE] += (S)(!isfinite(z[l]));
        }
        st(CS_ZU, l) = z[l];
      }
    }
    // ---- z_d solve (ADMM.py:587-588)
    {
      S z[RPT];
#pragma unroll
      for (int l = 0; l < RPT; ++l) { r[l] = st(CS_GD, l) / (S)2 + a.azd * x[l]; z[l] = st(CS_ZD, l); }
      iters[2] = c.template cg_solve<MGA_SYS_ZD>(z, r, a.azd, a.czd, a.max_cg, a.cg_tol, al ? al + 2 * sys_stride : nullptr,
                                                 be ? be + 2 * sys_stride : nullptr, a.B);
#pragma unroll
      for (int l = 0; l < RPT; ++l) {
        if (active && t0 + l < T) {
          const S d0 = x[l] - z[l], d1 = z[l] - st(CS_ZD, l);
          d[MGA_DIAG_X_ZD2] += d0 * d0;
          d[MGA_DIAG_DZD2] += d1 * d1;
          d[MGA_DIAG_NONFINITE] += (S)(!isfinite(z[l]));
          // dual ascents (ADMM.py:595-597): gamma_u uses the z_u stored above
          st(CS_GU, l) = st(CS_GU, l) + a.rho_u * (x[l] - st(CS_ZU, l));
          st(CS_GD, l) = st(CS_GD, l) + a.rho_d * d0;
        }
        st(CS_ZD, l) = z[l];
      }
    }
    // ---- phi prox + gamma ascent (ADMM.py:600-605) and the remaining diagnostics (ADMM.py:612-637)
    {
      S ldx[RPT], lux[RPT];
      c.publish_p(x);
      c.ldr(x, ldx);
      if (a.want_diag) c.lu(x, lux);
#pragma unroll
      for (int l = 0; l < RPT; ++l) {
        const int t = t0 + l;
        if (active && t < T) {
          const S gam = st(CS_GAM, l), pho = st(CS_PHI, l), q = ldx[l];
          const S ph = cl_soft<S>(q - gam / a.rho, a.thr);
          const S gn = gam + a.rho * (ph - q);
          d[MGA_DIAG_NONFINITE] += (S)(!isfinite(ph) || !isfinite(gn));
          const S e = ph - q, f = ph - pho;
          d[MGA_DIAG_PHI_LDX2] += e * e;
          d[MGA_DIAG_DPHI2] += f * f;
          d[MGA_DIAG_DGTV] += fabs(q);
          d[MGA_DIAG_DGLR] += q * q;
          if (a.want_diag) {
            d[MGA_DIAG_GLR] += x[l] * lux[l];
            if (maskm) {                                           // ||x * mask - y|| (ADMM.py:620-621)
              const S h = x[l] * mw[(size_t)t * N] - yw[(size_t)t * N];
              d[MGA_DIAG_RECOVER2] += h * h;
            } else if (t < t_in) {
              const S h = x[l] - yw[(size_t)t * N];
              d[MGA_DIAG_RECOVER2] += h * h;
            }
          }
          st(CS_PHI, l) = ph;
          st(CS_GAM, l) = gn;
        }
      }
    }
    // ---- diagnostics out (one double atomic per CTA and column) and the outer stop test (ADMM.py:645)
    {
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
      for (int k = 0; k < MGA_DIAG_COLS; ++k) {
        const S v = warp_sum<S>(d[k]);
        if (lane == 0) dredS[k * 32 + w] = v;
      }
      __syncthreads();
      if (threadIdx.x < MGA_DIAG_COLS && a.diag && (a.want_diag || threadIdx.x == MGA_DIAG_NONFINITE)) {
        S t = 0;
        for (int k = 0; k < nw; ++k) t += dredS[threadIdx.x * 32 + k];
        if (t != (S)0) atomicAdd(a.diag + (size_t)it * MGA_DIAG_COLS + threadIdx.x, (double)t);
      }
      __syncthreads();
    }
    if (a.cg_iters && rank == 0 && threadIdx.x == 0) {
      a.cg_iters[it * 3 + 0] = iters[0]; a.cg_iters[it * 3 + 1] = iters[1]; a.cg_iters[it * 3 + 2] = iters[2];
    }
    outer_done = it + 1;
    if (a.admm_tol > 0) {
      // whole-window norms, as the reference takes them in the signal dtype (B = 1)
      S nrm[6] = {d[MGA_DIAG_X_ZU2], d[MGA_DIAG_DZU2], d[MGA_DIAG_PHI_LDX2], d[MGA_DIAG_DPHI2], d[MGA_DIAG_X_ZD2], d[MGA_DIAG_DZD2]};
      c.template csum_vec<6>(nrm, dredS, dredS + MGA_DIAG_COLS * 32);
      const S pri = fmax(fmax(sqrt(nrm[0]), sqrt(nrm[2])), sqrt(nrm[4])), dua = fmax(fmax(sqrt(nrm[1]), sqrt(nrm[3])), sqrt(nrm[5]));
      if ((double)pri < a.admm_tol && (double)dua < a.admm_tol) break;
    }
  }
  if (a.outer_done && rank == 0 && threadIdx.x == 0 && b == 0) *a.outer_done = outer_done;
  // ---- results
#pragma unroll
  for (int l = 0; l < RPT; ++l) {
    const int t = t0 + l;
    if (active && t < T) {
      const size_t at = ((size_t)b * T + t) * N + orig;
      a.x_out[at] = st(CS_X, l);
      if (a.out[0]) a.out[0][at] = st(CS_ZU, l);
      if (a.out[1]) a.out[1][at] = st(CS_ZD, l);
      if (a.out[2]) a.out[2][at] = st(CS_PHI, l);
      if (a.out[3]) a.out[3][at] = st(CS_GAM, l);
      if (a.out[4]) a.out[4][at] = st(CS_GU, l);
      if (a.out[5]) a.out[5][at] = st(CS_GD, l);
    }
  }
  cluster.sync();      // no CTA leaves while a neighbour may still push into its shared memory
}

// ---- host side ------------------------------------------------------------------------------------------------
struct ClGeom {
  int CL, TL, TR, RPT, NP;      // cluster size, steps per CTA, row groups per CTA (threads = NP * TR), steps per thread
  size_t smem;
  size_t smem2;                 // with compact tables (the two-CTAs-per-SM instantiation)
};

template <typename S>
static bool cl_geometry(const mga_plan* p, ClGeom* out, int maxcl = kClPortable) {
  const GraphDev& g = p->g;
  if (g.temporal == MGA_TEMPORAL_BAND || g.N > 1024 || g.T > maxcl * kClMaxTL) return false;
  ClGeom q;
  q.NP = ((g.N + 31) / 32) * 32;
  q.CL = std::min(g.T, maxcl);
  q.TL = (g.T + q.CL - 1) / q.CL;
  q.TR = 1;                                       // one thread per node and slab: a table entry is read once for all its steps
  if (const char* e = std::getenv("MGA_CLUSTER_TR")) q.TR = std::max(1, std::min(std::min(q.TL, 1024 / q.NP), std::atoi(e)));
  q.RPT = (q.TL + q.TR - 1) / q.TR;
  q.TL = q.TR * q.RPT;
  q.CL = (g.T + q.TL - 1) / q.TL;                 // no empty CTAs (T = 9 -> TL = 2 -> 5 CTAs)
  const size_t vec = ((size_t)CS_COUNT * q.TL + 2 * (q.TL + 1)) * q.NP * sizeof(S) + (32 + 2 * kClMaxCL) * sizeof(S) +
                     ((size_t)MGA_DIAG_COLS * 32 + 6 * kClMaxCL) * sizeof(S) + 32 + (size_t)(g.N + 1) * 4;
  const size_t nd = (size_t)g.N * g.kd, nu = (size_t)g.N * g.ku, sl_u = g.u_wT > 1 ? q.TL : 1;
  q.smem = vec + (nd + nu + g.nnz) * 4 + ((size_t)(g.d_wT > 1 ? q.TL : 1) * (nd + g.nnz) + sl_u * nu) * 4;
  q.smem2 = vec + (nd + nu + 2 * (size_t)g.nnz) * 2 + 8 + ((size_t)(g.d_wT > 1 ? q.TL + 1 : 1) * nd + sl_u * nu) * 4;
  if (q.smem > (size_t)p->max_smem_optin) return false;
  *out = q;
  return true;
}

bool cluster_eligible(const mga_plan* p, int dtype) {
  ClGeom q;
  return dtype == MGA_F64 ? cl_geometry<double>(p, &q) : cl_geometry<float>(p, &q);
}

template <typename S, int RPT>
static int cl_launch(mga_plan* p, const ClArgs<S>& a, const ClGeom& q, cudaStream_t st) {
  // batches that oversubscribe the GPU with one cluster per SM group: two CTAs per SM when shared memory and registers allow
  const bool two = a.B * q.CL > p->sm_count && q.NP * q.TR <= 384 && 2 * (q.smem2 + 1024) <= (size_t)p->max_smem_sm &&
                   (size_t)a.N * a.kd < 65536 && !std::getenv("MGA_CLUSTER_ONE");
  const size_t smem = two ? q.smem2 : q.smem;
  const int nthr = q.NP * q.TR;
  auto kern = nthr > 384 ? k_admm_cluster<S, RPT, 1024, 1>
              : !two     ? k_admm_cluster<S, RPT, 384, 1>
              : nthr <= 320 ? k_admm_cluster<S, RPT, 320, 2>      // PEMS04-sized graphs: 96 instead of 80 registers
                            : k_admm_cluster<S, RPT, 384, 2>;
  MGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (q.CL > kClPortable) MGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(a.B * q.CL));
  cfg.blockDim = dim3(q.NP * q.TR);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = q.CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (q.CL > kClPortable) {            // a GPC must have room for the whole cluster; MGA_ERR_UNSUPPORTED sends the caller back to 8
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc < 1) { cudaGetLastError(); return MGA_ERR_UNSUPPORTED; }
  }
  MGA_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  MGA_LAUNCH_CHECK("k_admm_cluster");
  return MGA_OK;
}

template <typename S>
static int cl_admm(mga_plan* p, const mga_params* m, const void* y, const void* mask, void* x_out, int64_t B, int n_outer, int max_cg, double cg_tol,
                   double admm_tol, double t_mean, double t_var, int diag_flags, const mga_admm_outputs* outs, cudaStream_t st,
                   int force_cl = 0) {
  const GraphDev& g = p->g;
  // Few windows: the largest cluster a GPC takes (T = 24: 12 CTAs of 2 steps instead of 8 of 3 - the slab's gathers are bound by
  // ONE SM's shared-memory pipe); batches: portable clusters, two of which fit a GPC.  MGA_CLUSTER_CL overrides.
  int maxcl = B <= 8 ? kClMaxCL : kClPortable;
  if (const char* e = std::getenv("MGA_CLUSTER_CL")) maxcl = std::max(1, std::min(kClMaxCL, std::atoi(e)));
  if (force_cl > 0) maxcl = force_cl;
  ClGeom q;
  if (!cl_geometry<S>(p, &q, maxcl) && !cl_geometry<S>(p, &q)) { set_error("cluster mode: the window does not fit"); return MGA_ERR_UNSUPPORTED; }
  const bool tol_mode = cg_tol > 0 || admm_tol > 0;
  if (tol_mode && B != 1) { set_error("cluster mode: the stop tests are per window, tolerance mode takes B = 1"); return MGA_ERR_UNSUPPORTED; }
  ClArgs<S> a{};
  a.N = g.N; a.NP = q.NP; a.T = g.T; a.t_in = g.t_in; a.CL = q.CL; a.TL = q.TL; a.transpose_exact = p->ldrt_gather ? 0 : 1; a.u_wT = g.u_wT; a.d_wT = g.d_wT; a.csr_slot = g.csr_slot; a.kd = g.kd; a.ku = g.ku; a.q1 = g.q1; a.nnz = g.nnz;
  a.n_outer = n_outer; a.max_cg = max_cg; a.want_diag = (diag_flags & 1) ? 1 : 0;
  a.B = B;
  a.nbr_d = g.nbr_d; a.d_w = g.d_w; a.nbr_u = g.nbr_u; a.u_w = g.u_w; a.csr_ptr = g.csr_ptr; a.csr_src = g.csr_src; a.csr_w = g.csr_w;
  if (p->has_s2 && !std::getenv("MGA_CLUSTER_NATURAL")) {       // the streaming path's reordered tables (same entries, same order within a row)
    const Graph2& g2 = p->g2;
    a.nbr_d = g2.nbr_d; a.d_w = g2.w_d; a.nbr_u = g2.nbr_u; a.u_w = g2.w_u;
    a.csr_ptr = g2.in_ptr; a.csr_src = g2.in_src; a.csr_w = g2.in_w; a.csr_slot = g2.in_slot; a.perm = g2.perm;
  }
  a.y = static_cast<const S*>(y); a.x_out = static_cast<S*>(x_out);
  a.mask = static_cast<const S*>(mask);
  a.out[0] = static_cast<S*>(outs->zu); a.out[1] = static_cast<S*>(outs->zd); a.out[2] = static_cast<S*>(outs->phi);
  a.out[3] = static_cast<S*>(outs->gamma); a.out[4] = static_cast<S*>(outs->gamma_u); a.out[5] = static_cast<S*>(outs->gamma_d);
  a.alpha = static_cast<S*>(outs->alpha); a.beta = static_cast<S*>(outs->beta);
  if (a.alpha && !a.beta) a.alpha = nullptr;
  a.cg_tol = cg_tol; a.admm_tol = admm_tol;
  a.rho = (S)m->rho; a.rho_u = (S)m->rho_u; a.rho_d = (S)m->rho_d; a.thr = (S)(m->mu_d1 / m->rho);
  a.ax = (S)((m->rho_u + m->rho_d) / 2); a.cx = (S)(m->rho / 2);
  a.azu = (S)(m->rho_u / 2); a.czu = (S)m->mu_u;
  a.azd = (S)(m->rho_d / 2); a.czd = (S)m->mu_d2;
  a.t_mean = (float)t_mean; a.t_var = (float)t_var;
  // diagnostics: the non-finite column is needed even with diagnostics off -> a scratch row block in the workspace
  const size_t diag_bytes = (size_t)std::max(n_outer, 1) * MGA_DIAG_COLS * sizeof(double);
  double* diag = a.want_diag ? outs->diag : nullptr;
  if (!diag) {
    int rc = ensure_workspace(p, p->ws, diag_bytes);
    if (rc) return rc;
    diag = static_cast<double*>(p->ws.base);
  }
  a.diag = diag;
  a.dx_sum = a.want_diag ? outs->dx_sum : nullptr;
  if (!(a.want_diag && (diag_flags & 2))) {
    MGA_CUDA(cudaMemsetAsync(diag, 0, diag_bytes, st));
    if (a.dx_sum) MGA_CUDA(cudaMemsetAsync(a.dx_sum, 0, (size_t)n_outer * g.T * g.N * sizeof(double), st));
  }
  // iteration counts of tolerance mode come back through the plan's pinned block (device-visible host memory)
  int* h_iters = reinterpret_cast<int*>(static_cast<char*>(p->pinned) + 1024);
  if (tol_mode) {
    if ((size_t)n_outer * 3 * sizeof(int) + 2048 > p->pinned_bytes) { set_error("cluster mode: too many outer iterations"); return MGA_ERR_UNSUPPORTED; }
    for (int k = 0; k < n_outer * 3; ++k) h_iters[64 + k] = -1;
    h_iters[0] = 0;
    a.cg_iters = h_iters + 64;
    a.outer_done = h_iters;
  }
  int rc;
  switch (q.RPT) {
    case 1: rc = cl_launch<S, 1>(p, a, q, st); break;
    case 2: rc = cl_launch<S, 2>(p, a, q, st); break;
    case 3: rc = cl_launch<S, 3>(p, a, q, st); break;
    default: rc = cl_launch<S, 4>(p, a, q, st); break;
  }
  if (rc == MGA_ERR_UNSUPPORTED && q.CL > kClPortable)      // no GPC takes this cluster: portable size
    return cl_admm<S>(p, m, y, mask, x_out, B, n_outer, max_cg, cg_tol, admm_tol, t_mean, t_var, diag_flags, outs, st, kClPortable);
  if (rc) return rc;
  if (tol_mode) {
    MGA_CUDA(cudaStreamSynchronize(st));
    const int done = h_iters[0];
    if (outs->cg_iters) for (int k = 0; k < n_outer * 3; ++k) outs->cg_iters[k] = k < done * 3 ? h_iters[64 + k] : -1;
    if (outs->outer_done) *outs->outer_done = done;
  } else {
    if (outs->cg_iters) for (int k = 0; k < n_outer * 3; ++k) outs->cg_iters[k] = -1;
    if (outs->outer_done) *outs->outer_done = n_outer;
  }
  return MGA_OK;
}

int cluster_admm(mga_plan* p, const mga_params* m, const void* y, const void* mask, void* x_out, int64_t B, int dtype, int n_outer, int max_cg,
                 double cg_tol, double admm_tol, double t_mean, double t_var, int diag_flags, const mga_admm_outputs* outs,
                 cudaStream_t st) {
  if (dtype == MGA_F64)
    return cl_admm<double>(p, m, y, mask, x_out, B, n_outer, max_cg, cg_tol, admm_tol, t_mean, t_var, diag_flags, outs, st);
  return cl_admm<float>(p, m, y, mask, x_out, B, n_outer, max_cg, cg_tol, admm_tol, t_mean, t_var, diag_flags, outs, st);
}

}  // namespace mga
