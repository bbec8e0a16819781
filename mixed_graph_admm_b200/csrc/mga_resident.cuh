// Resident mode: the whole of combined_loop (ADMM.py:528-648) for one window inside one CTA.
//
// Mapping: thread i owns node i for ALL time steps.  The four CG vectors (x, r, p, Ap) of the
// thread's T lattice points live in registers; only the two vectors other threads gather from
// (p and q = L_d p) are staged in shared memory, row-major (t, node) with one extra zero slot
// per row that absorbs the "-1 = no neighbour" entries (quirk Q6).  The thread's own rows of the
// ELL tables (neighbour offsets + weights) are time-invariant and sit in registers; the in-list
// (transposed CSR) that turns the reference's scatter_add into a gather sits in shared memory.
// The seven ADMM state vectors are parked between solves either in shared memory (when they
// fit) or in a per-CTA L2-resident scratch.  HBM traffic per window is y in, x out.
//
// A persistent grid (<= CTAs that fit on the chip) strides over the batch; there is no
// inter-CTA communication, so windows shard over CTAs — and over GPUs — with no collective.
#include <algorithm>
#include <cstdio>

#pragma once
#include "mga_common.cuh"

namespace mga {

enum { ST_X = 0, ST_ZU, ST_ZD, ST_GU, ST_GD, ST_GAM, ST_PHI, ST_COUNT };

struct ResArgs {
  int N, T, t_in, n_outer, n_cg, NP, q1, want_diag, state_in_smem, nnz;
  int64_t B;
  int kd, ku;
  const int* nbr_d; const float* d_w; const int* nbr_u; const float* u_w;
  const int* csr_ptr; const int* csr_src; const float* csr_w;
  const float* y; float* x_out;
  float* out[ST_COUNT];       // optional per-window outputs (index by ST_*; ST_X unused)
  float* scratch;             // gridDim.x * ST_COUNT * T * N floats when !state_in_smem
  double* diag; double* dx_sum;
  float* alpha; float* beta;
  float rho, rho_u, rho_d, thr;
  float ax, cx, azu, czu, azd, czd;
  float t_mean, t_var;
};

template <int TT, int K>
struct Ctx {
  // per-thread constants
  int i, T, t_in, NP;
  bool active;
  int nd[K];      // smem column of the j-th temporal neighbour (N = zero slot)
  float wd[K];
  int nu[K - 1];  // spatial neighbours
  float wu[K - 1];
  int e0, e1;
  float* pbuf;
  float* qbuf;
  const int2* ent;  // (src column, weight bits)
  float* red;       // 2 x 32
  int red_sel;

  __device__ __forceinline__ float bsum(float v) {
    float* r = red + 32 * red_sel;
    red_sel ^= 1;
    return block_sum<float>(v, r);
  }

  __device__ __forceinline__ void put(float* buf, const float (&v)[TT]) {
#pragma unroll
    for (int t = 0; t < TT; ++t)
      if (t < T) buf[t * NP + i] = v[t];
  }

  // q = L_d v, reading v from pbuf (ADMM.py:166-177)
  __device__ __forceinline__ void ldr_from_pbuf(const float (&v)[TT], float (&q)[TT]) {
    q[0] = 0.f;
#pragma unroll
    for (int t = 1; t < TT; ++t) {
      float acc = 0.f;
      if (t < T) {
        const float* row = pbuf + (t - 1) * NP;
#pragma unroll
        for (int j = 0; j < K; ++j) acc += wd[j] * row[nd[j]];
      }
      q[t] = v[t] - acc;
    }
  }

  // f[t] = sum over the in-list of w * buf[t+1][src]   (ADMM.py:200-209 as a gather)
  __device__ __forceinline__ void father_sum(const float* buf, float (&f)[TT]) {
#pragma unroll
    for (int t = 0; t < TT; ++t) f[t] = 0.f;
    for (int e = e0; e < e1; ++e) {
      const int2 en = ent[e];
      const float w = __int_as_float(en.y);
      const float* col = buf + en.x;
#pragma unroll
      for (int t = 0; t < TT - 1; ++t)
        if (t + 1 < T) f[t] += w * col[(t + 1) * NP];
    }
  }

  // out = A v for the x / z_d systems: diag(v) + c * L_d^T L_d v  (ADMM.py:371-387, 392-394)
  template <bool XSYS>
  __device__ __forceinline__ void apply_cldr(const float (&v)[TT], float (&out)[TT], float a, float c) {
    put(pbuf, v);
    __syncthreads();
    float q[TT];
    ldr_from_pbuf(v, q);
    put(qbuf, q);
    __syncthreads();
    float f[TT];
    father_sum(qbuf, f);
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      const float l = (t == T - 1) ? q[t] : q[t] - f[t];     // q[0] == 0, so Q1 is moot here
      if (XSYS) out[t] = ((t < t_in ? v[t] : 0.f) + a * v[t]) + c * l;
      else out[t] = c * l + a * v[t];
    }
  }

  // out = mu_u L_u v + (rho_u/2) v  (ADMM.py:389-390)
  __device__ __forceinline__ void apply_lu(const float (&v)[TT], float (&out)[TT], float a, float c) {
    put(pbuf, v);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      float acc = 0.f;
      if (t < T) {
        const float* row = pbuf + t * NP;
#pragma unroll
        for (int j = 0; j < K - 1; ++j) acc += wu[j] * row[nu[j]];
      }
      out[t] = c * (v[t] - acc) + a * v[t];
    }
  }

  template <int SYS>
  __device__ __forceinline__ void apply(const float (&v)[TT], float (&out)[TT], float a, float c) {
    if (SYS == MGA_SYS_ZU) apply_lu(v, out, a, c);
    else if (SYS == MGA_SYS_X) apply_cldr<true>(v, out, a, c);
    else apply_cldr<false>(v, out, a, c);
  }

  // CG_solver, fixed iteration count (ADMM.py:329-368 with an unreachable tolerance).
  // On entry r holds the right-hand side and x the warm start.
  template <int SYS>
  __device__ __forceinline__ void cg(float (&x)[TT], float (&r)[TT], float a, float c, int n_cg, float* alpha_out,
                                     float* beta_out, int64_t B) {
    float p[TT], ap[TT];
    apply<SYS>(x, ap, a, c);
    float loc = 0.f;
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      r[t] = r[t] - ap[t];
      p[t] = r[t];
      loc += r[t] * r[t];
    }
    float rr = bsum(loc);
    for (int k = 0; k < n_cg; ++k) {
      apply<SYS>(p, ap, a, c);
      loc = 0.f;
#pragma unroll
      for (int t = 0; t < TT; ++t) loc += p[t] * ap[t];
      const float alpha = rr / bsum(loc);
      loc = 0.f;
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        x[t] = x[t] + alpha * p[t];
        r[t] = r[t] - alpha * ap[t];
        loc += r[t] * r[t];
      }
      const float rrn = bsum(loc);
      const float beta = rrn / rr;
      rr = rrn;
      if (alpha_out && threadIdx.x == 0) {
        alpha_out[(size_t)k * B] = alpha;
        beta_out[(size_t)k * B] = beta;
      }
#pragma unroll
      for (int t = 0; t < TT; ++t) p[t] = r[t] + beta * p[t];
    }
  }
};

__device__ __forceinline__ float soft_thr(float s, float d) {
  const float u = fabsf(s) - d;
  const float sg = (float)((s > 0.f) - (s < 0.f));
  return sg * u * (float)(u > 0.f);   // ADMM.py:407-408
}

template <int TT, int K, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_admm_resident(const ResArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, T = a.T, NP = a.NP, t_in = a.t_in;
  float* pbuf = reinterpret_cast<float*>(smem_raw);
  float* qbuf = pbuf + T * NP;
  float* red = qbuf + T * NP;                       // 64 floats
  float* dred = red + 64;                           // MGA_DIAG_COLS x 32 floats
  int2* ent = reinterpret_cast<int2*>(dred + MGA_DIAG_COLS * 32);
  float* st_smem = reinterpret_cast<float*>(ent + ((a.nnz + 1) & ~1));
  const int i = threadIdx.x;
  const bool active = i < N;

  Ctx<TT, K> c;
  c.i = i; c.T = T; c.t_in = t_in; c.NP = NP; c.active = active;
  c.pbuf = pbuf; c.qbuf = qbuf; c.ent = ent; c.red = red; c.red_sel = 0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    int nb = -1;
    float w = 0.f;
    if (active && j < a.kd) { nb = a.nbr_d[i * a.kd + j]; w = a.d_w[i * a.kd + j]; }
    c.nd[j] = nb >= 0 ? nb : N;
    c.wd[j] = nb >= 0 ? w : 0.f;
  }
#pragma unroll
  for (int j = 0; j < K - 1; ++j) {
    int nb = -1;
    float w = 0.f;
    if (active && j < a.ku) { nb = a.nbr_u[i * a.ku + j]; w = a.u_w[i * a.ku + j]; }
    c.nu[j] = nb >= 0 ? nb : N;
    c.wu[j] = nb >= 0 ? w : 0.f;
  }
  c.e0 = active ? a.csr_ptr[i] : 0;
  c.e1 = active ? a.csr_ptr[i + 1] : 0;
  for (int e = i; e < a.nnz; e += blockDim.x) ent[e] = make_int2(a.csr_src[e], __float_as_int(a.csr_w[e]));
  for (int k = i; k < 2 * T * NP; k += blockDim.x) pbuf[k] = 0.f;
  __syncthreads();

  const int NS = a.state_in_smem ? NP : N;
  float* state = a.state_in_smem ? st_smem : a.scratch + (size_t)blockIdx.x * ST_COUNT * T * N;
  const int col = active ? i : 0;
#define LD(V, t) (state[((V) * T + (t)) * NS + col])
#define ST(V, t, val) do { if (active) state[((V) * T + (t)) * NS + col] = (val); } while (0)

  for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
    const float* yw = a.y + (size_t)b * t_in * N + col;
    float x[TT];
    // ---- initial_guess (ADMM.py:766-781) and initial state (ADMM.py:537-544)
    {
      float sy = 0.f, sty = 0.f;
      for (int t = 0; t < t_in; ++t) {
        const float v = active ? yw[(size_t)t * N] : 0.f;
        sy += v;
        sty += (float)t * v;
      }
      const float my = sy / (float)t_in, mty = sty / (float)t_in;
      const float w = (mty - a.t_mean * my) / a.t_var;
      const float cc = my - w * a.t_mean;
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        float v = 0.f;
        if (active && t < T) v = t < t_in ? yw[(size_t)t * N] : w * (float)t + cc;
        x[t] = v;
        if (t < T) {
          ST(ST_X, t, v); ST(ST_ZU, t, v); ST(ST_ZD, t, v);
          ST(ST_GU, t, 0.1f); ST(ST_GD, t, 0.1f); ST(ST_GAM, t, 0.1f);
        }
      }
      c.put(pbuf, x);
      __syncthreads();
      float q[TT];
      c.ldr_from_pbuf(x, q);
#pragma unroll
      for (int t = 0; t < TT; ++t)
        if (t < T) ST(ST_PHI, t, q[t]);
      __syncthreads();
    }

    for (int it = 0; it < a.n_outer; ++it) {
      float dg[MGA_DIAG_COLS];
#pragma unroll
      for (int k = 0; k < MGA_DIAG_COLS; ++k) dg[k] = 0.f;
      float* al = a.alpha ? a.alpha + ((size_t)it * 3) * a.n_cg * a.B + b : nullptr;
      float* be = a.beta ? a.beta + ((size_t)it * 3) * a.n_cg * a.B + b : nullptr;
      const size_t sys_stride = (size_t)a.n_cg * a.B;
      float r[TT];
      // ---- RHS_x (ADMM.py:552-559): Ldr_T(gamma + rho phi)/2 + (rho_u zu + rho_d zd)/2 - (gu+gd)/2 + H^T y
      {
        float v[TT], f[TT];
#pragma unroll
        for (int t = 0; t < TT; ++t) v[t] = (active && t < T) ? LD(ST_GAM, t) + a.rho * LD(ST_PHI, t) : 0.f;
        c.put(qbuf, v);
        __syncthreads();
        c.father_sum(qbuf, f);
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          float l = (t == T - 1) ? v[t] : ((t == 0 && !a.q1) ? -f[t] : v[t] - f[t]);
          float o = 0.f;
          if (active && t < T) {
            const float hty = t < t_in ? yw[(size_t)t * N] : 0.f;
            o = l / 2.f + (a.rho_u * LD(ST_ZU, t) + a.rho_d * LD(ST_ZD, t)) / 2.f - (LD(ST_GU, t) + LD(ST_GD, t)) / 2.f + hty;
          }
          r[t] = o;
        }
      }
      // ---- x solve (ADMM.py:571); x registers hold x_old
      c.template cg<MGA_SYS_X>(x, r, a.ax, a.cx, a.n_cg, al, be, a.B);
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        if (active && t < T) {
          if (a.want_diag) {
            const float dx = x[t] - LD(ST_X, t);
            dg[MGA_DIAG_DX2] += dx * dx;
            if (a.dx_sum) atomicAdd(a.dx_sum + ((size_t)it * T + t) * N + i, (double)dx);
          }
          ST(ST_X, t, x[t]);
        }
      }
      // ---- z_u solve (ADMM.py:579-580) + its dual ascent (ADMM.py:595)
      {
        float z[TT];
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          const bool on = active && t < T;
          z[t] = on ? LD(ST_ZU, t) : 0.f;
          r[t] = on ? LD(ST_GU, t) / 2.f + a.azu * LD(ST_X, t) : 0.f;
        }
        c.template cg<MGA_SYS_ZU>(z, r, a.azu, a.czu, a.n_cg, al ? al + sys_stride : nullptr,
                                  be ? be + sys_stride : nullptr, a.B);
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          if (active && t < T) {
            const float d0 = LD(ST_X, t) - z[t];
            if (a.want_diag) {
              const float d1 = z[t] - LD(ST_ZU, t);
              dg[MGA_DIAG_X_ZU2] += d0 * d0;
              dg[MGA_DIAG_DZU2] += d1 * d1;
            }
            ST(ST_GU, t, LD(ST_GU, t) + a.rho_u * d0);
            ST(ST_ZU, t, z[t]);
          }
        }
      }
      // ---- z_d solve (ADMM.py:587-588) + its dual ascent (ADMM.py:597)
      {
        float z[TT];
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          const bool on = active && t < T;
          z[t] = on ? LD(ST_ZD, t) : 0.f;
          r[t] = on ? LD(ST_GD, t) / 2.f + a.azd * LD(ST_X, t) : 0.f;
        }
        c.template cg<MGA_SYS_ZD>(z, r, a.azd, a.czd, a.n_cg, al ? al + 2 * sys_stride : nullptr,
                                  be ? be + 2 * sys_stride : nullptr, a.B);
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          if (active && t < T) {
            const float d0 = LD(ST_X, t) - z[t];
            if (a.want_diag) {
              const float d1 = z[t] - LD(ST_ZD, t);
              dg[MGA_DIAG_X_ZD2] += d0 * d0;
              dg[MGA_DIAG_DZD2] += d1 * d1;
            }
            ST(ST_GD, t, LD(ST_GD, t) + a.rho_d * d0);
            ST(ST_ZD, t, z[t]);
          }
        }
      }
      // ---- phi prox + gamma ascent (ADMM.py:600-605) and the remaining diagnostics (ADMM.py:612-637)
      {
        // x was parked in the state block during the z solves (keeps it out of the CG register budget)
#pragma unroll
        for (int t = 0; t < TT; ++t) x[t] = (active && t < T) ? LD(ST_X, t) : 0.f;
        c.put(pbuf, x);
        __syncthreads();
        float q[TT];
        c.ldr_from_pbuf(x, q);
        int bad = 0;
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          if (active && t < T) {
            const float gv = LD(ST_GAM, t), po = LD(ST_PHI, t);
            const float ph = soft_thr(q[t] - gv / a.rho, a.thr);
            const float gn = gv + a.rho * (ph - q[t]);
            ST(ST_PHI, t, ph);
            ST(ST_GAM, t, gn);
            bad |= !isfinite(x[t]) || !isfinite(ph) || !isfinite(gn);
            if (a.want_diag) {
              const float e = ph - q[t], f = ph - po;
              dg[MGA_DIAG_PHI_LDX2] += e * e;
              dg[MGA_DIAG_DPHI2] += f * f;
              dg[MGA_DIAG_DGTV] += fabsf(q[t]);
              dg[MGA_DIAG_DGLR] += q[t] * q[t];
              if (t < t_in) {
                const float h = x[t] - yw[(size_t)t * N];
                dg[MGA_DIAG_RECOVER2] += h * h;
              }
              float acc = 0.f;
              const float* row = pbuf + t * NP;
#pragma unroll
              for (int j = 0; j < K - 1; ++j) acc += c.wu[j] * row[c.nu[j]];
              dg[MGA_DIAG_GLR] += x[t] * (x[t] - acc);
            }
          }
        }
        dg[MGA_DIAG_NONFINITE] = (float)bad;
        // one-sync multi-column block reduction
        const int lane = i & 31, wp = i >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int k = 0; k < MGA_DIAG_COLS; ++k) {
          const float v = warp_sum<float>(dg[k]);
          if (lane == 0) dred[k * 32 + wp] = v;
        }
        __syncthreads();
        if (i < MGA_DIAG_COLS && a.diag) {
          float tot = 0.f;
          for (int k = 0; k < nw; ++k) tot += dred[i * 32 + k];
          if (tot != 0.f) atomicAdd(a.diag + (size_t)it * MGA_DIAG_COLS + i, (double)tot);
        }
        __syncthreads();
      }
    }
    // ---- results
    if (active) {
#pragma unroll
      for (int t = 0; t < TT; ++t)
        if (t < T) a.x_out[((size_t)b * T + t) * N + i] = x[t];
      for (int v = ST_ZU; v < ST_COUNT; ++v)
        if (a.out[v])
          for (int t = 0; t < T; ++t) a.out[v][((size_t)b * T + t) * N + i] = LD(v, t);
    }
    __syncthreads();
  }
#undef LD
#undef ST
}

inline size_t res_core_bytes(const GraphDev& g, int NP) {
  return (size_t)2 * g.T * NP * 4 + 64 * 4 + MGA_DIAG_COLS * 32 * 4 + (size_t)((g.nnz + 1) & ~1) * 8;
}


template <int TT, int K, int MAXT>
inline int launch_res(mga_plan* p, ResArgs& a, int threads, cudaStream_t st) {
  const GraphDev& g = p->g;
  auto kern = k_admm_resident<TT, K, MAXT>;
  const size_t core = res_core_bytes(g, a.NP);
  const size_t with_state = core + (size_t)ST_COUNT * g.T * a.NP * 4;
  // State in shared memory only if it does not cost residency: compare CTAs/SM both ways.
  int occ_core = 0, occ_state = 0;
  MGA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem_optin));
  MGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_core, kern, threads, core));
  if (with_state <= (size_t)p->max_smem_optin)
    MGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_state, kern, threads, with_state));
  if (occ_core < 1) { set_error("resident kernel does not fit on an SM"); return MGA_ERR_UNSUPPORTED; }
  a.state_in_smem = (occ_state >= occ_core) ? 1 : 0;
  const int occ = a.state_in_smem ? occ_state : occ_core;
  const size_t smem = a.state_in_smem ? with_state : core;
  int64_t grid = std::min<int64_t>(a.B, (int64_t)occ * p->sm_count);
  if (!a.state_in_smem) {
    const size_t need = (size_t)grid * ST_COUNT * g.T * g.N * sizeof(float);
    int rc = ensure_workspace(p, p->ws, need);
    if (rc) return rc;
    a.scratch = static_cast<float*>(p->ws.base);
  }
  kern<<<(unsigned)grid, threads, smem, st>>>(a);
  MGA_LAUNCH_CHECK("k_admm_resident");
  return MGA_OK;
}

template <int TT, int K>
inline int pick_threads(mga_plan* p, ResArgs& a, int threads, cudaStream_t st) {
  if (threads <= 256) return launch_res<TT, K, 256>(p, a, threads, st);
  if (threads <= 384) return launch_res<TT, K, 384>(p, a, threads, st);
  return launch_res<TT, K, 512>(p, a, threads, st);
}


}  // namespace mga
