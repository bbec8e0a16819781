// Streaming mode: ADMM state lives in HBM / L2, one thread per lattice point, one fused kernel
// per CG phase.  Works for every shape and both dtypes; it is the only mode for windows that
// do not fit one SM (T=288, N=20k) and for the tolerance-driven (parity) runs.
//
// Per CG iteration (ADMM.py:348-366) the x / z_d systems run three kernels
//     cg_p1a:  p' = r + beta p ; q = L_d p'          reads r,p   writes p',q   (16 B/pt)
//     cg_p1b:  Ap = D p' + c L_d^T q ; <p',Ap>        reads p',q  writes Ap     (12 B/pt)
//     cg_p2 :  x += a p' ; r -= a Ap ; <r,r>          reads x,p',r,Ap writes x,r (24 B/pt)
// and the z_u system two (cg_p1_lu fuses the p update with the 1-hop L_u).  Neighbour values of
// p' are recomputed from (r, p) at the gather so the p update needs no separate pass; p is
// ping-ponged between two buffers.  Per-window dot products: warp shuffle -> block -> one double
// atomicAdd per CTA into a per-iteration slot, so no kernel ever waits on another.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "mga_common.cuh"

namespace mga {

constexpr int kBlock = 256;

// ---------------------------------------------------------------------------------------------
// value loaders: how a kernel reads the vector an operator is applied to
template <typename S>
struct LoadVec {
  const S* v;
  __device__ __forceinline__ S operator()(int64_t k) const { return v[k]; }
};
template <typename S>
struct LoadPNew {   // p' = r + beta p  (first iteration: p' = r)
  const S* r;
  const S* p;
  S beta;
  bool first;
  __device__ __forceinline__ S operator()(int64_t k) const { return first ? r[k] : r[k] + beta * p[k]; }
};
template <typename S>
struct LoadAxpy {   // a + s * b   (gamma + rho * phi, ADMM.py:559)
  const S* a;
  const S* b;
  S s;
  __device__ __forceinline__ S operator()(int64_t k) const { return a[k] + s * b[k]; }
};

// L_u v at (t, i); `w0` = offset of window start.  ADMM.py:138-148
template <typename S, class Load>
__device__ __forceinline__ S lu_at(const GraphDev& g, const Load& ld, int64_t w0, int t, int i, S self) {
  const int64_t row = w0 + (int64_t)t * g.N;
  const int* nb = g.nbr_u + (size_t)i * g.ku;
  const float* w = g.u_w + ((size_t)(g.u_wT > 1 ? t : 0) * g.N + i) * g.ku;
  S acc = 0;
  for (int j = 0; j < g.ku; ++j) {
    const int c = nb[j];
    if (c >= 0) acc += (S)w[j] * ld(row + c);
  }
  return self - acc;
}

// L_d v at (t, i).  ADMM.py:150-177
template <typename S, class Load>
__device__ __forceinline__ S ldr_at(const GraphDev& g, const Load& ld, int64_t w0, int t, int i, S self) {
  if (t == 0) return (S)0;
  if (g.temporal == MGA_TEMPORAL_BAND) {
    S acc = 0;
    for (int s = 0; s < g.skip; ++s) {
      const int ts = t - 1 - s;
      if (ts >= 0) acc += (S)g.band_w[((size_t)t * g.skip + s) * g.N + i] * ld(w0 + (int64_t)ts * g.N + i);
    }
    return self - acc;
  }
  const int64_t row = w0 + (int64_t)(t - 1) * g.N;
  const int* nb = g.nbr_d + (size_t)i * g.kd;
  const float* w = g.d_w + ((size_t)(g.d_wT > 1 ? t - 1 : 0) * g.N + i) * g.kd;
  S acc = 0;
  for (int j = 0; j < g.kd; ++j) {
    const int c = nb[j];
    if (c >= 0) acc += (S)w[j] * ld(row + c);
  }
  return self - acc;
}

// (L_d^T + Q1) v at (t, i): the "father" sum over the in-list, then the row rules of
// ADMM.py:217-223 (and 183-186 / 191-193 for the line-graph variants).
template <typename S, class Load>
__device__ __forceinline__ S ldrt_at(const GraphDev& g, const Load& ld, int64_t w0, int t, int i, S self) {
  if (t == g.T - 1) return self;
  S f = 0;
  if (g.temporal == MGA_TEMPORAL_BAND) {
    for (int s = 0; s < g.skip; ++s) {
      const int ts = t + 1 + s;
      if (ts < g.T) f += (S)g.band_w[((size_t)ts * g.skip + s) * g.N + i] * ld(w0 + (int64_t)ts * g.N + i);
    }
  } else {
    const int64_t row = w0 + (int64_t)(t + 1) * g.N;
    const int e0 = g.csr_ptr[i], e1 = g.csr_ptr[i + 1];
    if (g.d_wT > 1) {
      const float* w = g.d_w + (size_t)t * g.N * g.kd;
      for (int e = e0; e < e1; ++e) f += (S)w[g.csr_slot[e]] * ld(row + g.csr_src[e]);
    } else {
      for (int e = e0; e < e1; ++e) f += (S)g.csr_w[e] * ld(row + g.csr_src[e]);
    }
  }
  if (t == 0 && !g.q1) return -f;
  return self - f;
}

struct Pt {
  int64_t b, w0, k;   // window, offset of window start, flat offset of this point
  int t, i;
  bool ok;
};

__device__ __forceinline__ Pt locate(const GraphDev& g, int chunks) {
  Pt p;
  const int n = g.T * g.N;
  p.b = blockIdx.x / chunks;
  const int local = (blockIdx.x % chunks) * kBlock + threadIdx.x;
  p.ok = local < n;
  const int l = p.ok ? local : 0;
  p.t = l / g.N;
  p.i = l - p.t * g.N;
  p.w0 = p.b * (int64_t)n;
  p.k = p.w0 + l;
  return p;
}

// diagonal part of a system matrix: a*x + [H^T H x], in the reference's order (ADMM.py:372-387)
template <typename S>
struct Diag {
  S a;          // coefficient of x
  int hth;      // 1: add H^T H x (rows t < t_in, or mask)
  const S* mask;
  int t_in;
  __device__ __forceinline__ S operator()(S x, int t, int64_t k) const {
    S hx = (S)0;
    if (hth) hx = mask ? x * mask[k] : (t < t_in ? x : (S)0);
    return hth ? hx + a * x : a * x;
  }
};

// ---------------------------------------------------------------------------------------------
// standalone operators
enum { kOpLu = 0, kOpLdr = 1, kOpLdrT = 2 };

// out = op(v)                          when lhs == 0
// out = diag(xin) + c * op(v)          when lhs == 1   (LHS_zu / second half of LHS_x, LHS_zd)
template <typename S, int OP>
__global__ void __launch_bounds__(kBlock) k_apply(GraphDev g, int chunks, const S* __restrict__ v,
                                                   const S* __restrict__ xin, S* __restrict__ out, int lhs,
                                                   Diag<S> dg, S c, int zu_order) {
  const Pt p = locate(g, chunks);
  if (!p.ok) return;
  LoadVec<S> ld{v};
  const S self = v[p.k];
  S o;
  if (OP == kOpLu) o = lu_at<S>(g, ld, p.w0, p.t, p.i, self);
  else if (OP == kOpLdr) o = ldr_at<S>(g, ld, p.w0, p.t, p.i, self);
  else o = ldrt_at<S>(g, ld, p.w0, p.t, p.i, self);
  if (lhs) {
    const S xv = xin[p.k];
    // LHS_zu / LHS_zd: c*op + a*x (ADMM.py:390, 394);  LHS_x: (HtHx + a*x) + c*op (ADMM.py:379)
    o = zu_order ? c * o + dg(xv, p.t, p.k) : dg(xv, p.t, p.k) + c * o;
  }
  out[p.k] = o;
}

// ---------------------------------------------------------------------------------------------
// CG kernels.  dots: (2*max_iter+1, B) doubles; RR(k) = dots[2k], PAP(k) = dots[2k+1].
template <typename S>
__device__ __forceinline__ void block_accumulate(S v, double* slot) {
  __shared__ S red[32];
  const S tot = block_sum<S>(v, red);
  if (threadIdx.x == 0) atomicAdd(slot, (double)tot);
}

// r = rhs - Ax0 ; p = r ; RR(0) += r.r ; optionally x_out = x_in (ping-pong of the iterate)
template <typename S>
__global__ void __launch_bounds__(kBlock) k_cg_init(GraphDev g, int chunks, const S* __restrict__ rhs,
                                                     const S* __restrict__ ax, S* __restrict__ r,
                                                     double* __restrict__ rr0) {
  const Pt p = locate(g, chunks);
  S rv = 0;
  if (p.ok) {
    rv = rhs[p.k] - ax[p.k];
    r[p.k] = rv;
  }
  block_accumulate<S>(rv * rv, rr0 + p.b);
}

template <typename S>
__device__ __forceinline__ S beta_of(const double* dots, int64_t B, int64_t b, int k) {
  // beta_{k-1} = rr_k / rr_{k-1} in the signal dtype (ADMM.py:356)
  return (S)dots[(size_t)(2 * k) * B + b] / (S)dots[(size_t)(2 * k - 2) * B + b];
}

// z_u system, phase 1: p' = r + beta p ; Ap = mu_u L_u p' + (rho_u/2) p' ; PAP(k) += p'.Ap
template <typename S>
__global__ void __launch_bounds__(kBlock) k_cg_p1_lu(GraphDev g, int chunks, int64_t B, int k,
                                                      const S* __restrict__ r, const S* __restrict__ p_old,
                                                      S* __restrict__ p_new, S* __restrict__ ap,
                                                      double* __restrict__ dots, Diag<S> dg, S c) {
  const Pt p = locate(g, chunks);
  S prod = 0;
  if (p.ok) {
    LoadPNew<S> ld{r, p_old, k > 0 ? beta_of<S>(dots, B, p.b, k) : (S)0, k == 0};
    const S pv = ld(p.k);
    const S l = lu_at<S>(g, ld, p.w0, p.t, p.i, pv);
    const S a = c * l + dg(pv, p.t, p.k);
    p_new[p.k] = pv;
    ap[p.k] = a;
    prod = pv * a;
  }
  block_accumulate<S>(prod, dots + (size_t)(2 * k + 1) * B + p.b);
}

// x / z_d systems, phase 1a: p' = r + beta p ; q = L_d p'
template <typename S>
__global__ void __launch_bounds__(kBlock) k_cg_p1a(GraphDev g, int chunks, int64_t B, int k,
                                                    const S* __restrict__ r, const S* __restrict__ p_old,
                                                    S* __restrict__ p_new, S* __restrict__ q,
                                                    const double* __restrict__ dots) {
  const Pt p = locate(g, chunks);
  if (!p.ok) return;
  LoadPNew<S> ld{r, p_old, k > 0 ? beta_of<S>(dots, B, p.b, k) : (S)0, k == 0};
  const S pv = ld(p.k);
  p_new[p.k] = pv;
  q[p.k] = ldr_at<S>(g, ld, p.w0, p.t, p.i, pv);
}

// phase 1b: Ap = diag(p') + c L_d^T q ; PAP(k) += p'.Ap
template <typename S>
__global__ void __launch_bounds__(kBlock) k_cg_p1b(GraphDev g, int chunks, int64_t B, int k,
                                                    const S* __restrict__ pv_, const S* __restrict__ q,
                                                    S* __restrict__ ap, double* __restrict__ dots, Diag<S> dg, S c,
                                                    int zu_order) {
  const Pt p = locate(g, chunks);
  S prod = 0;
  if (p.ok) {
    LoadVec<S> ld{q};
    const S pv = pv_[p.k];
    const S l = ldrt_at<S>(g, ld, p.w0, p.t, p.i, q[p.k]);
    const S a = zu_order ? c * l + dg(pv, p.t, p.k) : dg(pv, p.t, p.k) + c * l;
    ap[p.k] = a;
    prod = pv * a;
  }
  block_accumulate<S>(prod, dots + (size_t)(2 * k + 1) * B + p.b);
}

// diagonal system (ablation 'DGTV' / 'UT': LHS_x has no graph term, ADMM.py:381, 385)
template <typename S>
__global__ void __launch_bounds__(kBlock) k_cg_p1_diag(GraphDev g, int chunks, int64_t B, int k,
                                                        const S* __restrict__ r, const S* __restrict__ p_old,
                                                        S* __restrict__ p_new, S* __restrict__ ap,
                                                        double* __restrict__ dots, Diag<S> dg) {
  const Pt p = locate(g, chunks);
  S prod = 0;
  if (p.ok) {
    LoadPNew<S> ld{r, p_old, k > 0 ? beta_of<S>(dots, B, p.b, k) : (S)0, k == 0};
    const S pv = ld(p.k);
    const S a = dg(pv, p.t, p.k);
    p_new[p.k] = pv;
    ap[p.k] = a;
    prod = pv * a;
  }
  block_accumulate<S>(prod, dots + (size_t)(2 * k + 1) * B + p.b);
}

// phase 2: alpha = rr/pAp ; x += alpha p ; r -= alpha Ap ; RR(k+1) += r.r   (ADMM.py:350-355)
template <typename S>
__global__ void __launch_bounds__(kBlock) k_cg_p2(GraphDev g, int chunks, int64_t B, int k, S* __restrict__ x,
                                                   S* __restrict__ r, const S* __restrict__ pv,
                                                   const S* __restrict__ ap, double* __restrict__ dots) {
  const Pt p = locate(g, chunks);
  S rv = 0;
  if (p.ok) {
    const S alpha = (S)dots[(size_t)(2 * k) * B + p.b] / (S)dots[(size_t)(2 * k + 1) * B + p.b];
    x[p.k] = x[p.k] + alpha * pv[p.k];
    rv = r[p.k] - alpha * ap[p.k];
    r[p.k] = rv;
  }
  block_accumulate<S>(rv * rv, dots + (size_t)(2 * k + 2) * B + p.b);
}

// alpha_k, beta_k for the lists (ADMM.py:351, 357)
template <typename S>
__global__ void k_cg_coeffs(int64_t B, int iters, const double* __restrict__ dots, S* __restrict__ alpha,
                            S* __restrict__ beta) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * iters) return;
  const int k = (int)(idx / B);
  const int64_t b = idx - (int64_t)k * B;
  const S rr = (S)dots[(size_t)(2 * k) * B + b], pap = (S)dots[(size_t)(2 * k + 1) * B + b];
  const S rrn = (S)dots[(size_t)(2 * k + 2) * B + b];
  if (alpha) alpha[idx] = rr / pap;
  if (beta) beta[idx] = rrn / rr;
}

// stop test: max_b sqrt(rr_b) < tol  (ADMM.py:360; NaN anywhere => not converged, like torch.max)
template <typename S>
__global__ void k_cg_check(int64_t B, const double* __restrict__ rr, S tol, int* __restrict__ flag) {
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  int bad = 0;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
    const S v = sqrt((S)rr[b]);
    if (!(v < tol)) bad = 1;
  }
  if (bad) atomicOr(&s_bad, 1);
  __syncthreads();
  if (threadIdx.x == 0) *flag = s_bad ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------
// elementwise / prologue kernels of combined_loop

// initial_guess (ADMM.py:766-781) + initial state (ADMM.py:537-544, phi filled by a later k_apply)
template <typename S>
__global__ void __launch_bounds__(kBlock) k_init(GraphDev g, int64_t B, const S* __restrict__ y, S* __restrict__ x,
                                                  S* __restrict__ zu, S* __restrict__ zd, S* __restrict__ gu,
                                                  S* __restrict__ gd, S* __restrict__ gam, float t_mean,
                                                  float t_var) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * g.N) return;
  const int64_t b = idx / g.N;
  const int i = (int)(idx - b * g.N);
  const S* yw = y + b * (int64_t)g.t_in * g.N + i;
  S sy = 0, sty = 0;
  for (int t = 0; t < g.t_in; ++t) {
    const S v = yw[(size_t)t * g.N];
    sy += v;
    sty += (S)(float)t * v;
  }
  const S my = sy / (S)g.t_in, mty = sty / (S)g.t_in;
  const S w = (mty - (S)t_mean * my) / (S)t_var;
  const S c = my - w * (S)t_mean;
  const int64_t o = b * (int64_t)g.T * g.N + i;
  for (int t = 0; t < g.T; ++t) {
    const S v = t < g.t_in ? yw[(size_t)t * g.N] : w * (S)(float)t + c;
    const int64_t k = o + (int64_t)t * g.N;
    x[k] = v;
    if (zu) { zu[k] = v; zd[k] = v; gu[k] = (S)0.1; gd[k] = (S)0.1; }
    if (gam) gam[k] = (S)0.1;
  }
}

// state init when x0 is given (mask mode: x0 from initial_interpolation on the host side)
template <typename S>
__global__ void __launch_bounds__(kBlock) k_init_from_x(int64_t n, const S* __restrict__ x, S* __restrict__ zu,
                                                         S* __restrict__ zd, S* __restrict__ gu, S* __restrict__ gd,
                                                         S* __restrict__ gam) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const S v = x[k];
  zu[k] = v; zd[k] = v; gu[k] = (S)0.1; gd[k] = (S)0.1;
  if (gam) gam[k] = (S)0.1;
}

// initial_interpolation (ADMM.py:783-811), B == 1 semantics per window
template <typename S>
__global__ void __launch_bounds__(kBlock) k_interp(GraphDev g, int64_t B, const S* __restrict__ y,
                                                    const S* __restrict__ mask, S* __restrict__ x) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * g.N) return;
  const int64_t b = idx / g.N;
  const int i = (int)(idx - b * g.N);
  const int64_t o = b * (int64_t)g.T * g.N + i;
  S cnt = 0, st = 0, sy = 0, sty = 0, st2 = 0;
  for (int t = 0; t < g.T; ++t) {
    const S m = mask[o + (int64_t)t * g.N], v = y[o + (int64_t)t * g.N], tt = (S)(float)t;
    cnt += m; st += tt * m; sy += v * m; sty += tt * v * m; st2 += tt * tt * m;
  }
  const S tm = st / cnt, ym = sy / cnt, tym = sty / cnt, t2m = st2 / cnt;
  const S w = (tym - tm * ym) / (t2m - tm * tm);
  const S c = ym - w * tm;
  for (int t = 0; t < g.T; ++t) {
    const int64_t k = o + (int64_t)t * g.N;
    x[k] = (w * (S)(float)t + c) * ((S)1 - mask[k]) + y[k];
  }
}

// RHS_x (ADMM.py:552-566)
template <typename S>
__global__ void __launch_bounds__(kBlock) k_rhs_x(GraphDev g, int chunks, int abl, const S* __restrict__ gam,
                                                   const S* __restrict__ phi, const S* __restrict__ zu,
                                                   const S* __restrict__ zd, const S* __restrict__ gu,
                                                   const S* __restrict__ gd, const S* __restrict__ y, int y_rows,
                                                   S* __restrict__ rhs, S rho, S rho_u, S rho_d) {
  const Pt p = locate(g, chunks);
  if (!p.ok) return;
  const S hty = p.t < y_rows ? y[p.b * (int64_t)y_rows * g.N + (int64_t)p.t * g.N + p.i] : (S)0;
  S o;
  if (abl == MGA_ABL_NONE || abl == MGA_ABL_DGLR) {
    LoadAxpy<S> ld{gam, phi, rho};
    const S l = ldrt_at<S>(g, ld, p.w0, p.t, p.i, ld(p.k));
    if (abl == MGA_ABL_NONE) o = l / (S)2 + (rho_u * zu[p.k] + rho_d * zd[p.k]) / (S)2 - (gu[p.k] + gd[p.k]) / (S)2 + hty;
    else o = l / (S)2 + rho_u * zu[p.k] / (S)2 - gu[p.k] / (S)2 + hty;
  } else {
    o = (rho_u * zu[p.k] + rho_d * zd[p.k]) / (S)2 - (gu[p.k] + gd[p.k]) / (S)2 + hty;
  }
  rhs[p.k] = o;
}

// RHS_zu / RHS_zd (ADMM.py:579, 587): gamma_z / 2 + rho_z / 2 * x
template <typename S>
__global__ void __launch_bounds__(kBlock) k_rhs_z(int64_t n, const S* __restrict__ gz, const S* __restrict__ x,
                                                   S* __restrict__ rhs, S half_rho) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) rhs[k] = gz[k] / (S)2 + half_rho * x[k];
}

template <typename S>
__global__ void __launch_bounds__(kBlock) k_dual(int64_t n, S rho_z, const S* __restrict__ x, const S* __restrict__ z,
                                                  S* __restrict__ gz) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) gz[k] = gz[k] + rho_z * (x[k] - z[k]);
}

template <typename S>
__device__ __forceinline__ S soft(S s, S d) {
  const S u = fabs(s) - d;
  const S sg = (S)((s > (S)0) - (s < (S)0));
  return sg * u * (S)(u > (S)0);   // ADMM.py:407-408
}

// phi_direct alone / phi + gamma ascent (ADMM.py:401-408, 603-605)
template <typename S>
__global__ void __launch_bounds__(kBlock) k_phi(GraphDev g, int chunks, const S* __restrict__ x,
                                                 const S* __restrict__ gam_in, S* __restrict__ gam_out,
                                                 S* __restrict__ phi, S rho, S thr) {
  const Pt p = locate(g, chunks);
  if (!p.ok) return;
  LoadVec<S> ld{x};
  const S ldx = ldr_at<S>(g, ld, p.w0, p.t, p.i, x[p.k]);
  const S gv = gam_in[p.k];
  const S ph = soft<S>(ldx - gv / rho, thr);
  phi[p.k] = ph;
  if (gam_out) gam_out[p.k] = gv + rho * (ph - ldx);
}

// The tail of one outer iteration in a single pass (ADMM.py:595-637): dual ascent, phi prox,
// gamma ascent and all diagnostics.  Sums go to diag[] (double atomics, one per CTA and column).
template <typename S>
__global__ void __launch_bounds__(kBlock) k_tail(GraphDev g, int chunks, int abl, int want_diag,
                                                  const S* __restrict__ x, const S* __restrict__ x_old,
                                                  const S* __restrict__ zu, const S* __restrict__ zu_old,
                                                  const S* __restrict__ zd, const S* __restrict__ zd_old,
                                                  S* __restrict__ gu, S* __restrict__ gd, S* __restrict__ gam,
                                                  S* __restrict__ phi, const S* __restrict__ y, int y_rows,
                                                  const S* __restrict__ mask, S rho, S rho_u, S rho_d, S thr,
                                                  double* __restrict__ diag, double* __restrict__ dx_sum) {
  const Pt p = locate(g, chunks);
  const bool with_phi = abl == MGA_ABL_NONE || abl == MGA_ABL_DGLR;
  const bool with_zd = abl != MGA_ABL_DGLR;
  S d[MGA_DIAG_COLS];
#pragma unroll
  for (int c = 0; c < MGA_DIAG_COLS; ++c) d[c] = 0;
  if (p.ok) {
    const S xv = x[p.k], zuv = zu[p.k];
    const S zdv = with_zd ? zd[p.k] : (S)0;
    gu[p.k] = gu[p.k] + rho_u * (xv - zuv);
    if (with_zd) gd[p.k] = gd[p.k] + rho_d * (xv - zdv);
    LoadVec<S> ld{x};
    S ldx = 0, ph = 0, ph_old = 0;
    int bad = !isfinite(xv) || !isfinite(zuv) || !isfinite(zdv);
    if (with_phi || want_diag) ldx = ldr_at<S>(g, ld, p.w0, p.t, p.i, xv);
    if (with_phi) {
      const S gv = gam[p.k];
      ph_old = phi[p.k];
      ph = soft<S>(ldx - gv / rho, thr);
      phi[p.k] = ph;
      const S gn = gv + rho * (ph - ldx);
      gam[p.k] = gn;
      bad |= !isfinite(ph) || !isfinite(gn);
    }
    d[MGA_DIAG_NONFINITE] = (S)bad;
    if (want_diag) {
      const S dx = xv - x_old[p.k];
      d[MGA_DIAG_DX2] = dx * dx;
      if (dx_sum) atomicAdd(dx_sum + (size_t)p.t * g.N + p.i, (double)dx);
      const S a = xv - zuv, bz = zuv - zu_old[p.k];
      d[MGA_DIAG_X_ZU2] = a * a;
      d[MGA_DIAG_DZU2] = bz * bz;
      d[MGA_DIAG_GLR] = xv * lu_at<S>(g, ld, p.w0, p.t, p.i, xv);
      if (mask) {
        const S h = xv * mask[p.k] - y[p.k];
        d[MGA_DIAG_RECOVER2] = h * h;
      } else if (p.t < y_rows) {
        const S h = xv - y[p.b * (int64_t)y_rows * g.N + (int64_t)p.t * g.N + p.i];
        d[MGA_DIAG_RECOVER2] = h * h;
      }
      if (with_phi) {
        const S e = ph - ldx, f = ph - ph_old;
        d[MGA_DIAG_PHI_LDX2] = e * e;
        d[MGA_DIAG_DPHI2] = f * f;
        d[MGA_DIAG_DGTV] = fabs(ldx);
      }
      if (with_zd) {
        const S e = xv - zdv, f = zdv - zd_old[p.k];
        d[MGA_DIAG_X_ZD2] = e * e;
        d[MGA_DIAG_DZD2] = f * f;
        d[MGA_DIAG_DGLR] = ldx * ldx;
      }
    }
  }
  // block reduction of all columns: warp shuffle, then one warp finishes
  __shared__ S red[MGA_DIAG_COLS][kBlock / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < MGA_DIAG_COLS; ++c) {
    const S v = warp_sum<S>(d[c]);
    if (lane == 0) red[c][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < MGA_DIAG_COLS) {
    S t = 0;
    for (int k = 0; k < kBlock / 32; ++k) t += red[threadIdx.x][k];
    if (t != (S)0) atomicAdd(diag + threadIdx.x, (double)t);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
struct Sys {
  int kind;        // 0: L_u, 1: cLdr, 2: diagonal only
  double a;        // coefficient of x
  int hth;         // H^T H term present
  double c;        // coefficient of the graph operator
  int zu_order;    // c*op + a*x instead of (hx + a*x) + c*op
};

static Sys system_of(int system, const mga_params& m) {
  Sys s{};
  if (system == MGA_SYS_X) {
    s.hth = 1;
    if (m.ablation == MGA_ABL_NONE) { s.kind = 1; s.a = (m.rho_u + m.rho_d) / 2; s.c = m.rho / 2; }
    else if (m.ablation == MGA_ABL_DGLR) { s.kind = 1; s.a = m.rho_u / 2; s.c = m.rho / 2; }
    else { s.kind = 2; s.a = (m.rho_u + m.rho_d) / 2; s.c = 0; }
  } else if (system == MGA_SYS_ZU) {
    s.kind = 0; s.a = m.rho_u / 2; s.c = m.mu_u; s.zu_order = 1;
  } else {
    s.kind = 1; s.a = m.rho_d / 2; s.c = m.mu_d2; s.zu_order = 1;
  }
  return s;
}

static inline int chunks_of(const GraphDev& g) { return (g.T * g.N + kBlock - 1) / kBlock; }

template <typename S>
static int launch_op(mga_plan* p, int op, const S* v, const S* xin, S* out, int lhs, Diag<S> dg, S c, int zu_order,
                     int64_t B, cudaStream_t st) {
  const GraphDev& g = p->g;
  const int ch = chunks_of(g);
  const unsigned grid = (unsigned)(B * ch);
  if (op == kOpLu) k_apply<S, kOpLu><<<grid, kBlock, 0, st>>>(g, ch, v, xin, out, lhs, dg, c, zu_order);
  else if (op == kOpLdr) k_apply<S, kOpLdr><<<grid, kBlock, 0, st>>>(g, ch, v, xin, out, lhs, dg, c, zu_order);
  else k_apply<S, kOpLdrT><<<grid, kBlock, 0, st>>>(g, ch, v, xin, out, lhs, dg, c, zu_order);
  MGA_LAUNCH_CHECK("k_apply");
  return MGA_OK;
}

// y = A x for a system; `tmp` is needed by cLdr systems
template <typename S>
static int apply_system(mga_plan* p, const Sys& s, const S* x, S* y, S* tmp, const S* mask, int64_t B,
                        cudaStream_t st) {
  Diag<S> dg{(S)s.a, s.hth, mask, p->g.t_in};
  Diag<S> none{(S)0, 0, nullptr, 0};
  if (s.kind == 0) return launch_op<S>(p, kOpLu, x, x, y, 1, dg, (S)s.c, s.zu_order, B, st);
  if (s.kind == 1) {
    int rc = launch_op<S>(p, kOpLdr, x, x, tmp, 0, none, (S)0, 0, B, st);
    if (rc) return rc;
    return launch_op<S>(p, kOpLdrT, tmp, x, y, 1, dg, (S)s.c, s.zu_order, B, st);
  }
  // diagonal system: reuse the Lu kernel with c = 0 would still gather; use LdrT-free path
  return launch_op<S>(p, kOpLu, x, x, y, 1, dg, (S)0, 1, B, st);
}

template <typename S>
static int apply_impl(mga_plan* p, int op, const mga_params* prm, const S* x, S* y, const S* mask, int64_t B,
                      cudaStream_t st) {
  const GraphDev& g = p->g;
  Diag<S> none{(S)0, 0, nullptr, 0};
  const size_t vec = (size_t)B * g.T * g.N * sizeof(S);
  switch (op) {
    case MGA_OP_LU: return launch_op<S>(p, kOpLu, x, x, y, 0, none, (S)0, 0, B, st);
    case MGA_OP_LDR: return launch_op<S>(p, kOpLdr, x, x, y, 0, none, (S)0, 0, B, st);
    case MGA_OP_LDRT: return launch_op<S>(p, kOpLdrT, x, x, y, 0, none, (S)0, 0, B, st);
    default: break;
  }
  int rc = ensure_workspace(p, p->ws, vec);
  if (rc) return rc;
  S* tmp = static_cast<S*>(p->ws.base);
  if (op == MGA_OP_CLDR) {
    rc = launch_op<S>(p, kOpLdr, x, x, tmp, 0, none, (S)0, 0, B, st);
    if (rc) return rc;
    return launch_op<S>(p, kOpLdrT, tmp, x, y, 0, none, (S)0, 0, B, st);
  }
  const int system = op == MGA_OP_LHS_X ? MGA_SYS_X : op == MGA_OP_LHS_ZU ? MGA_SYS_ZU : MGA_SYS_ZD;
  return apply_system<S>(p, system_of(system, *prm), x, y, tmp, system == MGA_SYS_X ? mask : nullptr, B, st);
}

// Buffers of one CG solve, carved from the plan workspace by the caller.
template <typename S>
struct CgBufs {
  S *r, *p0, *p1, *ap, *q;
  double* dots;   // (2*max_iter+1, B)
  int* flag;      // device int for the stop test
};

template <typename S>
static size_t cg_bytes(const GraphDev& g, int64_t B, int max_iter) {
  const size_t vec = (((size_t)B * g.T * g.N * sizeof(S)) + 255) & ~(size_t)255;
  return 5 * vec + (((size_t)(2 * max_iter + 1) * B * sizeof(double) + 255) & ~(size_t)255) + 256;
}

template <typename S>
static CgBufs<S> cg_carve(char* base, const GraphDev& g, int64_t B, int max_iter) {
  const size_t vec = (((size_t)B * g.T * g.N * sizeof(S)) + 255) & ~(size_t)255;
  CgBufs<S> b;
  b.r = reinterpret_cast<S*>(base);
  b.p0 = reinterpret_cast<S*>(base + vec);
  b.p1 = reinterpret_cast<S*>(base + 2 * vec);
  b.ap = reinterpret_cast<S*>(base + 3 * vec);
  b.q = reinterpret_cast<S*>(base + 4 * vec);
  b.dots = reinterpret_cast<double*>(base + 5 * vec);
  b.flag = reinterpret_cast<int*>(base + 5 * vec + ((((size_t)(2 * max_iter + 1) * B * sizeof(double)) + 255) & ~(size_t)255));
  return b;
}

// CG_solver (ADMM.py:329-368).  x holds x0 on entry and the solution on exit.
template <typename S>
static int cg_impl(mga_plan* p, const Sys& s, const S* rhs, S* x, const S* mask_first, int64_t B, int max_iter,
                   double tol, int32_t* iters_out, S* alpha, S* beta, CgBufs<S> w, cudaStream_t st) {
  const GraphDev& g = p->g;
  const int ch = chunks_of(g);
  const unsigned grid = (unsigned)(B * ch);
  Diag<S> dg{(S)s.a, s.hth, nullptr, g.t_in};   // iterations never see the mask (quirk Q4)
  MGA_CUDA(cudaMemsetAsync(w.dots, 0, (size_t)(2 * max_iter + 1) * B * sizeof(double), st));
  // r = rhs - A x0 (w.ap holds A x0, w.q the L_d intermediate)
  int rc = apply_system<S>(p, s, x, w.ap, w.q, mask_first, B, st);
  if (rc) return rc;
  k_cg_init<S><<<grid, kBlock, 0, st>>>(g, ch, rhs, w.ap, w.r, w.dots);
  MGA_LAUNCH_CHECK("k_cg_init");
  int done = -1;
  int* h_flag = static_cast<int*>(p->pinned);
  for (int k = 0; k < max_iter; ++k) {
    const S* p_old = (k & 1) ? w.p0 : w.p1;     // k == 0 never reads it
    S* p_new = (k & 1) ? w.p1 : w.p0;
    if (s.kind == 0) {
      k_cg_p1_lu<S><<<grid, kBlock, 0, st>>>(g, ch, B, k, w.r, p_old, p_new, w.ap, w.dots, dg, (S)s.c);
      MGA_LAUNCH_CHECK("k_cg_p1_lu");
    } else if (s.kind == 1) {
      k_cg_p1a<S><<<grid, kBlock, 0, st>>>(g, ch, B, k, w.r, p_old, p_new, w.q, w.dots);
      MGA_LAUNCH_CHECK("k_cg_p1a");
      k_cg_p1b<S><<<grid, kBlock, 0, st>>>(g, ch, B, k, p_new, w.q, w.ap, w.dots, dg, (S)s.c, s.zu_order);
      MGA_LAUNCH_CHECK("k_cg_p1b");
    } else {
      k_cg_p1_diag<S><<<grid, kBlock, 0, st>>>(g, ch, B, k, w.r, p_old, p_new, w.ap, w.dots, dg);
      MGA_LAUNCH_CHECK("k_cg_p1_diag");
    }
    k_cg_p2<S><<<grid, kBlock, 0, st>>>(g, ch, B, k, x, w.r, p_new, w.ap, w.dots);
    MGA_LAUNCH_CHECK("k_cg_p2");
    if (tol > 0) {
      k_cg_check<S><<<1, 1024, 0, st>>>(B, w.dots + (size_t)(2 * k + 2) * B, (S)tol, w.flag);
      MGA_LAUNCH_CHECK("k_cg_check");
      MGA_CUDA(cudaMemcpyAsync(h_flag, w.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      MGA_CUDA(cudaStreamSynchronize(st));
      if (*h_flag) { done = k + 1; break; }
    }
  }
  const int used = done > 0 ? done : max_iter;
  if ((alpha || beta) && used > 0) {
    const int64_t tot = B * used;
    k_cg_coeffs<S><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(B, used, w.dots, alpha, beta);
    MGA_LAUNCH_CHECK("k_cg_coeffs");
  }
  if (iters_out) *iters_out = done;
  return MGA_OK;
}

template <typename S>
static int cg_entry(mga_plan* p, int system, const mga_params* prm, const S* rhs, S* x, const S* mask_first,
                    int64_t B, int max_iter, double tol, int32_t* iters_out, S* alpha, S* beta, cudaStream_t st) {
  int rc = ensure_workspace(p, p->ws, cg_bytes<S>(p->g, B, max_iter));
  if (rc) return rc;
  CgBufs<S> w = cg_carve<S>(static_cast<char*>(p->ws.base), p->g, B, max_iter);
  return cg_impl<S>(p, system_of(system, *prm), rhs, x, system == MGA_SYS_X ? mask_first : nullptr, B, max_iter,
                    tol, iters_out, alpha, beta, w, st);
}

// combined_loop (ADMM.py:528-648)
template <typename S>
static int admm_impl(mga_plan* p, const mga_params* prm, const S* y, int y_rows, const S* mask, S* x_out, int64_t B,
                     int n_outer, int max_cg, double cg_tol, double admm_tol, double t_mean, double t_var,
                     int diag_flags, const mga_admm_outputs* outs, cudaStream_t st) {
  const GraphDev& g = p->g;
  const int abl = prm->ablation;
  const bool with_phi = abl == MGA_ABL_NONE || abl == MGA_ABL_DGLR;
  const bool with_zd = abl != MGA_ABL_DGLR;
  const bool want_diag = (diag_flags & 1) != 0, accumulate = (diag_flags & 2) != 0;
  const int64_t n = B * (int64_t)g.T * g.N;
  const size_t vec = (((size_t)n * sizeof(S)) + 255) & ~(size_t)255;
  // state vectors the caller did not ask for live in the workspace; x/zu/zd are ping-ponged so
  // the previous iterate stays available for the dual residuals (ADMM.py:547-550, 612-636)
  const int n_state = 11;  // x1, zu0, zu1, zd0, zd1, gu, gd, gam, phi, rhs, spare
  const size_t need = n_state * vec + cg_bytes<S>(g, B, max_cg) + 256 +
                      (want_diag && !outs->diag ? (size_t)n_outer * MGA_DIAG_COLS * sizeof(double) + 256 : 0);
  int rc = ensure_workspace(p, p->ws, need);
  if (rc) return rc;
  char* base = static_cast<char*>(p->ws.base);
  auto V = [&](int k) { return reinterpret_cast<S*>(base + (size_t)k * vec); };
  S* xa = x_out;   // buffers of the x ping-pong: the final x must land in x_out
  S* xb = V(0);
  S *zu0 = V(1), *zu1 = V(2), *zd0 = V(3), *zd1 = V(4);
  S *gu = V(5), *gd = V(6), *gam = V(7), *phi = V(8), *rhs = V(9);
  CgBufs<S> w = cg_carve<S>(base + n_state * vec, g, B, max_cg);
  double* diag = outs->diag;
  if (want_diag && !diag) diag = reinterpret_cast<double*>(base + n_state * vec + cg_bytes<S>(g, B, max_cg));
  double* dx_sum = want_diag ? outs->dx_sum : nullptr;
  if (want_diag && !accumulate) {
    MGA_CUDA(cudaMemsetAsync(diag, 0, (size_t)n_outer * MGA_DIAG_COLS * sizeof(double), st));
    if (dx_sum) MGA_CUDA(cudaMemsetAsync(dx_sum, 0, (size_t)n_outer * g.T * g.N * sizeof(double), st));
  }
  // non-finite detection needs a diag row even when diagnostics are off: use a 1-row scratch
  double* nf_row = nullptr;
  if (!want_diag) {
    nf_row = reinterpret_cast<double*>(w.flag + 8);
    MGA_CUDA(cudaMemsetAsync(nf_row, 0, MGA_DIAG_COLS * sizeof(double), st));
  }

  const int ch = chunks_of(g);
  const unsigned grid = (unsigned)(B * ch);
  const unsigned grid_bn = (unsigned)((B * g.N + kBlock - 1) / kBlock);
  const unsigned grid_n = (unsigned)((n + kBlock - 1) / kBlock);
  // the x ping-pong must end in x_out: with an odd number of solves start in xb
  S* x_cur = (n_outer % 2 == 0) ? xa : xb;
  S* x_nxt = (n_outer % 2 == 0) ? xb : xa;
  if (n_outer == 0) x_cur = xa;
  if (!mask) {
    k_init<S><<<grid_bn, kBlock, 0, st>>>(g, B, y, x_cur, zu0, zd0, gu, gd, with_phi ? gam : nullptr, (float)t_mean,
                                         (float)t_var);
    MGA_LAUNCH_CHECK("k_init");
  } else {
    k_interp<S><<<grid_bn, kBlock, 0, st>>>(g, B, y, mask, x_cur);
    MGA_LAUNCH_CHECK("k_interp");
    k_init_from_x<S><<<grid_n, kBlock, 0, st>>>(n, x_cur, zu0, zd0, gu, gd, with_phi ? gam : nullptr);
    MGA_LAUNCH_CHECK("k_init_from_x");
  }
  Diag<S> none{(S)0, 0, nullptr, 0};
  if (with_phi) {
    rc = launch_op<S>(p, kOpLdr, x_cur, x_cur, phi, 0, none, (S)0, 0, B, st);   // ADMM.py:541
    if (rc) return rc;
  }
  S *zu_cur = zu0, *zu_nxt = zu1, *zd_cur = zd0, *zd_nxt = zd1;
  const Sys sx = system_of(MGA_SYS_X, *prm), szu = system_of(MGA_SYS_ZU, *prm), szd = system_of(MGA_SYS_ZD, *prm);
  const S rho = (S)prm->rho, rho_u = (S)prm->rho_u, rho_d = (S)prm->rho_d;
  const S thr = (S)(prm->mu_d1 / prm->rho);
  const size_t coef_stride = (size_t)max_cg * B;
  int outer_done = 0;
  const size_t vbytes = (size_t)n * sizeof(S);
  for (int it = 0; it < n_outer; ++it) {
    int32_t iters[3] = {-1, -1, -1};
    k_rhs_x<S><<<grid, kBlock, 0, st>>>(g, ch, abl, gam, phi, zu_cur, zd_cur, gu, gd, y, y_rows, rhs, rho, rho_u, rho_d);
    MGA_LAUNCH_CHECK("k_rhs_x");
    // x solve, warm start x_old (ADMM.py:571)
    MGA_CUDA(cudaMemcpyAsync(x_nxt, x_cur, vbytes, cudaMemcpyDeviceToDevice, st));
    S* al = outs->alpha ? static_cast<S*>(outs->alpha) + ((size_t)it * 3 + 0) * coef_stride : nullptr;
    S* be = outs->beta ? static_cast<S*>(outs->beta) + ((size_t)it * 3 + 0) * coef_stride : nullptr;
    rc = cg_impl<S>(p, sx, rhs, x_nxt, mask, B, max_cg, cg_tol, &iters[0], al, be, w, st);
    if (rc) return rc;
    // z_u solve (ADMM.py:579-580)
    k_rhs_z<S><<<grid_n, kBlock, 0, st>>>(n, gu, x_nxt, rhs, (S)(prm->rho_u / 2));
    MGA_LAUNCH_CHECK("k_rhs_z");
    MGA_CUDA(cudaMemcpyAsync(zu_nxt, zu_cur, vbytes, cudaMemcpyDeviceToDevice, st));
    al = outs->alpha ? static_cast<S*>(outs->alpha) + ((size_t)it * 3 + 1) * coef_stride : nullptr;
    be = outs->beta ? static_cast<S*>(outs->beta) + ((size_t)it * 3 + 1) * coef_stride : nullptr;
    rc = cg_impl<S>(p, szu, rhs, zu_nxt, nullptr, B, max_cg, cg_tol, &iters[1], al, be, w, st);
    if (rc) return rc;
    if (with_zd) {   // ADMM.py:586-588
      k_rhs_z<S><<<grid_n, kBlock, 0, st>>>(n, gd, x_nxt, rhs, (S)(prm->rho_d / 2));
      MGA_LAUNCH_CHECK("k_rhs_z");
      MGA_CUDA(cudaMemcpyAsync(zd_nxt, zd_cur, vbytes, cudaMemcpyDeviceToDevice, st));
      al = outs->alpha ? static_cast<S*>(outs->alpha) + ((size_t)it * 3 + 2) * coef_stride : nullptr;
      be = outs->beta ? static_cast<S*>(outs->beta) + ((size_t)it * 3 + 2) * coef_stride : nullptr;
      rc = cg_impl<S>(p, szd, rhs, zd_nxt, nullptr, B, max_cg, cg_tol, &iters[2], al, be, w, st);
      if (rc) return rc;
    }
    double* drow = want_diag ? diag + (size_t)it * MGA_DIAG_COLS : nf_row;
    k_tail<S><<<grid, kBlock, 0, st>>>(g, ch, abl, want_diag ? 1 : 0, x_nxt, x_cur, zu_nxt, zu_cur,
                                       with_zd ? zd_nxt : zd_cur, zd_cur, gu, gd, gam, phi, y, y_rows, mask, rho,
                                       rho_u, rho_d, thr, drow,
                                       dx_sum ? dx_sum + (size_t)it * g.T * g.N : nullptr);
    MGA_LAUNCH_CHECK("k_tail");
    std::swap(x_cur, x_nxt);
    std::swap(zu_cur, zu_nxt);
    if (with_zd) std::swap(zd_cur, zd_nxt);
    if (outs->cg_iters) { outs->cg_iters[it * 3 + 0] = iters[0]; outs->cg_iters[it * 3 + 1] = iters[1]; outs->cg_iters[it * 3 + 2] = iters[2]; }
    outer_done = it + 1;
    if (admm_tol > 0 && want_diag) {   // stop test (ADMM.py:645): whole-batch norms, host sync
      double* h = static_cast<double*>(p->pinned) + 8;
      MGA_CUDA(cudaMemcpyAsync(h, drow, MGA_DIAG_COLS * sizeof(double), cudaMemcpyDeviceToHost, st));
      MGA_CUDA(cudaStreamSynchronize(st));
      auto nrm = [&](int c) { return (double)(S)std::sqrt((S)h[c]); };
      double pri = nrm(MGA_DIAG_X_ZU2), dua = nrm(MGA_DIAG_DZU2);
      if (with_phi) { pri = std::max(pri, nrm(MGA_DIAG_PHI_LDX2)); dua = std::max(dua, nrm(MGA_DIAG_DPHI2)); }
      if (with_zd) { pri = std::max(pri, nrm(MGA_DIAG_X_ZD2)); dua = std::max(dua, nrm(MGA_DIAG_DZD2)); }
      if (pri < admm_tol && dua < admm_tol) break;
    }
  }
  if (x_cur != x_out) MGA_CUDA(cudaMemcpyAsync(x_out, x_cur, vbytes, cudaMemcpyDeviceToDevice, st));
  auto give = [&](void* dst, const S* src) -> int {
    if (dst) MGA_CUDA(cudaMemcpyAsync(dst, src, vbytes, cudaMemcpyDeviceToDevice, st));
    return MGA_OK;
  };
  if ((rc = give(outs->zu, zu_cur))) return rc;
  if (with_zd && (rc = give(outs->zd, zd_cur))) return rc;
  if ((rc = give(outs->gamma_u, gu))) return rc;
  if (with_zd && (rc = give(outs->gamma_d, gd))) return rc;
  if (with_phi && (rc = give(outs->phi, phi))) return rc;
  if (with_phi && (rc = give(outs->gamma, gam))) return rc;
  if (outs->outer_done) *outs->outer_done = outer_done;
  return MGA_OK;
}

// ---------------------------------------------------------------------------------------------
int stream_apply(mga_plan* p, int op, const mga_params* prm, const void* x, void* y, const void* mask, int64_t B,
                 int dtype, cudaStream_t st) {
  if (dtype == MGA_F32)
    return apply_impl<float>(p, op, prm, static_cast<const float*>(x), static_cast<float*>(y),
                             static_cast<const float*>(mask), B, st);
  return apply_impl<double>(p, op, prm, static_cast<const double*>(x), static_cast<double*>(y),
                            static_cast<const double*>(mask), B, st);
}

int stream_cg(mga_plan* p, int system, const mga_params* prm, const void* rhs, void* x, const void* mask_first,
              int64_t B, int dtype, int max_iter, double tol, int32_t* iters_out, void* alpha, void* beta,
              cudaStream_t st) {
  if (dtype == MGA_F32)
    return cg_entry<float>(p, system, prm, static_cast<const float*>(rhs), static_cast<float*>(x),
                           static_cast<const float*>(mask_first), B, max_iter, tol, iters_out,
                           static_cast<float*>(alpha), static_cast<float*>(beta), st);
  return cg_entry<double>(p, system, prm, static_cast<const double*>(rhs), static_cast<double*>(x),
                          static_cast<const double*>(mask_first), B, max_iter, tol, iters_out,
                          static_cast<double*>(alpha), static_cast<double*>(beta), st);
}

int stream_admm(mga_plan* p, const mga_params* prm, const void* y, int y_rows, const void* mask, void* x_out,
                int64_t B, int dtype, int n_outer, int max_cg, double cg_tol, double admm_tol, double t_mean,
                double t_var, int want_diag, const mga_admm_outputs* outs, cudaStream_t st) {
  if (dtype == MGA_F32)
    return admm_impl<float>(p, prm, static_cast<const float*>(y), y_rows, static_cast<const float*>(mask),
                            static_cast<float*>(x_out), B, n_outer, max_cg, cg_tol, admm_tol, t_mean, t_var,
                            want_diag, outs, st);
  return admm_impl<double>(p, prm, static_cast<const double*>(y), y_rows, static_cast<const double*>(mask),
                           static_cast<double*>(x_out), B, n_outer, max_cg, cg_tol, admm_tol, t_mean, t_var,
                           want_diag, outs, st);
}

}  // namespace mga

using namespace mga;

// ---- small C-ABI entry points that map one-to-one on a kernel --------------------------------
extern "C" {

#define MGA_DISPATCH(S_EXPR)            \
  if (dtype == MGA_F32) { using S = float; S_EXPR; } else { using S = double; S_EXPR; }

static int pre(mga_plan* p, int64_t B, int dtype, const char* who) {
  if (!p || B <= 0 || (dtype != MGA_F32 && dtype != MGA_F64)) { set_error(std::string(who) + ": bad argument"); return MGA_ERR_INVALID; }
  cudaError_t e = cudaSetDevice(p->device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  return MGA_OK;
}

int mga_initial_guess(mga_plan* p, const void* y, void* x, int64_t B, int dtype, double t_mean, double t_var,
                      void* stream) {
  int rc = pre(p, B, dtype, "mga_initial_guess");
  if (rc) return rc;
  if (!y || !x) { set_error("mga_initial_guess: NULL buffer"); return MGA_ERR_INVALID; }
  const unsigned grid = (unsigned)((B * p->g.N + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)stream;
  MGA_DISPATCH((k_init<S><<<grid, kBlock, 0, st>>>(p->g, B, static_cast<const S*>(y), static_cast<S*>(x), nullptr,
                                                   nullptr, nullptr, nullptr, nullptr, (float)t_mean, (float)t_var)));
  MGA_LAUNCH_CHECK("k_init");
  return MGA_OK;
}

int mga_rhs_x(mga_plan* p, const mga_params* prm, const void* gamma, const void* phi, const void* zu, const void* zd,
              const void* gamma_u, const void* gamma_d, const void* y, int y_rows, void* rhs, int64_t B, int dtype,
              void* stream) {
  int rc = pre(p, B, dtype, "mga_rhs_x");
  if (rc) return rc;
  if (!prm || !zu || !gamma_u || !y || !rhs) { set_error("mga_rhs_x: NULL buffer"); return MGA_ERR_INVALID; }
  const int ch = chunks_of(p->g);
  const unsigned grid = (unsigned)(B * ch);
  cudaStream_t st = (cudaStream_t)stream;
  MGA_DISPATCH((k_rhs_x<S><<<grid, kBlock, 0, st>>>(
      p->g, ch, prm->ablation, static_cast<const S*>(gamma), static_cast<const S*>(phi), static_cast<const S*>(zu),
      static_cast<const S*>(zd), static_cast<const S*>(gamma_u), static_cast<const S*>(gamma_d),
      static_cast<const S*>(y), y_rows, static_cast<S*>(rhs), (S)prm->rho, (S)prm->rho_u, (S)prm->rho_d)));
  MGA_LAUNCH_CHECK("k_rhs_x");
  return MGA_OK;
}

int mga_dual_ascent(mga_plan* p, double rho_z, const void* x, const void* z, void* gz, int64_t B, int dtype,
                    void* stream) {
  int rc = pre(p, B, dtype, "mga_dual_ascent");
  if (rc) return rc;
  if (!x || !z || !gz) { set_error("mga_dual_ascent: NULL buffer"); return MGA_ERR_INVALID; }
  const int64_t n = B * (int64_t)p->g.T * p->g.N;
  const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
  cudaStream_t st = (cudaStream_t)stream;
  MGA_DISPATCH((k_dual<S><<<grid, kBlock, 0, st>>>(n, (S)rho_z, static_cast<const S*>(x), static_cast<const S*>(z),
                                                   static_cast<S*>(gz))));
  MGA_LAUNCH_CHECK("k_dual");
  return MGA_OK;
}

static int phi_common(mga_plan* p, const mga_params* prm, const void* x, const void* gin, void* gout, void* phi,
                      int64_t B, int dtype, void* stream, const char* who) {
  int rc = pre(p, B, dtype, who);
  if (rc) return rc;
  if (!prm || !x || !gin || !phi) { set_error(std::string(who) + ": NULL buffer"); return MGA_ERR_INVALID; }
  const int ch = chunks_of(p->g);
  const unsigned grid = (unsigned)(B * ch);
  cudaStream_t st = (cudaStream_t)stream;
  MGA_DISPATCH((k_phi<S><<<grid, kBlock, 0, st>>>(p->g, ch, static_cast<const S*>(x), static_cast<const S*>(gin),
                                                  static_cast<S*>(gout), static_cast<S*>(phi), (S)prm->rho,
                                                  (S)(prm->mu_d1 / prm->rho))));
  MGA_LAUNCH_CHECK("k_phi");
  return MGA_OK;
}

int mga_prox_phi_dual(mga_plan* p, const mga_params* prm, const void* x, void* gamma_inout, void* phi_out, int64_t B,
                      int dtype, void* stream) {
  return phi_common(p, prm, x, gamma_inout, gamma_inout, phi_out, B, dtype, stream, "mga_prox_phi_dual");
}

int mga_phi_direct(mga_plan* p, const mga_params* prm, const void* x, const void* gamma, void* phi_out, int64_t B,
                   int dtype, void* stream) {
  return phi_common(p, prm, x, gamma, nullptr, phi_out, B, dtype, stream, "mga_phi_direct");
}

}  // extern "C"
