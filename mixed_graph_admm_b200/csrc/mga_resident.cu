// Resident-mode dispatch: eligibility, argument block, choice of the (TT, K) instantiation.
// The kernel itself is in mga_resident.cuh.
#include <algorithm>

#include "mga_resident.cuh"

namespace mga {

#define MGA_DECL(CH, K) int resident_launch_##CH##_##K(mga_plan*, ResArgs&, const ResGeom&, cudaStream_t);
MGA_DECL(1, 4) MGA_DECL(1, 6) MGA_DECL(1, 8) MGA_DECL(1, 10) MGA_DECL(2, 4) MGA_DECL(2, 6) MGA_DECL(2, 8) MGA_DECL(2, 10)
MGA_DECL(3, 4) MGA_DECL(3, 6) MGA_DECL(3, 8) MGA_DECL(3, 10)
#undef MGA_DECL
#define MGA_DECL(CH, K) int resident_cg_launch_##CH##_##K(mga_plan*, ResArgs&, const CgArgs&, const ResGeom&, cudaStream_t);
MGA_DECL(1, 4) MGA_DECL(1, 6) MGA_DECL(1, 8) MGA_DECL(1, 10) MGA_DECL(2, 4) MGA_DECL(2, 6) MGA_DECL(2, 8) MGA_DECL(2, 10)
MGA_DECL(3, 4) MGA_DECL(3, 6) MGA_DECL(3, 8) MGA_DECL(3, 10)
#undef MGA_DECL

constexpr int kResMaxT = 24;
constexpr int kResMaxK = 10;    // slots per forward table after self links and pads are dropped

static bool geometry(const mga_plan* p, ResGeom* geo) {
  const GraphDev& g = p->g;
  if (!p->has_sched) return false;
  if (!res_geometry(g, p->r_kd, p->r_ku, p->r_ell_total, res_forced_ch(), geo) &&
      !res_geometry(g, p->r_kd, p->r_ku, p->r_ell_total, 0, geo))
    return false;
  return geo->core_bytes <= (size_t)p->max_smem_optin;
}

bool resident_eligible(const mga_plan* p, int dtype) {
  const GraphDev& g = p->g;
  if (dtype != MGA_F32) return false;
  if (g.T > kResMaxT) return false;
  if (g.u_wT != 1 || g.d_wT != 1) return false;
  if (!p->has_sched || std::max(p->r_kd, p->r_ku) > kResMaxK) return false;
  ResGeom geo;
  return geometry(p, &geo);
}

int resident_smem_bytes(const mga_plan* p, int* threads) {
  ResGeom geo;
  if (!geometry(p, &geo)) { if (threads) *threads = 0; return 0; }
  if (threads) *threads = geo.threads;
  return (int)geo.core_bytes;
}

static int pick(mga_plan* p, ResArgs& a, const ResGeom& geo, cudaStream_t st) {
  const int kk = geo.Kt;
#define MGA_CASE(C_, K_) if (geo.CH == C_ && kk == K_) return resident_launch_##C_##_##K_(p, a, geo, st);
  MGA_CASE(1, 4) MGA_CASE(1, 6) MGA_CASE(1, 8) MGA_CASE(1, 10) MGA_CASE(2, 4) MGA_CASE(2, 6) MGA_CASE(2, 8) MGA_CASE(2, 10)
  MGA_CASE(3, 4) MGA_CASE(3, 6) MGA_CASE(3, 8) MGA_CASE(3, 10)
#undef MGA_CASE
  set_error("resident: no instantiation for this (CH, K)");
  return MGA_ERR_UNSUPPORTED;
}

static void fill_tables(const mga_plan* p, const ResGeom& geo, ResArgs& a) {
  const GraphDev& g = p->g;
  a.N = g.N; a.T = g.T; a.t_in = g.t_in; a.q1 = g.q1; a.nnz = g.nnz;
  a.NT = geo.NT; a.S = geo.S; a.TP = geo.TP;
  a.kd = p->r_kd; a.ku = p->r_ku;
  a.transpose_exact = p->ldrt_gather ? 0 : 1;
  a.w_self = p->r_w_self;
  a.perm = p->r_perm; a.nbr_d = p->r_nbr_d; a.d_w = p->r_w_d; a.nbr_u = p->r_nbr_u; a.u_w = p->r_w_u;
  a.ell_ptr = p->r_ell_ptr; a.ell_ent = reinterpret_cast<const int2*>(p->r_ell_ent); a.ell_total = p->r_ell_total;
}

// CG_solver for one system, fixed iteration count, x_inout holds x0 / the solution (ADMM.py:329-368)
int resident_cg(mga_plan* p, int system, const mga_params* m, const void* rhs, void* x, int64_t B, int n_cg,
                void* alpha, void* beta, cudaStream_t st) {
  ResGeom geo;
  if (!geometry(p, &geo)) { set_error("resident: shape does not fit"); return MGA_ERR_UNSUPPORTED; }
  ResArgs a{};
  fill_tables(p, geo, a);
  a.B = B; a.n_cg = n_cg;
  CgArgs g{};
  g.system = system; g.n_cg = n_cg;
  g.rhs = static_cast<const float*>(rhs);
  g.x = static_cast<float*>(x);
  g.alpha = static_cast<float*>(alpha);
  g.beta = static_cast<float*>(beta);
  if (g.alpha && !g.beta) g.alpha = nullptr;
  if (system == MGA_SYS_X) { g.a = (float)((m->rho_u + m->rho_d) / 2); g.c = (float)(m->rho / 2); }
  else if (system == MGA_SYS_ZU) { g.a = (float)(m->rho_u / 2); g.c = (float)m->mu_u; }
  else { g.a = (float)(m->rho_d / 2); g.c = (float)m->mu_d2; }
  const int kk = geo.Kt;
#define MGA_CASE(C_, K_) if (geo.CH == C_ && kk == K_) return resident_cg_launch_##C_##_##K_(p, a, g, geo, st);
  MGA_CASE(1, 4) MGA_CASE(1, 6) MGA_CASE(1, 8) MGA_CASE(1, 10) MGA_CASE(2, 4) MGA_CASE(2, 6) MGA_CASE(2, 8) MGA_CASE(2, 10)
  MGA_CASE(3, 4) MGA_CASE(3, 6) MGA_CASE(3, 8) MGA_CASE(3, 10)
#undef MGA_CASE
  set_error("resident: no instantiation for this (CH, K)");
  return MGA_ERR_UNSUPPORTED;
}

int resident_admm(mga_plan* p, const mga_params* m, const void* y, const void* mask, void* x_out, int64_t B, int n_outer, int n_cg,
                  double t_mean, double t_var, int diag_flags, const mga_admm_outputs* outs, cudaStream_t st) {
  const GraphDev& g = p->g;
  ResArgs a{};
  ResGeom geo;
  if (!geometry(p, &geo)) { set_error("resident: shape does not fit"); return MGA_ERR_UNSUPPORTED; }
  fill_tables(p, geo, a);
  a.n_outer = n_outer; a.n_cg = n_cg;
  a.B = B;
  a.y = static_cast<const float*>(y);
  a.mask = static_cast<const float*>(mask);
  a.band_w = g.temporal == MGA_TEMPORAL_BAND ? g.band_w : nullptr;
  a.band_uniform = g.temporal == MGA_TEMPORAL_BAND ? p->band_uniform : nullptr;
  a.skip = g.skip;
  a.band_floats = g.temporal == MGA_TEMPORAL_BAND ? (g.T * g.skip + 3) / 4 * 4 : 0;
  if (a.band_w && a.mask) { set_error("resident: mask mode on the banded line graph is not built"); return MGA_ERR_UNSUPPORTED; }
  a.x_out = static_cast<float*>(x_out);
  a.out[ST_ZU] = static_cast<float*>(outs->zu);
  a.out[ST_ZD] = static_cast<float*>(outs->zd);
  a.out[ST_GU] = static_cast<float*>(outs->gamma_u);
  a.out[ST_GD] = static_cast<float*>(outs->gamma_d);
  a.out[ST_GAM] = static_cast<float*>(outs->gamma);
  a.out[ST_PHI] = static_cast<float*>(outs->phi);
  a.want_diag = (diag_flags & 1) ? 1 : 0;
  a.diag = a.want_diag ? outs->diag : nullptr;
  a.dx_sum = a.want_diag ? outs->dx_sum : nullptr;
  a.alpha = static_cast<float*>(outs->alpha);
  a.beta = static_cast<float*>(outs->beta);
  if (a.alpha && !a.beta) a.alpha = nullptr;
  a.rho = (float)m->rho; a.rho_u = (float)m->rho_u; a.rho_d = (float)m->rho_d;
  a.thr = (float)(m->mu_d1 / m->rho);
  a.ax = (float)((m->rho_u + m->rho_d) / 2); a.cx = (float)(m->rho / 2);
  a.azu = (float)(m->rho_u / 2); a.czu = (float)m->mu_u;
  a.azd = (float)(m->rho_d / 2); a.czd = (float)m->mu_d2;
  a.t_mean = (float)t_mean; a.t_var = (float)t_var;
  if (p->pipe) {
    a.ready = p->pipe->ready; a.done = p->pipe->done; a.host_done = p->pipe->host_done; a.abort_flag = p->pipe->abort_flag;
    a.chunk = p->pipe->chunk; a.chunk_up = p->pipe->chunk_up; a.epoch = p->pipe->epoch;
  }
  if (a.want_diag && !(diag_flags & 2)) {
    if (a.diag) MGA_CUDA(cudaMemsetAsync(a.diag, 0, (size_t)n_outer * MGA_DIAG_COLS * sizeof(double), st));
    if (a.dx_sum) MGA_CUDA(cudaMemsetAsync(a.dx_sum, 0, (size_t)n_outer * g.T * g.N * sizeof(double), st));
  }
  if (outs->cg_iters) for (int k = 0; k < n_outer * 3; ++k) outs->cg_iters[k] = -1;
  if (outs->outer_done) *outs->outer_done = n_outer;
  return pick(p, a, geo, st);
}

}  // namespace mga
