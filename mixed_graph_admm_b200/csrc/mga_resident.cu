// Resident-mode dispatch: eligibility, argument block, choice of the (TT, K) instantiation.
// The kernel itself is in mga_resident.cuh.
#include <algorithm>

#include "mga_resident.cuh"

namespace mga {

#define MGA_DECL(TT, K) int resident_launch_##TT##_##K(mga_plan*, ResArgs&, int, cudaStream_t);
MGA_DECL(12, 5) MGA_DECL(12, 7) MGA_DECL(12, 9) MGA_DECL(24, 5) MGA_DECL(24, 7) MGA_DECL(24, 9)
#undef MGA_DECL

constexpr int kResMaxN = 512;
constexpr int kResMaxT = 24;
constexpr int kResMaxK = 9;

bool resident_eligible(const mga_plan* p, int dtype) {
  const GraphDev& g = p->g;
  if (dtype != MGA_F32) return false;
  if (g.temporal == MGA_TEMPORAL_BAND) return false;
  if (g.N > kResMaxN || g.T > kResMaxT) return false;
  if (g.u_wT != 1 || g.d_wT != 1) return false;
  if (std::max(g.kd, g.ku + 1) > kResMaxK) return false;
  int th = 0;
  return resident_smem_bytes(p, &th) <= p->max_smem_optin;
}

int resident_smem_bytes(const mga_plan* p, int* threads) {
  const GraphDev& g = p->g;
  const int NP = ((g.N + 1 + 31) / 32) * 32;
  if (threads) *threads = ((g.N + 31) / 32) * 32;
  return (int)res_core_bytes(g, NP);
}

static int pick(mga_plan* p, ResArgs& a, int threads, cudaStream_t st) {
  const int k = std::max(p->g.kd, p->g.ku + 1);
  const bool small_t = p->g.T <= 12;
  if (k <= 5) return small_t ? resident_launch_12_5(p, a, threads, st) : resident_launch_24_5(p, a, threads, st);
  if (k <= 7) return small_t ? resident_launch_12_7(p, a, threads, st) : resident_launch_24_7(p, a, threads, st);
  return small_t ? resident_launch_12_9(p, a, threads, st) : resident_launch_24_9(p, a, threads, st);
}

int resident_admm(mga_plan* p, const mga_params* m, const void* y, void* x_out, int64_t B, int n_outer, int n_cg,
                  double t_mean, double t_var, int diag_flags, const mga_admm_outputs* outs, cudaStream_t st) {
  const GraphDev& g = p->g;
  ResArgs a{};
  a.N = g.N; a.T = g.T; a.t_in = g.t_in; a.n_outer = n_outer; a.n_cg = n_cg; a.q1 = g.q1; a.nnz = g.nnz;
  a.NP = ((g.N + 1 + 31) / 32) * 32;
  a.B = B; a.kd = g.kd; a.ku = g.ku;
  a.nbr_d = g.nbr_d; a.d_w = g.d_w; a.nbr_u = g.nbr_u; a.u_w = g.u_w;
  a.csr_ptr = g.csr_ptr; a.csr_src = g.csr_src; a.csr_w = g.csr_w;
  a.y = static_cast<const float*>(y);
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
  if (a.want_diag && !(diag_flags & 2)) {
    if (a.diag) MGA_CUDA(cudaMemsetAsync(a.diag, 0, (size_t)n_outer * MGA_DIAG_COLS * sizeof(double), st));
    if (a.dx_sum) MGA_CUDA(cudaMemsetAsync(a.dx_sum, 0, (size_t)n_outer * g.T * g.N * sizeof(double), st));
  }
  const int threads = ((g.N + 31) / 32) * 32;
  if (outs->cg_iters) for (int k = 0; k < n_outer * 3; ++k) outs->cg_iters[k] = -1;
  if (outs->outer_done) *outs->outer_done = n_outer;
  return pick(p, a, threads, st);
}

}  // namespace mga
