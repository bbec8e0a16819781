// Host-side schedule of the resident kernel's gathers (see mga_schedule.cpp).
#pragma once
#include <vector>

namespace mga {

struct ResidentSchedule {
  int N = 0;
  int kd = 0, ku = 0;         // slots per row AFTER self links and "-1" pads were dropped
  std::vector<float> w_self;  // (N) internal order: weight of the node's link to itself in the temporal table
  int overload = 0;           // gathers left over the conflict-free bound after bank-class balancing (0 = none)
  std::vector<int> perm;      // perm[internal] = original node
  std::vector<int> inv;       // inv[original] = internal
  std::vector<int> nbr_d;     // (N, kd) internal ids; N .. N+7 = zero rows (zero-weight padding, one per bank class)
  std::vector<float> w_d;
  std::vector<int> nbr_u;     // (N, ku)
  std::vector<float> w_u;
  std::vector<int> ell_ptr;   // (n_warps + 1) first step of each 32-row warp
  std::vector<int> ell_node;  // (steps_total * 32) internal ids, N .. N+7 = padding
  std::vector<float> ell_w;
};

// order[new] = old: reverse Cuthill-McKee on the symmetrised union of the two neighbour tables (-1 = none)
std::vector<int> graph_rcm_order(int N, int kd, const int* nbr_d, int ku, const int* nbr_u);

void build_resident_schedule(int N, int kd, const int* nbr_d, const float* d_w, int ku, const int* nbr_u,
                             const float* u_w, const int* csr_ptr, const int* csr_src, const float* csr_w,
                             ResidentSchedule* out);

}  // namespace mga
