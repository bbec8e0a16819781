// One (CH, K) instantiation of the resident kernel per translation unit, so the variants
// compile in parallel (build.py passes -DMGA_CH=.. -DMGA_K=..).
#include "mga_resident.cuh"

#ifndef MGA_CH
#error "compile with -DMGA_CH=<1|2|3> -DMGA_K=<4|6|8|10>"
#endif

#define MGA_CAT_(a, b, c) a##b##_##c
#define MGA_CAT(a, b, c) MGA_CAT_(a, b, c)

namespace mga {
int MGA_CAT(resident_launch_, MGA_CH, MGA_K)(mga_plan* p, ResArgs& a, const ResGeom& geo, cudaStream_t st) {
  return pick_threads<MGA_CH, MGA_K>(p, a, geo, st);
}
int MGA_CAT(resident_cg_launch_, MGA_CH, MGA_K)(mga_plan* p, ResArgs& a, const CgArgs& g, const ResGeom& geo, cudaStream_t st) {
  return pick_threads_cg<MGA_CH, MGA_K>(p, a, g, geo, st);
}
}  // namespace mga
