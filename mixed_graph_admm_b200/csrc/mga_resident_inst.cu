// One (TT, K) instantiation of the resident kernel per translation unit, so the variants
// compile in parallel (build.py passes -DMGA_TT=.. -DMGA_K=..).
#include "mga_resident.cuh"

#ifndef MGA_TT
#error "compile with -DMGA_TT=<12|24> -DMGA_K=<5|7|9>"
#endif

#define MGA_CAT_(a, b, c) a##b##_##c
#define MGA_CAT(a, b, c) MGA_CAT_(a, b, c)

namespace mga {
int MGA_CAT(resident_launch_, MGA_TT, MGA_K)(mga_plan* p, ResArgs& a, int threads, cudaStream_t st) {
  return pick_threads<MGA_TT, MGA_K>(p, a, threads, st);
}
}  // namespace mga
