// Host-side memory check of plan creation (no GPU): mga_plan.cu / mga_schedule.cpp / mga_knn.cpp are compiled with
// AddressSanitizer and linked with the other objects of libmga; the few CUDA runtime calls mga_plan_create makes are
// answered here by host stand-ins (the executable's symbols win over the shared libcudart), so every table the plan
// builds - resident schedule, RCM tables, time-tiled tables, row orders - is built and "uploaded" (copied) under ASAN.
// Descriptors come from profiles/asan_plan_dump.py (the graphs of the fuzz tests).  Built and run by profiles/asan_plan.sh.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mga.h"

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static void fill(cudaDeviceProp* p) {
  std::memset(p, 0, sizeof(*p));
  p->multiProcessorCount = 148; p->sharedMemPerBlockOptin = 232448; p->sharedMemPerMultiprocessor = 233472;
  p->l2CacheSize = 126 << 20; p->major = 10; p->minor = 0;
}
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) { fill(p); return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return cudaSuccess; }
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n ? n : 1); return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::malloc(n ? n : 1); return cudaSuccess; }
cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
}

int main(int argc, char** argv) {
  int bad = 0, done = 0;
  for (int a = 1; a < argc; ++a) {
    FILE* f = std::fopen(argv[a], "rb");
    if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[a]); return 2; }
    int32_t h[9];      // n_nodes, T, t_in, ku, u_w_T, kd, d_w_T, ldrt_mode, temporal
    int64_t n[4];      // element counts: nbr_u, u_w, nbr_d, d_w
    if (std::fread(h, sizeof h, 1, f) != 1 || std::fread(n, sizeof n, 1, f) != 1) return 2;
    std::vector<int64_t> nbr_u(n[0]), nbr_d(n[2]);
    std::vector<float> u_w(n[1]), d_w(n[3]);
    if ((n[0] && std::fread(nbr_u.data(), 8, n[0], f) != (size_t)n[0]) || (n[1] && std::fread(u_w.data(), 4, n[1], f) != (size_t)n[1]) ||
        (n[2] && std::fread(nbr_d.data(), 8, n[2], f) != (size_t)n[2]) || (n[3] && std::fread(d_w.data(), 4, n[3], f) != (size_t)n[3]))
      return 2;
    std::fclose(f);
    mga_graph_desc d{};
    d.n_nodes = h[0]; d.T = h[1]; d.t_in = h[2]; d.ku = h[3]; d.u_w_T = h[4]; d.kd = h[5]; d.d_w_T = h[6]; d.ldrt_mode = h[7]; d.temporal = h[8];
    d.nbr_u = n[0] ? nbr_u.data() : nullptr; d.u_w = n[1] ? u_w.data() : nullptr;
    d.nbr_d = n[2] ? nbr_d.data() : nullptr; d.d_w = n[3] ? d_w.data() : nullptr;
    mga_plan* p = nullptr;
    const int rc = mga_plan_create(&d, 0, &p);
    if (rc) { std::fprintf(stderr, "%s: mga_plan_create rc %d: %s\n", argv[a], rc, mga_last_error()); ++bad; }
    if (p) mga_plan_destroy(p);
    ++done;
  }
  std::printf("%d plans created and destroyed, %d refused\n", done, bad);
  return 0;
}
