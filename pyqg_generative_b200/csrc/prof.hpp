// prof.hpp -- device-side timing of individual kernels inside the running step loop (qgb_profile_* entry points): launches of
// the selected slot(s) are bracketed by CUDA events on the launching stream.  Events come from a pool that is created on first
// use and reused, so a profiled loop does not create events after its first pass.
#pragma once
#include <cuda_runtime.h>

#include <vector>

namespace qgb {

// slots 0-7: conv layers of network 0, 8-15: of network 1, then the non-CNN kernels of a step
enum { PROF_SPECTRAL = 16, PROF_LATENT = 17, PROF_FINISH = 18, PROF_DIAG = 19, PROF_SLOTS = 20 };

struct Profiler {
  int net = -1, layer = -1;     // (net, layer >= 0): that conv layer only;  layer == -2: every slot;  -1: off
  struct Rec { int slot; cudaEvent_t a, b; long long units; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  bool wants(int slot) const { return layer == -2 || (layer >= 0 && slot == net * 8 + layer); }
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e = nullptr; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
  int start(int slot, cudaStream_t st) {
    if (!wants(slot)) return -1;
    Rec r{slot, get(), get(), 0};
    cudaEventRecord(r.a, st);
    recs.push_back(r);
    return (int)recs.size() - 1;
  }
  void stop(int idx, cudaStream_t st, long long units) {
    if (idx < 0) return;
    cudaEventRecord(recs[idx].b, st);
    recs[idx].units = units;
  }
  void reset() { recs.clear(); used = 0; net = -1; layer = -1; }
  void destroy() { for (auto e : pool) cudaEventDestroy(e); pool.clear(); reset(); }
};

}  // namespace qgb
