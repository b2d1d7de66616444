// Shared host/device helpers for the floodsr_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/floodsr_b200.h"
#include "fsr_plan.h"

namespace fsr {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define FSR_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw ::fsr::Error(FSR_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                                         __FILE__ + ":" + std::to_string(__LINE__) + ")");          \
  } while (0)

#define FSR_REQUIRE(cond, msg)                                                  \
  do {                                                                          \
    if (!(cond)) throw ::fsr::Error(FSR_E_INVALID, std::string(msg));           \
  } while (0)

// Launch bookkeeping: every kernel launch of the engine goes through this so bench.py can report
// gpu_launches truthfully (fsr_launch_count).
struct LaunchCounter {
  int64_t n = 0;
};
extern thread_local LaunchCounter* g_launch_counter;
inline void count_launch() {
  if (g_launch_counter) g_launch_counter->n++;
}
#define FSR_LAUNCH_CHECK()                 \
  do {                                     \
    ::fsr::count_launch();                 \
    FSR_CUDA(cudaGetLastError());          \
  } while (0)

// Optional per-stage device timing (bench.py roofline): CUDA events bracket each launch group on the
// launching stream; elapsed times are summed per category when fetched.
enum ProfCat { PROF_PROLOGUE = 0, PROF_LR_CONV, PROF_LR_MISC, PROF_CONVT, PROF_HEAD, PROF_INVERT, PROF_BLEND, PROF_NCAT };
struct Profiler {
  bool on = false;
  struct Rec {
    int cat;
    cudaEvent_t a, b;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  double ms[PROF_NCAT] = {0};
  long long count[PROF_NCAT] = {0};
  // Consecutive LR convolution launches share ONE event pair: an event record between two kernels would cut the
  // programmatic dependent launch edge between them.  The pair is closed by the next scope of any kind (or collect()).
  long long open_idx = -1;
  cudaStream_t open_s = nullptr;
  void close_open() {
    if (open_idx >= 0) cudaEventRecord(recs[(size_t)open_idx].b, open_s);
    open_idx = -1;
  }
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    FSR_CUDA(cudaEventCreate(&e));
    return e;
  }
  void collect() {
    close_open();
    for (auto& r : recs) {
      FSR_CUDA(cudaEventSynchronize(r.b));
      float t = 0.f;
      FSR_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
      ms[r.cat] += t;
      count[r.cat] += 1;
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    recs.clear();
  }
  void reset() {
    collect();
    for (int i = 0; i < PROF_NCAT; ++i) ms[i] = 0, count[i] = 0;
  }
  ~Profiler() {
    for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : pool) cudaEventDestroy(e);
  }
};
struct ProfScope {
  Profiler* p;
  cudaStream_t s;
  size_t idx = 0;
  bool deferred = false;  // the end event is recorded by the next scope (runs of LR convolutions)
  ProfScope(Profiler& prof, int cat, cudaStream_t stream) : p(prof.on ? &prof : nullptr), s(stream) {
    if (!p) return;
    if (p->open_idx >= 0) {
      if (cat == PROF_LR_CONV && p->recs[(size_t)p->open_idx].cat == cat && p->open_s == s) {
        idx = (size_t)p->open_idx;
        deferred = true;
        return;
      }
      p->close_open();
    }
    Profiler::Rec r{cat, p->get(), p->get()};
    FSR_CUDA(cudaEventRecord(r.a, s));
    p->recs.push_back(r);
    idx = p->recs.size() - 1;
    if (cat == PROF_LR_CONV) {
      p->open_idx = (long long)idx;
      p->open_s = s;
      deferred = true;
    }
  }
  ~ProfScope() {
    if (p && !deferred) cudaEventRecord(p->recs[idx].b, s);
  }
};

// Where one tile's pixels come from: a window of a (virtually zero-padded) raster.  A batch of B
// independent tiles is the raster [B*T, T] with origins (t*T, 0).
struct TileGrid {
  const int2* origins;  // device array [n_tiles] of (y0, x0) in HR raster coordinates
  int H, W;             // valid HR raster extent (pixels beyond it read as 0, ResUNet_16x_DEM.py:215-235)
  int Hl, Wl;           // valid LR raster extent
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// SM count of the CURRENT device (grid caps, wave sizing): looked up per device, not per process
inline int current_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cache[dev] && cudaDeviceGetAttribute(&cache[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cache[dev] = 148;
  return cache[dev];
}
// one-time per-device action (function attributes are per device): true the first time it is asked for (key, device)
inline bool first_on_device(bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (done[dev]) return false;
  done[dev] = true;
  return true;
}

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  if (act == FSR_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == FSR_ACT_LEAKY) return v > 0.0f ? v : v * alpha;
  return v;
}

// activations of element-wise ops (the convolution epilogues implement NONE / RELU / LEAKY only)
__device__ __forceinline__ float apply_act2(float v, int act, float alpha, float beta) {
  if (act == FSR_ACT_CLIP) return fminf(fmaxf(v, alpha), beta);
  if (act == FSR_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return apply_act(v, act, alpha);
}

// Source sample of output coordinate `o` of an f-fold bilinear upsampling (ONNX Resize, mode linear): index of the lower
// neighbour, clamped, and the weight of the upper one.  n_in = source extent.
__device__ __forceinline__ void up_linear_coord(int o, int f, int n_in, int mode, int& i0, int& i1, float& w1) {
  float x;
  if (mode == FSR_UP_LINEAR_ALIGN_CORNERS) x = n_in * f > 1 ? (float)o * (float)(n_in - 1) / (float)(n_in * f - 1) : 0.0f;
  else if (mode == FSR_UP_LINEAR_ASYMMETRIC) x = (float)o / (float)f;
  else x = ((float)o + 0.5f) / (float)f - 0.5f;
  x = fminf(fmaxf(x, 0.0f), (float)(n_in - 1));
  i0 = (int)floorf(x);
  i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
  w1 = x - (float)i0;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace fsr
