// Shared host/device helpers for the floodsr_b200 CUDA engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <stdexcept>
#include <string>

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

// Where one tile's pixels come from: a window of a (virtually zero-padded) raster.  A batch of B
// independent tiles is the raster [B*T, T] with origins (t*T, 0).
struct TileGrid {
  const int2* origins;  // device array [n_tiles] of (y0, x0) in HR raster coordinates
  int H, W;             // valid HR raster extent (pixels beyond it read as 0, ResUNet_16x_DEM.py:215-235)
  int Hl, Wl;           // valid LR raster extent
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float apply_act(float v, int act, float alpha) {
  if (act == FSR_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == FSR_ACT_LEAKY) return v > 0.0f ? v : v * alpha;
  return v;
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
