// Micro-benchmark: tcgen05.ld throughput/latency (run on a B200).
#include <cstdio>
#include "../../floodsr_b200/csrc/tc_common.cuh"
using namespace fsr::tc;

__global__ void __launch_bounds__(512, 1) tmem_bench(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (mode == 0) {          // back-to-back x32 loads, wait each (latency-bound per warp)
    for (int it = 0; it < iters; ++it) {
      float v[32];
      tmem_ld32(tm + (it % 8) * 32, v);
      tmem_ld_wait();
      acc += v[0] + v[13] + v[31];
    }
  } else if (mode == 1) {   // 3 loads in flight, one wait
    for (int it = 0; it < iters; it += 3) {
      float a[32], b[32], c[32];
      tmem_ld32(tm + 0, a);
      tmem_ld32(tm + 32, b);
      tmem_ld32(tm + 64, c);
      tmem_ld_wait();
      acc += a[0] + a[31] + b[5] + b[30] + c[9] + c[31];
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tslot, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 148);
  float* sink; cudaMalloc(&sink, 4 * 148 * 512);
  const int iters = 3000;
  for (int threads : {128, 256, 512}) for (int mode : {0, 1}) {
    tmem_bench<<<148, threads>>>(mode, iters, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h / iters;
    printf("threads %3d mode %d: %.1f cycles per x32 load per warp; SM-wide %.1f B/cycle %s\n", threads, mode, cyc,
           (threads / 32) * 4096.0 / cyc, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
