// Micro-benchmark: tcgen05.mma issue rate for different smem operand layouts (run on a B200).
#include <cstdio>
#include <cstdlib>
#include "../../floodsr_b200/csrc/tc_common.cuh"
using namespace fsr::tc;

__device__ __forceinline__ uint64_t desc_generic(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// mode 0: no-swizzle, A plane stride 2080 (head layout); 1: no-swizzle, plane stride 2048; 2: SWIZZLE_128B canonical
__global__ void __launch_bounds__(128, 1) mma_bench(int mode, int N, int iters, int a_step, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&tslot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  if (warp == 0) {
    const uint32_t idesc = idesc_16(128, N, 1);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    uint64_t da[3], db;
    for (int k = 0; k < 3; ++k) {
      if (mode == 0) da[k] = desc_generic(a + k * a_step, 2080, 128, 0);
      else if (mode == 1) da[k] = desc_generic(a + k * a_step, 2048, 128, 0);
      else da[k] = desc_generic(a + k * 32, 16, 1024, 2);
    }
    db = mode == 2 ? desc_generic(b, 16, 1024, 2) : desc_generic(b, N * 16, 128, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; it += 6) {
      if (elect_one()) {
        umma_bf16(tm, da[0], db, idesc, 1u);
        umma_bf16(tm, da[1], db, idesc, 1u);
        umma_bf16(tm, da[2], db, idesc, 1u);
        umma_bf16(tm + 128, da[0], db, idesc, 1u);
        umma_bf16(tm + 128, da[1], db, idesc, 1u);
        umma_bf16(tm + 128, da[2], db, idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int iters = 3000;
  for (int grid : {1, 148}) for (int N : {32, 64, 96, 128, 256}) for (int mode : {0, 1, 2}) for (int step : {0, 16}) {
    mma_bench<<<grid, 128, 96 * 1024>>>(mode, N, iters, step, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("grid %3d N %3d mode %d step %2d: %.1f cycles/MMA (ideal %d) %s\n", grid, N, mode, step, (double)h / iters, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
