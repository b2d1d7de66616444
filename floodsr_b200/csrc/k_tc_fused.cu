// Fused high-resolution end of ResUNet_16x_DEM on tcgen05: 16x transposed convolution (kernel == stride) -> activation
// -> conv3x3 over concat(features, dem_hr) -> activation -> conv1x1 -> invert_depth_log1p, in ONE persistent kernel.
// ~90 % of the FLOPs behind `session.run` (floodsr/engine/ort.py:193) plus the invert_depth_log1p_np call after it
// (ort.py:196, preprocessing.py:154-164).  The 512 x 512 x 32 feature map F never exists in HBM (the unfused pair
// k_tc_head.cu / k_tc_head2.cu writes and re-reads 33.5 MB of it per tile): a CTA builds each full-width row of F in
// shared memory from the 32 x 32 x 32 low-resolution map L and consumes it at once.
//
// Work unit: one 512-pixel input row y of one tile.  Rows of all tiles are split into equal contiguous ranges per CTA
// (one CTA per SM); a range is walked as items = runs of rows inside one tile (+1 halo row above and below).
//
//   convT row    D_t[(kx, co), cell] = sum_ci Wt[y & 15][kx][ci][co] * L[y >> 4][cell][ci]
//                8 MMAs (4 blocks of 4 kx x 32 co = 128 lanes, K = 32, N = 32 cells) into 128 TMEM columns.
//   F builder    4 warps read D_t, add bias, apply the activation, convert to 16 bit and scatter the row into shared
//                memory as the head's K-major operand [4 channel planes][1 + 512 + 1 px][8 ch] (halo pixels stay zero).
//   head         per 128-pixel strip s: the input row feeds output rows y+1, y, y-1 (ky = 0, 1, 2) with ONE N = 96 MMA
//                per K step into the strip's three accumulators (TMEM columns [96 s, 96 s + 96)): 6 feature K steps (kx
//                taps = 16-byte start offsets into the row) + 1 DEM/bias K step ((hi, lo) 16-bit pairs, built in smem).
//                The three accumulators are a ring of 3: the weight operand is stored [ky2|ky1|ky0|ky2|ky1] along N, so
//                the rotation of the ring is a start offset and every steady-state MMA covers all 96 columns.
//   epilogue     2 warpgroups (strips alternate): read the finished accumulator, re-zero it, activation, 1x1
//                projection, log1p inversion, coalesced fp32 stores.
//
// Warps: 0 loader (TMA) | 1, 21 head MMA issuers (strips 0-1 / 2-3: one issuing warp cannot keep the tensor core fed) |
// 2 DEM prefetch | 3 DEM operand builder | 4-11 epilogue | 12-19 F builders (blocks 0-1 / 2-3) | 20 convT MMA issuer.
// All hand-offs are mbarriers.
#include <stdlib.h>

#include <type_traits>

#include "fsr_engine.cuh"
#include "tc_common.cuh"

namespace fsr {

using namespace tc;

CUtensorMap make_cp8_wide_tensor_map(const void* base, int W, int H, int N, int chunks, long long plane, int bw, int bh, int bn, int kc,
                                     int parts);

namespace {

constexpr int kC = 32;                            // feature channels == head mid channels
constexpr int kStrips = 4;                        // 128-pixel strips per 512-pixel row
constexpr int kW = 512, kCells = 32, kUp = 16;
constexpr int kRowPx = kW + 2;                    // + zero halo pixel left and right
constexpr int kFPlane = (kRowPx - 1) * 16;        // 8208 B between the 8-channel planes of an F row: the right halo pixel of
                                                  // one plane shares its slot with the left halo pixel of the next (both are
                                                  // always zero), which puts the four planes 4 banks apart -> conflict-free stores
constexpr int kFRow = 32896;                      // stage stride (3 * 8208 + 8224, rounded up to 128 B)
constexpr int kFStages = 3;
constexpr int kHwBlocks = 5;                      // [ky2|ky1|ky0|ky2|ky1]
constexpr int kHwStep = 2 * kHwBlocks * kC * 16;  // one K step of the head weights: [2 planes][160][8] = 5120 B
constexpr int kHwBytes = 7 * kHwStep;             // 6 feature K steps + DEM/bias operand
constexpr int kWtRow = 4 * 4 * 128 * 16;          // convT weights of one ky: [4 blocks][4 K planes][128 rows][8] = 32 KB
constexpr int kLRow = 4 * kCells * 16;            // L cells of one LR row: [4 K planes][32 cells][8] = 2 KB
constexpr int kWtHalf = kWtRow / 2;               // a weight stage holds two of the four kx blocks (+ the L row)
constexpr int kWtStage = kWtHalf + kLRow;
constexpr int kWtStages = 3;
constexpr int kA2Row = kW * 16;                   // DEM operand of one row: [512 px][8] = 8 KB
constexpr int kDemRow = 2176;                     // fp32 DEM halo row (514 floats), 128-byte aligned
constexpr int kZero = 2048;
constexpr int kThreads = 22 * 32;
constexpr int kSmemBytes = kHwBytes + kWtStages * kWtStage + kFStages * (kFRow + kA2Row + kDemRow) + kZero + 1024;
constexpr int kDcol = kStrips * 3 * kC;           // first TMEM column of the convT accumulator (384)

struct FusedParams {
  int H, N;             // tile height (rows), tiles in this launch; width is 512
  long long total_rows; // N * H
  int act_t;            // activation after the transposed convolution
  float alpha_t, alpha_h;
  int half;             // 16-bit format: 0 bf16, 1 fp16
  float max_depth, denom;
  const __nv_bfloat16* hw;      // head weights, kHwBytes
  const __nv_bfloat16* wt;      // convT weights [16 ky][kWtRow]
  const float* dem;     // [N][H][512] normalised DEM (src.on == 0)
  DemSource src;        // or: the raw raster + per-tile statistics (src.on == 1)
  float* pred_m;        // [N][H][512]
  float* pred_norm;     // or nullptr
  unsigned* flags;      // FSR_FLAG_PRED_NONFINITE is raised here
  float bias_t[kC];     // convT bias
  float w2[kC];         // 1x1 projection
  float b2;
};

struct RowIter {  // items of this CTA's row range; every warp role iterates the same sequence
  long long r, r_end;
  int H;
  int img, y0, rows;
  __device__ RowIter(const FusedParams& p) : H(p.H) {
    r = p.total_rows * (long long)blockIdx.x / gridDim.x;
    r_end = p.total_rows * (long long)(blockIdx.x + 1) / gridDim.x;
  }
  __device__ bool next() {
    if (r >= r_end) return false;
    const long long t = r / H;
    y0 = (int)(r - t * H);
    const long long left = r_end - r;
    rows = (H - y0) < left ? (H - y0) : (int)left;
    img = (int)t;
    r += rows;
    return true;
  }
};

// try_wait with a short suspend-time hint: the rings here are only 2-3 deep, so a hand-off must be seen within a fraction
// of a row time (a 2 us hint serialised builder and head), while plain spinning would steal shared-memory bandwidth
__device__ __forceinline__ void wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(96u)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 26)) __trap();
  }
}

// MMA with the 64-bit shared-memory descriptors given as 32-bit halves (layout: tc_common.cuh, smem_desc_kmajor)
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .b64 da, db;\n"
      ".reg .pred p;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.u32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint32_t dlo(uint32_t smem_addr, uint32_t lbo_bytes) { return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ void commit_to(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint16_t to16(float v, int half) {
  if (half) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float from16(uint16_t u, int half) {
  if (half) return __half2float(__ushort_as_half(u));
  return __bfloat162float(__ushort_as_bfloat16(u));
}
template <int ACT>
__device__ __forceinline__ float act_fn(float v, float alpha) {
  if (ACT == FSR_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == FSR_ACT_LEAKY) return v > 0.0f ? v : v * alpha;
  return v;
}

template <int ACT, int ACT_T, int HALF>
__global__ void __launch_bounds__(kThreads, 1)  // 80 registers: the register file is split per SM sub-partition (16 K each), which
                                                // holds 6 of the 22 warps: 6 x 32 x 88 would not fit
fused_hr_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ FusedParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem_hw = smem_raw;                                  // head weights
  uint8_t* smem_wt = smem_hw + kHwBytes;                        // kWtStages x (convT weights of one ky + L row)
  uint8_t* smem_f = smem_wt + kWtStages * kWtStage;             // kFStages x F row
  uint8_t* smem_a2 = smem_f + kFStages * kFRow;                 // kFStages x DEM operand row
  uint8_t* smem_zero = smem_a2 + kFStages * kA2Row;             // zeros: upper K plane of every DEM operand
  uint8_t* smem_dem = smem_zero + kZero;                        // kFStages x fp32 DEM halo row
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dem + kFStages * kDemRow);
  uint64_t* hw_full = bars;
  uint64_t* wt_full = bars + 1;                  // [kWtStages]  TMA -> convT issuer
  uint64_t* wt_empty = wt_full + kWtStages;      // [kWtStages]  convT MMAs done -> loader
  uint64_t* d_full = wt_empty + kWtStages;       //              convT MMAs done -> F builder
  uint64_t* d_empty = d_full + 1;                //              F builder (4 warps) has read D -> convT issuer
  uint64_t* f_full = d_empty + 1;                // [kFStages]   F builder (4 warps) + DEM builder -> head issuer
  uint64_t* f_empty = f_full + kFStages;         // [kFStages]   head MMAs done -> F builder, DEM prefetch
  uint64_t* dem_full = f_empty + kFStages;       // [kFStages]   DEM prefetch (32 cp.async arrivals) -> DEM builder
  uint64_t* slot_full = dem_full + kFStages;     // [kStrips][3] head MMAs done -> epilogue
  uint64_t* slot_empty = slot_full + kStrips * 3;  // [kStrips][3] epilogue (4 warps) -> head issuer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_empty + kStrips * 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmL);
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(hw_full, 1);
      for (int i = 0; i < kWtStages; ++i) {
        mbar_init(&wt_full[i], 1);
        mbar_init(&wt_empty[i], 1);
      }
      mbar_init(d_full, 1);
      mbar_init(d_empty, 8);
      for (int i = 0; i < kFStages; ++i) {
        mbar_init(&f_full[i], 9);
        mbar_init(&f_empty[i], 2);
        mbar_init(&dem_full[i], 32);
      }
      for (int i = 0; i < kStrips * 3; ++i) {
        mbar_init(&slot_full[i], 1);
        mbar_init(&slot_empty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // zero plane, and the F stages (their halo pixels must stay zero; everything else is rewritten per row)
  for (int i = threadIdx.x; i < kZero / 16; i += kThreads) reinterpret_cast<uint4*>(smem_zero)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < kFStages * kFRow / 16; i += kThreads) reinterpret_cast<uint4*>(smem_f)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp >= 4 && warp < 8) {
    // every head MMA accumulates: start from zeroed accumulators (the epilogue re-zeroes a slot after reading it)
    for (int c = 0; c < kDcol / 32; ++c) tmem_zero32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + c * 32);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1

  if (warp == 0) {
    // ===================== loader: head weights once; per valid input row the convT weights of its ky + its L cells ====
    if (lane == 0) {
      mbar_expect_tx(hw_full, kHwBytes);
      bulk_load_1d(smem_hw, p.hw, kHwBytes, hw_full);
      int st = 0;
      uint32_t ph = 1;
      for (RowIter it(p); it.next();) {
        for (int i = 0; i < it.rows + 2; ++i) {
          const int y = it.y0 - 1 + i;
          if (y < 0 || y >= p.H) continue;
          for (int h = 0; h < 2; ++h) {
            wait_relaxed(&wt_empty[st], ph);
            mbar_expect_tx(&wt_full[st], kWtStage);
            uint8_t* dst = smem_wt + st * kWtStage;
            bulk_load_1d(dst, reinterpret_cast<const uint8_t*>(p.wt) + (size_t)(y & (kUp - 1)) * kWtRow + h * kWtHalf, kWtHalf, &wt_full[st]);
            tma_load_5d(dst + kWtHalf, &tmL, &wt_full[st], 0, y / kUp, it.img, 0, 0);
            if (++st == kWtStages) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 20) {
    // ===================== convT MMA issuer: D_t[(kx, co), cell] for one row ===============================
    const uint32_t idesc = idesc_16(128, kCells, HALF);
    const uint32_t wt0 = smem_u32(smem_wt);
    const bool leader = elect_one();
    int st = 0;
    uint32_t ph = 0, dph = 1;
    for (RowIter it(p); it.next();) {
      for (int i = 0; i < it.rows + 2; ++i) {
        const int y = it.y0 - 1 + i;
        if (y < 0 || y >= p.H) continue;
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&wt_full[st], ph);
          if (h == 0) mbar_wait(d_empty, dph);
          tc_fence_after();
          if (leader) {
            const uint32_t wbase = wt0 + st * kWtStage;
            const uint32_t b0 = dlo(wbase + kWtHalf, kCells * 16);          // L: [4 planes][32 cells][16 B]
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
              const uint32_t a0 = dlo(wbase + bb * (4 * 2048), 2048);        // block 2 h + bb: [4 planes][128 rows][16 B]
              const uint32_t dcol = tmem_base + kDcol + (2 * h + bb) * kCells;
              umma2(dcol, a0, b0, desc_hi, idesc, 0u);
              umma2(dcol, a0 + ((2 * 2048) >> 4), b0 + ((2 * kCells * 16) >> 4), desc_hi, idesc, 1u);
            }
            commit_to(smem_u32(&wt_empty[st]));
            if (h == 1) commit_to(smem_u32(d_full));
          }
          __syncwarp();
          if (++st == kWtStages) { st = 0; ph ^= 1; }
        }
        dph ^= 1;
      }
    }
  } else if (warp >= 12 && warp < 20) {
    // ===================== F builders: TMEM D_t -> + bias, activation, 16 bit -> shared-memory operand row ===========
    // warp group gb handles blocks 2 gb and 2 gb + 1 (kx = 4 b + q); lane == co, TMEM lane quarter q == kx within the block
    const int gb = (warp - 12) >> 2;
    const int q = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + kDcol;
    const float bias = p.bias_t[lane];
    int fs = 0;
    uint32_t fph = 1, dph = 0;
    for (RowIter it(p); it.next();) {
      for (int i = 0; i < it.rows + 2; ++i) {
        const int y = it.y0 - 1 + i;
        uint8_t* frow = smem_f + fs * kFRow;
        wait_relaxed(&f_empty[fs], fph);
        if (y < 0 || y >= p.H) {
          // zero padding row of the head convolution
          for (int k = (warp - 12) * 32 + lane; k < kFRow / 16; k += 256) reinterpret_cast<uint4*>(frow)[k] = make_uint4(0, 0, 0, 0);
        } else {
          wait_relaxed(d_full, dph);
          dph ^= 1;
          tc_fence_after();
          float v0[32], v1[32];
          tmem_ld32(lane_addr + (2 * gb) * kCells, v0);
          tmem_ld32(lane_addr + (2 * gb + 1) * kCells, v1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty);
          // lanes 2k / 2k+1 hold channels 2k / 2k+1: one shuffle per cell lets the even lane store the channel pair of block
          // 2 gb and the odd lane that of block 2 gb + 1 as 32-bit words (32 lanes -> 32 distinct banks)
          const bool odd = lane & 1;
          __half2 hmax = __floats2half2_rn(0.0f, 0.0f);  // largest |value| stored (HALF): Inf = it did not fit fp16 (flag_unstorable, k_tc_conv.cu)
          uint8_t* dst0 = frow + (uint32_t)(lane >> 3) * kFPlane + (uint32_t)((lane & 7) >> 1) * 4 +
                          (uint32_t)(1 + 4 * (2 * gb + (odd ? 1 : 0)) + q) * 16;
#pragma unroll
          for (int c = 0; c < kCells; ++c) {
            const float a = act_fn<ACT_T>(v0[c] + bias, p.alpha_t);   // block 2 gb, own channel
            const float b = act_fn<ACT_T>(v1[c] + bias, p.alpha_t);   // block 2 gb + 1, own channel

            const float other = __shfl_xor_sync(0xffffffffu, odd ? a : b, 1);
            const uint32_t w = odd ? pack_x2(other, b, HALF) : pack_x2(a, other, HALF);
            *reinterpret_cast<uint32_t*>(dst0 + c * (kUp * 16)) = w;
#ifndef FSR_NO_BUILDER_CHECK
            if (HALF) hmax = __hmax2(hmax, __habs2(*reinterpret_cast<const __half2*>(&w)));  // one instruction per two values
#endif
          }
          if (__hisinf(__low2half(hmax)) || __hisinf(__high2half(hmax))) atomicOr(p.flags, FSR_FLAG_PRED_NONFINITE);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&f_full[fs]);
        if (++fs == kFStages) { fs = 0; fph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ===================== DEM prefetcher: fp32 halo row (514 px) -> smem ===================================
    // src.on: the row comes from the raw raster window of the tile (zero beyond the raster, as the reference's padding to
    // whole tiles); the operand builder normalises it.  Pixels outside the TILE are the head convolution's zero padding.
    int fs = 0;
    uint32_t fph = 1;
    for (RowIter it(p); it.next();) {
      int oy = 0, ox = 0;
      if (p.src.on) {
        const int2 org = __ldg(p.src.origins + it.img);
        oy = org.x;
        ox = org.y;
      }
      for (int i = 0; i < it.rows + 2; ++i) {
        const int y = it.y0 - 1 + i;
        const bool yok = y >= 0 && y < p.H && (!p.src.on || oy + y < p.src.H);
        const float* row = p.src.on ? p.src.ras + (size_t)(oy + (yok ? y : 0)) * p.src.W + ox
                                    : p.dem + ((size_t)it.img * p.H + (yok ? y : 0)) * kW;
        const int x_end = p.src.on ? (p.src.W - ox < kW ? p.src.W - ox : kW) : kW;
        const uint32_t dst = smem_u32(smem_dem + fs * kDemRow);
        wait_relaxed(&f_empty[fs], fph);
        for (int k = lane; k < kRowPx; k += 32) {
          const int x = k - 1;
          const bool ok = yok && x >= 0 && x < x_end;
          const float* src = row + (ok ? x : 0);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + k * 4), "l"(src), "r"(ok ? 4 : 0) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&dem_full[fs])) : "memory");
        if (++fs == kFStages) { fs = 0; fph ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================== DEM operand builder: [512 px][hi(-1,0,+1), lo(-1,0,+1), 1, 1] ====================
    const uint32_t ones = (uint32_t)to16(1.0f, HALF) * 0x10001u;
    int fs = 0;
    uint32_t ph = 0;
    for (RowIter it(p); it.next();) {
      float p_clip = 0.f, dem_min = 0.f, range_f = 1.f;
      bool zero_out = false;
      if (p.src.on) dem_norm_spec(p.src.stats + (size_t)it.img * 3, p_clip, dem_min, range_f, zero_out);
      for (int i = 0; i < it.rows + 2; ++i) {
        wait_relaxed(&dem_full[fs], ph);
        if (p.src.on) {
          // raw raster values -> normalised DEM, in place; pixels outside the tile stay the convolution's zero padding
          const int y = it.y0 - 1 + i;
          const bool in_tile_row = y >= 0 && y < p.H;
          float* drow = reinterpret_cast<float*>(smem_dem + fs * kDemRow);
          for (int k = lane; k < kRowPx; k += 32) {
            const bool in_tile = in_tile_row && k >= 1 && k <= kW;
            drow[k] = in_tile ? dem_normalise_px(drow[k], p.src, p_clip, dem_min, range_f, zero_out) : 0.0f;
          }
          __syncwarp();
        }
        // a lane owns 4 consecutive pixels: 6 loads + 6 conversions serve 4 pixels.  (The lane-per-pixel mapping that avoids the
        // bank conflicts of these 16-byte stores converts every value three times; this warp is on the critical path of a
        // 1.4 us row and that version measured 4.6 -> 5.6 ms for the kernel.)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int px0 = g * 128 + lane * 4;
          const float* drow = reinterpret_cast<const float*>(smem_dem + fs * kDemRow) + px0;  // halo index of px0 - 1
          uint16_t hi[6], lo[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const float d = drow[k];
            hi[k] = to16(d, HALF);
            lo[k] = to16(d - from16(hi[k], HALF), HALF);
          }
          uint4* dst = reinterpret_cast<uint4*>(smem_a2 + fs * kA2Row) + px0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint4 v;
            v.x = (uint32_t)hi[e] | ((uint32_t)hi[e + 1] << 16);
            v.y = (uint32_t)hi[e + 2] | ((uint32_t)lo[e] << 16);
            v.z = (uint32_t)lo[e + 1] | ((uint32_t)lo[e + 2] << 16);
            v.w = ones;
            dst[e] = v;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&f_full[fs]);
        if (++fs == kFStages) { fs = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 21) {
    // ===================== head MMA issuers (warp 1: strips 0-1, warp 21: strips 2-3) ====================================================================
    const uint32_t idesc0 = idesc_16(128, 0, HALF);  // + (N >> 3) << 17
    const int s_begin = warp == 1 ? 0 : 2;
    const uint32_t idesc96 = idesc0 + (3u << 19), idesc32 = idesc0 + (1u << 19);
    mbar_wait(hw_full, 0);
    const uint32_t b_lo0 = dlo(smem_u32(smem_hw), kHwBlocks * kC * 16);
    const uint32_t f_lo0 = dlo(smem_u32(smem_f), kFPlane);
    const uint32_t zero_addr = smem_u32(smem_zero);
    const uint32_t bar_slot_full = smem_u32(slot_full), bar_f_empty = smem_u32(f_empty);
    const bool leader = elect_one();
    constexpr uint32_t kB = kHwStep >> 4, kP = (2 * kFPlane) >> 4;
    int fs = 0;
    uint32_t fph = 0;
    int go = 0;  // output rows of earlier items (ring position base)
    for (RowIter it(p); it.next(); go += it.rows) {
      for (int i = 0; i < it.rows + 2; ++i) {
        mbar_wait(&f_full[fs], fph);
        tc_fence_after();
        const int R = go + i;                 // ring index of the output row this input row opens
        const int r0 = (R + 1) % 3;           // slot of output row i - 2 ( (R - 2) mod 3 )
        const int t0 = (3 - r0) % 3;          // first weight block so that slot order matches [ky2|ky1|ky0] rotation
        const bool steady = i >= 2 && i < it.rows;
        const uint32_t f_row = f_lo0 + fs * (kFRow >> 4);
        for (int s = s_begin; s < s_begin + 2; ++s) {
          if (i < it.rows) mbar_wait(&slot_empty[s * 3 + R % 3], ((R / 3) & 1) ^ 1);  // read + re-zeroed by the epilogue
          tc_fence_after();
          const uint32_t a = f_row + s * 128;   // 128 pixels x 16 B >> 4
          const uint32_t a2_addr = smem_u32(smem_a2) + fs * kA2Row + s * 2048;
          const uint32_t a2 = dlo(a2_addr, zero_addr - a2_addr);
          const uint32_t d = tmem_base + s * (3 * kC);
          if (leader) {
            if (steady) {
              const uint32_t b = b_lo0 + t0 * kC;
              umma2(d, a, b, desc_hi, idesc96, 1u);
              umma2(d, a + kP, b + kB, desc_hi, idesc96, 1u);
              umma2(d, a + 1, b + 2 * kB, desc_hi, idesc96, 1u);
              umma2(d, a + kP + 1, b + 3 * kB, desc_hi, idesc96, 1u);
              umma2(d, a + 2, b + 4 * kB, desc_hi, idesc96, 1u);
              umma2(d, a + kP + 2, b + 5 * kB, desc_hi, idesc96, 1u);
              umma2(d, a2, b + 6 * kB, desc_hi, idesc96, 1u);
            } else {
              // item borders: one N = 32 MMA set per existing target row j = i - 2 + t (weight block t <-> ky = 2 - t)
              for (int t = 0; t < 3; ++t) {
                const int j = i - 2 + t;
                if (j < 0 || j >= it.rows) continue;
                const uint32_t dj = d + ((go + j) % 3) * kC;
                const uint32_t b = b_lo0 + t * kC;
                umma2(dj, a, b, desc_hi, idesc32, 1u);
                umma2(dj, a + kP, b + kB, desc_hi, idesc32, 1u);
                umma2(dj, a + 1, b + 2 * kB, desc_hi, idesc32, 1u);
                umma2(dj, a + kP + 1, b + 3 * kB, desc_hi, idesc32, 1u);
                umma2(dj, a + 2, b + 4 * kB, desc_hi, idesc32, 1u);
                umma2(dj, a + kP + 2, b + 5 * kB, desc_hi, idesc32, 1u);
                umma2(dj, a2, b + 6 * kB, desc_hi, idesc32, 1u);
              }
            }
            if (i >= 2) commit_to(bar_slot_full + (s * 3 + r0) * 8);
          }
          __syncwarp();
        }
        if (leader) commit_to(bar_f_empty + fs * 8);
        __syncwarp();
        if (++fs == kFStages) { fs = 0; fph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 4-11): strips alternate between the two warpgroups ================
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float w2r[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) w2r[c] = p.w2[c];
    int go = 0;
    for (RowIter it(p); it.next(); go += it.rows) {
      for (int j = 0; j < it.rows; ++j) {
        const int r = go + j;
        const int slot = r % 3;
        const uint32_t par = (r / 3) & 1;
        for (int s = grp; s < kStrips; s += 2) {
          wait_relaxed(&slot_full[s * 3 + slot], par);
          tc_fence_after();
          const uint32_t taddr = lane_addr + s * (3 * kC) + slot * kC;
          float v[kC];
          tmem_ld32(taddr, v);
          tmem_ld_wait();
          tmem_zero32(taddr);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&slot_empty[s * 3 + slot]);
          float o0 = p.b2, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
#pragma unroll
          for (int c = 0; c < kC; c += 4) {
            o0 = fmaf(act_fn<ACT>(v[c], p.alpha_h), w2r[c], o0);
            o1 = fmaf(act_fn<ACT>(v[c + 1], p.alpha_h), w2r[c + 1], o1);
            o2 = fmaf(act_fn<ACT>(v[c + 2], p.alpha_h), w2r[c + 2], o2);
            o3 = fmaf(act_fn<ACT>(v[c + 3], p.alpha_h), w2r[c + 3], o3);
          }
          const float out = (o0 + o1) + (o2 + o3);
          const size_t off = ((size_t)it.img * p.H + (it.y0 + j)) * kW + s * 128 + m;
#ifndef FSR_NO_EPI_CHECK
          if (!(fabsf(out) <= 3.0e38f)) atomicOr(p.flags, FSR_FLAG_PRED_NONFINITE);  // Inf / NaN: never in a healthy run
#endif
          if (p.pred_norm) p.pred_norm[off] = out;
          const float yn = fminf(fmaxf(out, 0.0f), 1.0f);
          p.pred_m[off] = fminf(fmaxf(expm1f(__fmul_rn(yn, p.denom)), 0.0f), p.max_depth);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------

bool fused_hr_ok(int H, int W, int lr_h, int lr_w, int cin_t, int cout_t, int k_t, int cmid, int ksz) {
  return W == kW && lr_w == kCells && k_t == kUp && H == lr_h * kUp && cin_t == 32 && cout_t == kC && cmid == kC && ksz == 3;
}

size_t fused_hw_elems() { return (size_t)kHwBytes / 2; }
size_t fused_wt_elems() { return (size_t)kUp * kWtRow / 2; }

// head weights w [3][3][33][32] (+ bias [32]) -> [7 K steps][2 planes][160 columns][8]; column n = blk * 32 + co with
// ky = 2 - blk % 3.  K steps 0-5 = (kx, 16-channel slice); step 6 = DEM/bias operand: k 0-2 = Wdem[ky][kx], k 3-5 the same
// (they multiply the lo parts), k 6/7 = (hi, lo) of bias[co] for ky == 1.
void fused_pack_head(const float* w, const float* bias, uint16_t* dst, uint16_t (*cvt)(float), float (*back)(uint16_t)) {
  const int cin_real = 33;
  size_t pos = 0;
  for (int kx = 0; kx < 3; ++kx)
    for (int j = 0; j < 2; ++j)
      for (int pl = 0; pl < 2; ++pl)
        for (int n = 0; n < kHwBlocks * kC; ++n)
          for (int e = 0; e < 8; ++e, ++pos) {
            const int ky = 2 - (n / kC) % 3, co = n % kC, ci = j * 16 + pl * 8 + e;
            dst[pos] = cvt(w[(((size_t)ky * 3 + kx) * cin_real + ci) * kC + co]);
          }
  for (int pl = 0; pl < 2; ++pl)
    for (int n = 0; n < kHwBlocks * kC; ++n)
      for (int e = 0; e < 8; ++e, ++pos) {
        const int ky = 2 - (n / kC) % 3, co = n % kC;
        uint16_t v = cvt(0.0f);
        if (pl == 0 && e < 6) v = cvt(w[(((size_t)ky * 3 + (e % 3)) * cin_real + 32) * kC + co]);
        if (pl == 0 && ky == 1 && bias) {
          const uint16_t bh = cvt(bias[co]);
          if (e == 6) v = bh;
          if (e == 7) v = cvt(bias[co] - back(bh));
        }
        dst[pos] = v;
      }
}

// convT weights w [16 ky][16 kx][32 ci][32 co] -> [ky][4 blocks][4 K planes][128 rows = (kx % 4) * 32 + co][8 ci]
void fused_pack_convt(const float* w, uint16_t* dst, uint16_t (*cvt)(float)) {
  size_t pos = 0;
  for (int ky = 0; ky < kUp; ++ky)
    for (int b = 0; b < 4; ++b)
      for (int pl = 0; pl < 4; ++pl)
        for (int row = 0; row < 128; ++row)
          for (int e = 0; e < 8; ++e, ++pos) {
            const int kx = 4 * b + row / kC, co = row % kC, ci = pl * 8 + e;
            dst[pos] = cvt(w[(((size_t)ky * kUp + kx) * 32 + ci) * kC + co]);
          }
}

void launch_fused_hr_tc(const __nv_bfloat16* lr, long long lr_plane, const __nv_bfloat16* wt_pack, const float* bias_t, int act_t,
                        float alpha_t, const __nv_bfloat16* hw_pack, const float* w2, const float* b2, int act_h, float alpha_h,
                        const float* dem, const DemSource& src, float* pred_m, float* pred_norm, int n_img, int H, float max_depth,
                        float denom, int half, int n_sms, unsigned* flags, cudaStream_t s) {
  FusedParams p{};
  p.H = H;
  p.N = n_img;
  p.total_rows = (long long)n_img * H;
  p.act_t = act_t;
  p.alpha_t = alpha_t;
  p.alpha_h = alpha_h;
  p.half = half;
  p.max_depth = max_depth;
  p.denom = denom;
  p.hw = hw_pack;
  p.wt = wt_pack;
  p.dem = dem;
  p.src = src;
  p.pred_m = pred_m;
  p.pred_norm = pred_norm;
  p.flags = flags;
  for (int c = 0; c < kC; ++c) {
    p.bias_t[c] = bias_t ? bias_t[c] : 0.0f;
    p.w2[c] = w2[c];
  }
  p.b2 = b2 ? b2[0] : 0.0f;
  // L as TMA source: one LR row of 32 cells, all 4 channel planes -> [4 planes][32 cells][8] (the convT B operand)
  CUtensorMap mL = make_cp8_wide_tensor_map(lr, kCells, H / kUp, n_img, 4, lr_plane, kCells, 1, 1, 4, 1);
  const int grid = p.total_rows < n_sms ? (int)p.total_rows : n_sms;
  auto go = [&](auto kernel) {
    FSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    kernel<<<grid, kThreads, kSmemBytes, s>>>(mL, p);
  };
  auto by_half = [&](auto ah, auto at) {
    constexpr int AH = decltype(ah)::value, AT = decltype(at)::value;
    if (half) go(fused_hr_kernel<AH, AT, 1>);
    else go(fused_hr_kernel<AH, AT, 0>);
  };
  auto by_act_t = [&](auto ah) {
    if (act_t == FSR_ACT_RELU) by_half(ah, std::integral_constant<int, FSR_ACT_RELU>{});
    else if (act_t == FSR_ACT_LEAKY) by_half(ah, std::integral_constant<int, FSR_ACT_LEAKY>{});
    else by_half(ah, std::integral_constant<int, FSR_ACT_NONE>{});
  };
  if (act_h == FSR_ACT_RELU) by_act_t(std::integral_constant<int, FSR_ACT_RELU>{});
  else if (act_h == FSR_ACT_LEAKY) by_act_t(std::integral_constant<int, FSR_ACT_LEAKY>{});
  else by_act_t(std::integral_constant<int, FSR_ACT_NONE>{});
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
