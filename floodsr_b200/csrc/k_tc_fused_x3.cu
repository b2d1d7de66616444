// Fused high-resolution end of ResUNet_16x_DEM in the fp32-tolerance mode (FSR_PREC_FP32): the same layers as
// k_tc_fused.cu (16x transposed convolution -> activation -> conv3x3 over concat(features, dem_hr) -> activation ->
// conv1x1 -> invert_depth_log1p; floodsr/engine/ort.py:193-196, preprocessing.py:154-164) with SPLIT fp16 operands:
// every activation and weight is the pair (hi, lo) = (fp16(v), fp16(v - hi)) and every product is three tcgen05 MMAs,
// a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, accumulated in fp32 in TMEM (tc_common.cuh).  ONNX Runtime computes these layers in
// fp32; the split products keep 22 significant bits, which holds the <= 1e-4 m tolerance of the north star.
//
// Differences from the 16-bit kernel, all forced by the doubled operands:
//   * the shared-memory ring of feature rows holds (row, part) stages: the F builders write the hi row, then the lo row; the
//     head issuers run  F_hi x [W_hi, W_lo]  (12 MMAs per strip) on the first and  F_lo x W_hi + the DEM steps (8 MMAs) on
//     the second, so the 512 x 512 x 32 feature map still never exists in HBM;
//   * the convT weight ring streams (kx block, K half) stages [hi|lo weights 8 KB + hi|lo L cells 2 KB], three MMAs each;
//   * weights carry a power-of-two factor (their lo parts stay normal fp16 numbers) that the F builders / the epilogue undo;
//     both biases are added in fp32 there instead of riding on constant-one operand columns;
//   * the DEM column is two K steps: [dem_hi(-1,0,+1), dem_lo(-1,0,+1)] x [Wd_hi, Wd_hi] and [dem_hi(-1,0,+1)] x [Wd_lo];
//   * DEM rows have their own 2-deep ring (prefetch -> operand builder -> head issuers).
//
// Warps: 0 loader (TMA) | 1, 25 head MMA issuers (strips 0-1 / 2-3) | 2 DEM prefetch | 3 DEM operand builder |
// 4-7 epilogue | 8-23 F builders (2 block pairs x 4 TMEM lane quarters x 2 halves of the row's cells) | 24 convT MMA issuer.
// All hand-offs are mbarriers.  The F builders are what the row time follows (measured), hence sixteen of them with 16 cells
// each; the epilogue has three MMAs' worth of time per product and needs only one warpgroup.
#include <stdlib.h>

#include <type_traits>

#include "fsr_engine.cuh"
#include "tc_common.cuh"

namespace fsr {

using namespace tc;

CUtensorMap make_cp8_wide_tensor_map(const void* base, int W, int H, int N, int chunks, long long plane, int bw, int bh, int bn, int kc,
                                     int parts);
void split_weight(float v, uint16_t& hi, uint16_t& lo);

namespace {

constexpr int kC = 32;                            // feature channels == head mid channels
constexpr int kStrips = 4;                        // 128-pixel strips per 512-pixel row
constexpr int kW = 512, kCells = 32, kUp = 16;
constexpr int kRowPx = kW + 2;                    // + zero halo pixel left and right
constexpr int kFPlane = (kRowPx - 1) * 16;        // 8208 B between the 8-channel planes of an F row (see k_tc_fused.cu)
constexpr int kFRow = 32896;                      // stage stride (3 * 8208 + 8224, rounded up to 128 B): one part of one row
constexpr int kFStages = 3;                       // (row, part) stages: 1.5 rows
constexpr int kHwBlocks = 5;                      // [ky2|ky1|ky0|ky2|ky1]
constexpr int kHwStep = 2 * kHwBlocks * kC * 16;  // one K step of the head weights: [2 planes][160][8] = 5120 B
constexpr int kHwSteps = 14;                      // 6 W_hi + 6 W_lo feature steps, 2 DEM steps
constexpr int kHwBytes = kHwSteps * kHwStep;      // 71 680 B
constexpr int kWtPart = 2 * 128 * 16;             // convT weights of one (kx block, K half, part): [2 planes][128 rows][8] = 4 KB
constexpr int kLPart = 2 * kCells * 16;           // L cells of one (LR row, K half, part): [2 planes][32 cells][8] = 1 KB
constexpr int kWtStage = 2 * kWtPart + 2 * kLPart;  // 10 240 B
constexpr int kWtStages = 3;
constexpr int kA2Row = kW * 16;                   // DEM operand of one row: [512 px][8] = 8 KB
constexpr int kA2Stages = 2;
constexpr int kDemRow = 2176;                     // fp32 DEM halo row (514 floats), 128-byte aligned
constexpr int kZero = 2048;
constexpr int kThreads = 26 * 32;
constexpr int kBuilders = 16;                     // F-builder warps (8 .. 23)
constexpr int kCellsPerBuilder = kCells / 2;
constexpr int kSmemBytes = kHwBytes + kWtStages * kWtStage + kFStages * kFRow + kA2Stages * (kA2Row + kDemRow) + kZero + 1024;
constexpr int kDcol = kStrips * 3 * kC;           // first TMEM column of the convT accumulator (384)
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

struct X3Params {
  int H, N;             // tile height (rows), tiles in this launch; width is 512
  long long total_rows; // N * H
  float alpha_t, alpha_h;
  float scale_t_inv, scale_h_inv;  // inverse power-of-two factors of the packed convT / head weights
  unsigned wait_epi_ns, wait_build_ns, wait_dem_ns;  // suspend-time hints of the epilogue / F builder / DEM + loader roles
  float max_depth, denom;
  const __nv_bfloat16* hw;      // head weights, kHwBytes
  const __nv_bfloat16* wt;      // convT weights [16 ky][4 blocks][2 K halves][2 parts][kWtPart]
  const float* dem;     // [N][H][512] normalised DEM (src.on == 0)
  DemSource src;        // or: the raw raster + per-tile statistics (src.on == 1)
  float* pred_m;        // [N][H][512]
  float* pred_norm;     // or nullptr
  unsigned* flags;      // FSR_FLAG_PRED_NONFINITE is raised here
  float bias_t[kC];     // convT bias
  float bias_h[kC];     // head conv3x3 bias
  float w2[kC];         // 1x1 projection
  float b2;
};

struct RowIter {  // items of this CTA's row range; every warp role iterates the same sequence
  long long r, r_end;
  int H;
  int img, y0, rows;
  __device__ RowIter(const X3Params& p) : H(p.H) {
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

// try_wait with a suspend-time hint.  Every poll is a shared-memory access, and the shared-memory data pipe is what bounds this
// kernel (tensor-core operand wavefronts 75 % + LSU wavefronts 22 % of peak in the ncu capture, a third of the latter barrier
// polls): roles that have a whole row time (~4 us here) of slack poll rarely, roles on the critical path poll often.
__device__ __forceinline__ void wait_relaxed(uint64_t* bar, uint32_t parity, uint32_t hint_ns = 96u) {
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
        : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 26)) __trap();
  }
}

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

// (hi, lo) fp16 pair of v in one word: hi in bits 0-15, lo in bits 16-31
__device__ __forceinline__ uint32_t split16(float v) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  return (uint32_t)__half_as_ushort(h) | ((uint32_t)__half_as_ushort(l) << 16);
}
template <int ACT>
__device__ __forceinline__ float act_fn(float v, float alpha) {
  if (ACT == FSR_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == FSR_ACT_LEAKY) return v > 0.0f ? v : v * alpha;
  return v;
}

template <int ACT, int ACT_T>
__global__ void __launch_bounds__(kThreads, 1)  // 72 registers: the register file is split per SM sub-partition (16 K each), which
                                                // holds 7 of the 26 warps
fused_hr_x3_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ X3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem_hw = smem_raw;                                  // head weights
  uint8_t* smem_wt = smem_hw + kHwBytes;                        // kWtStages x (convT weights hi|lo + L cells hi|lo)
  uint8_t* smem_f = smem_wt + kWtStages * kWtStage;             // kFStages x one part of an F row
  uint8_t* smem_a2 = smem_f + kFStages * kFRow;                 // kA2Stages x DEM operand row
  uint8_t* smem_zero = smem_a2 + kA2Stages * kA2Row;            // zeros: upper K plane of every DEM operand
  uint8_t* smem_dem = smem_zero + kZero;                        // kA2Stages x fp32 DEM halo row
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dem + kA2Stages * kDemRow);
  uint64_t* hw_full = bars;
  uint64_t* wt_full = bars + 1;                  // [kWtStages]  TMA -> convT issuer
  uint64_t* wt_empty = wt_full + kWtStages;      // [kWtStages]  convT MMAs done -> loader
  uint64_t* d_full = wt_empty + kWtStages;       //              convT MMAs done -> F builder
  uint64_t* d_empty = d_full + 1;                //              F builders (8 warps) have read D -> convT issuer
  uint64_t* f_full = d_empty + 1;                // [kFStages]   F builders (8 warps) -> head issuers
  uint64_t* f_empty = f_full + kFStages;         // [kFStages]   head MMAs done (2 issuers) -> F builders
  uint64_t* a2_full = f_empty + kFStages;        // [kA2Stages]  DEM operand builder -> head issuers
  uint64_t* a2_empty = a2_full + kA2Stages;      // [kA2Stages]  head MMAs done (2 issuers) -> DEM prefetch
  uint64_t* dem_full = a2_empty + kA2Stages;     // [kA2Stages]  DEM prefetch (32 cp.async arrivals) -> DEM builder
  uint64_t* slot_full = dem_full + kA2Stages;    // [kStrips][3] head MMAs done -> epilogue
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
      mbar_init(d_empty, kBuilders);
      for (int i = 0; i < kFStages; ++i) {
        mbar_init(&f_full[i], kBuilders);
        mbar_init(&f_empty[i], 2);
      }
      for (int i = 0; i < kA2Stages; ++i) {
        mbar_init(&a2_full[i], 1);
        mbar_init(&a2_empty[i], 2);
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
    // ===================== loader: head weights once; per valid input row 8 (kx block, K half) stages ===================
    if (lane == 0) {
      mbar_expect_tx(hw_full, kHwBytes);
      bulk_load_1d(smem_hw, p.hw, kHwBytes / 2, hw_full);
      bulk_load_1d(smem_hw + kHwBytes / 2, reinterpret_cast<const uint8_t*>(p.hw) + kHwBytes / 2, kHwBytes / 2, hw_full);
      int st = 0;
      uint32_t ph = 1;
      for (RowIter it(p); it.next();) {
        for (int i = 0; i < it.rows + 2; ++i) {
          const int y = it.y0 - 1 + i;
          if (y < 0 || y >= p.H) continue;
          for (int bk = 0; bk < 8; ++bk) {  // bk = block * 2 + K half
            wait_relaxed(&wt_empty[st], ph, p.wait_dem_ns);
            mbar_expect_tx(&wt_full[st], kWtStage);
            uint8_t* dst = smem_wt + st * kWtStage;
            bulk_load_1d(dst, reinterpret_cast<const uint8_t*>(p.wt) + ((size_t)(y & (kUp - 1)) * 8 + bk) * (2 * kWtPart), 2 * kWtPart, &wt_full[st]);
            tma_load_5d(dst + 2 * kWtPart, &tmL, &wt_full[st], 0, y / kUp, it.img, 2 * (bk & 1), 0);
            if (++st == kWtStages) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 24) {
    // ===================== convT MMA issuer: D_t[(kx, co), cell] for one row, three MMAs per (block, K half) ==========
    const uint32_t idesc = idesc_16(128, kCells, 1);
    const uint32_t wt0 = smem_u32(smem_wt);
    const bool leader = elect_one();
    int st = 0;
    uint32_t ph = 0, dph = 1;
    for (RowIter it(p); it.next();) {
      for (int i = 0; i < it.rows + 2; ++i) {
        const int y = it.y0 - 1 + i;
        if (y < 0 || y >= p.H) continue;
        for (int bk = 0; bk < 8; ++bk) {
          mbar_wait(&wt_full[st], ph);
          if (bk == 0) mbar_wait(d_empty, dph);
          tc_fence_after();
          if (leader) {
            const uint32_t wbase = wt0 + st * kWtStage;
            const uint32_t a_hi = dlo(wbase, 2048), a_lo = dlo(wbase + kWtPart, 2048);                // [2 planes][128 rows][16 B]
            const uint32_t b_hi = dlo(wbase + 2 * kWtPart, kCells * 16), b_lo = dlo(wbase + 2 * kWtPart + kLPart, kCells * 16);
            const uint32_t dcol = tmem_base + kDcol + (bk >> 1) * kCells;
            umma2(dcol, a_hi, b_hi, desc_hi, idesc, (bk & 1) ? 1u : 0u);
            umma2(dcol, a_lo, b_hi, desc_hi, idesc, 1u);
            umma2(dcol, a_hi, b_lo, desc_hi, idesc, 1u);
            commit_to(smem_u32(&wt_empty[st]));
            if (bk == 7) commit_to(smem_u32(d_full));
          }
          __syncwarp();
          if (++st == kWtStages) { st = 0; ph ^= 1; }
        }
        dph ^= 1;
      }
    }
  } else if (warp >= 8 && warp < 8 + kBuilders) {
    // ===================== F builders: TMEM D_t -> scale, + bias, activation, split -> hi stage, then lo stage =========
    // warp (gb, q, ch) handles blocks 2 gb and 2 gb + 1 (kx = 4 b + q) for cells [16 ch, 16 ch + 16); lane == co, TMEM lane
    // quarter q == kx within the block
    const int gb = ((warp - 8) >> 2) & 1;
    const int ch = (warp - 8) >> 3;
    const int q = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + kDcol;
    const float bias = p.bias_t[lane];
    const float sc = p.scale_t_inv;
    int fs = 0;
    uint32_t fph = 1, dph = 0;
    for (RowIter it(p); it.next();) {
      for (int i = 0; i < it.rows + 2; ++i) {
        const int y = it.y0 - 1 + i;
        const int fs_a = fs;
        const uint32_t ph_a = fph;
        if (++fs == kFStages) { fs = 0; fph ^= 1; }
        const int fs_b = fs;
        const uint32_t ph_b = fph;
        if (++fs == kFStages) { fs = 0; fph ^= 1; }
        uint8_t* frow_a = smem_f + fs_a * kFRow;
        uint8_t* frow_b = smem_f + fs_b * kFRow;
        if (y < 0 || y >= p.H) {
          // zero padding row of the head convolution (both parts)
          wait_relaxed(&f_empty[fs_a], ph_a, p.wait_build_ns);
          for (int k = (warp - 8) * 32 + lane; k < kFRow / 16; k += kBuilders * 32) reinterpret_cast<uint4*>(frow_a)[k] = make_uint4(0, 0, 0, 0);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&f_full[fs_a]);
          wait_relaxed(&f_empty[fs_b], ph_b, p.wait_build_ns);
          for (int k = (warp - 8) * 32 + lane; k < kFRow / 16; k += kBuilders * 32) reinterpret_cast<uint4*>(frow_b)[k] = make_uint4(0, 0, 0, 0);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&f_full[fs_b]);
          continue;
        }
        wait_relaxed(d_full, dph, p.wait_build_ns);
        dph ^= 1;
        tc_fence_after();
        float v0[kCellsPerBuilder], v1[kCellsPerBuilder];
        tmem_ld16(lane_addr + (2 * gb) * kCells + ch * kCellsPerBuilder, v0);
        tmem_ld16(lane_addr + (2 * gb + 1) * kCells + ch * kCellsPerBuilder, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d_empty);
        // lanes 2k / 2k+1 hold channels 2k / 2k+1: one shuffle per cell lets the even lane store the channel pair of block
        // 2 gb and the odd lane that of block 2 gb + 1 as 32-bit words (32 lanes -> 32 distinct banks); the (hi, lo) halves
        // travel together through the shuffle
        const bool odd = lane & 1;
        const uint32_t off0 = (uint32_t)(lane >> 3) * kFPlane + (uint32_t)((lane & 7) >> 1) * 4 +
                              (uint32_t)(1 + 4 * (2 * gb + (odd ? 1 : 0)) + q + ch * kCellsPerBuilder * kUp) * 16;
        // The kernel's time follows the builders' instruction count (measured: +2 instructions per cell = +9 %), so a cell is
        // ~16 instructions: PACKED conversions (one F2FP for the two blocks' values; scalar F2F runs on a quarter-rate pipe), one
        // shuffle carrying the (hi, lo) the neighbouring lane needs, byte permutes instead of shifts and selects.  The even lane
        // stores the channel pair (2k, 2k + 1) of block 2 gb, the odd lane that of block 2 gb + 1.
        const uint32_t sel_send = odd ? 0x5410u : 0x7632u;
        const uint32_t sel_hi = odd ? 0x3254u : 0x5410u, sel_lo = odd ? 0x3276u : 0x7610u;
        uint32_t low[kCellsPerBuilder];
        __half2 hmax = __floats2half2_rn(0.0f, 0.0f);  // largest |hi| stored: Inf = a feature value that does not fit fp16 (see
                                                       // flag_unstorable in k_tc_conv.cu)
        wait_relaxed(&f_empty[fs_a], ph_a, p.wait_build_ns);
#pragma unroll
        for (int c = 0; c < kCellsPerBuilder; ++c) {
          const float fa = act_fn<ACT_T>(fmaf(v0[c], sc, bias), p.alpha_t), fb = act_fn<ACT_T>(fmaf(v1[c], sc, bias), p.alpha_t);
          const __half2 h = __floats2half2_rn(fa, fb);   // (block 2 gb, block 2 gb + 1), own channel
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(fa - hf.x, fb - hf.y);
#ifndef FSR_NO_BUILDER_CHECK
          hmax = __hmax2(hmax, __habs2(h));
#endif
          const uint32_t uh = *reinterpret_cast<const uint32_t*>(&h), ul = *reinterpret_cast<const uint32_t*>(&l);
          const uint32_t other = __shfl_xor_sync(0xffffffffu, __byte_perm(uh, ul, sel_send), 1);
          *reinterpret_cast<uint32_t*>(frow_a + off0 + c * (kUp * 16)) = __byte_perm(uh, other, sel_hi);
          low[c] = __byte_perm(ul, other, sel_lo);
        }
        if (__hisinf(__low2half(hmax)) || __hisinf(__high2half(hmax))) atomicOr(p.flags, FSR_FLAG_PRED_NONFINITE);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&f_full[fs_a]);
        wait_relaxed(&f_empty[fs_b], ph_b, p.wait_build_ns);
#pragma unroll
        for (int c = 0; c < kCellsPerBuilder; ++c) *reinterpret_cast<uint32_t*>(frow_b + off0 + c * (kUp * 16)) = low[c];
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&f_full[fs_b]);
      }
    }
  } else if (warp == 2) {
    // ===================== DEM prefetcher: fp32 halo row (514 px) -> smem ===================================
    // src.on: the row comes from the raw raster window of the tile (zero beyond the raster, as the reference's padding to
    // whole tiles); the operand builder normalises it.  Pixels outside the TILE are the head convolution's zero padding.
    int as = 0;
    uint32_t aph = 1;
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
        const uint32_t dst = smem_u32(smem_dem + as * kDemRow);
        wait_relaxed(&a2_empty[as], aph, p.wait_dem_ns);
        for (int k = lane; k < kRowPx; k += 32) {
          const int x = k - 1;
          const bool ok = yok && x >= 0 && x < x_end;
          const float* src = row + (ok ? x : 0);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + k * 4), "l"(src), "r"(ok ? 4 : 0) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&dem_full[as])) : "memory");
        if (++as == kA2Stages) { as = 0; aph ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================== DEM operand builder: [512 px][hi(-1,0,+1), lo(-1,0,+1), 0, 0] =====================
    int as = 0;
    uint32_t ph = 0;
    for (RowIter it(p); it.next();) {
      float p_clip = 0.f, dem_min = 0.f, range_f = 1.f;
      bool zero_out = false;
      if (p.src.on) dem_norm_spec(p.src.stats + (size_t)it.img * 3, p_clip, dem_min, range_f, zero_out);
      for (int i = 0; i < it.rows + 2; ++i) {
        wait_relaxed(&dem_full[as], ph, p.wait_dem_ns);
        if (p.src.on) {
          // raw raster values -> normalised DEM, in place; pixels outside the tile stay the convolution's zero padding
          const int y = it.y0 - 1 + i;
          const bool in_tile_row = y >= 0 && y < p.H;
          float* drow = reinterpret_cast<float*>(smem_dem + as * kDemRow);
          for (int k = lane; k < kRowPx; k += 32) {
            const bool in_tile = in_tile_row && k >= 1 && k <= kW;
            drow[k] = in_tile ? dem_normalise_px(drow[k], p.src, p_clip, dem_min, range_f, zero_out) : 0.0f;
          }
          __syncwarp();
        }
        // consecutive lanes take consecutive pixels: every LDS / STS.128 of the warp is conflict-free (a lane-owns-4-pixels
        // mapping made each 16-byte store a 16-way bank conflict: 12 % of the kernel's LSU wavefronts in the ncu capture)
#pragma unroll 4
        for (int g = 0; g < kW / 32; ++g) {
          const int px = g * 32 + lane;
          const float* drow = reinterpret_cast<const float*>(smem_dem + as * kDemRow) + px;  // halo index of px - 1
          const float d0 = drow[0], d1 = drow[1], d2 = drow[2];
          const __half2 h01 = __floats2half2_rn(d0, d1), h2z = __floats2half2_rn(d2, 0.0f);
          const float2 f01 = __half22float2(h01), f2z = __half22float2(h2z);
          const __half2 l01 = __floats2half2_rn(d0 - f01.x, d1 - f01.y), l2z = __floats2half2_rn(d2 - f2z.x, 0.0f);
          const uint32_t uh01 = *reinterpret_cast<const uint32_t*>(&h01), uh2 = *reinterpret_cast<const uint32_t*>(&h2z);
          const uint32_t ul01 = *reinterpret_cast<const uint32_t*>(&l01), ul2 = *reinterpret_cast<const uint32_t*>(&l2z);
          uint4 v;
          v.x = uh01;                              // hi(-1), hi(0)
          v.y = __byte_perm(uh2, ul01, 0x5410);    // hi(+1), lo(-1)
          v.z = __byte_perm(ul01, ul2, 0x5432);    // lo(0), lo(+1)
          v.w = 0u;
          reinterpret_cast<uint4*>(smem_a2 + as * kA2Row)[px] = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[as]);
        if (++as == kA2Stages) { as = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 25) {
    // ===================== head MMA issuers (warp 1: strips 0-1, warp 25: strips 2-3) ====================================
    const uint32_t idesc0 = idesc_16(128, 0, 1);  // + (N >> 3) << 17
    const int s_begin = warp == 1 ? 0 : 2;
    const uint32_t idesc96 = idesc0 + (3u << 19), idesc32 = idesc0 + (1u << 19);
    mbar_wait(hw_full, 0);
    const uint32_t b_lo0 = dlo(smem_u32(smem_hw), kHwBlocks * kC * 16);
    const uint32_t f_lo0 = dlo(smem_u32(smem_f), kFPlane);
    const uint32_t zero_addr = smem_u32(smem_zero);
    const uint32_t bar_slot_full = smem_u32(slot_full), bar_f_empty = smem_u32(f_empty), bar_a2_empty = smem_u32(a2_empty);
    const bool leader = elect_one();
    constexpr uint32_t kB = kHwStep >> 4, kP = (2 * kFPlane) >> 4;
    int fs = 0, as = 0;
    uint32_t fph = 0, aph = 0;
    int go = 0;  // output rows of earlier items (ring position base)
    for (RowIter it(p); it.next(); go += it.rows) {
      for (int i = 0; i < it.rows + 2; ++i) {
        const int R = go + i;                 // ring index of the output row this input row opens
        const int r0 = (R + 1) % 3;           // slot of output row i - 2 ( (R - 2) mod 3 )
        const int t0 = (3 - r0) % 3;          // first weight block so that slot order matches [ky2|ky1|ky0] rotation
        const bool steady = i >= 2 && i < it.rows;
        // ---- stage A: F_hi x [W_hi, W_lo] ----
        mbar_wait(&f_full[fs], fph);
        tc_fence_after();
        {
          const uint32_t f_row = f_lo0 + fs * (kFRow >> 4);
          for (int s = s_begin; s < s_begin + 2; ++s) {
            if (i < it.rows) mbar_wait(&slot_empty[s * 3 + R % 3], ((R / 3) & 1) ^ 1);  // read + re-zeroed by the epilogue
            tc_fence_after();
            const uint32_t a = f_row + s * 128;   // 128 pixels x 16 B >> 4
            const uint32_t d = tmem_base + s * (3 * kC);
            if (leader) {
              if (steady) {
                const uint32_t b = b_lo0 + t0 * kC;
#pragma unroll
                for (int k = 0; k < 6; ++k) {  // k = 2 kx + j
                  const uint32_t ak = a + (k >> 1) + (k & 1) * kP;
                  umma2(d, ak, b + k * kB, desc_hi, idesc96, 1u);
                  umma2(d, ak, b + (6 + k) * kB, desc_hi, idesc96, 1u);
                }
              } else {
                // item borders: one N = 32 MMA set per existing target row j = i - 2 + t (weight block t <-> ky = 2 - t)
                for (int t = 0; t < 3; ++t) {
                  const int j = i - 2 + t;
                  if (j < 0 || j >= it.rows) continue;
                  const uint32_t dj = d + ((go + j) % 3) * kC;
                  const uint32_t b = b_lo0 + t * kC;
#pragma unroll
                  for (int k = 0; k < 6; ++k) {
                    const uint32_t ak = a + (k >> 1) + (k & 1) * kP;
                    umma2(dj, ak, b + k * kB, desc_hi, idesc32, 1u);
                    umma2(dj, ak, b + (6 + k) * kB, desc_hi, idesc32, 1u);
                  }
                }
              }
            }
            __syncwarp();
          }
          if (leader) commit_to(bar_f_empty + fs * 8);
          __syncwarp();
          if (++fs == kFStages) { fs = 0; fph ^= 1; }
        }
        // ---- stage B: F_lo x W_hi + the two DEM steps; closes output row i - 2 ----
        mbar_wait(&f_full[fs], fph);
        mbar_wait(&a2_full[as], aph);
        tc_fence_after();
        {
          const uint32_t f_row = f_lo0 + fs * (kFRow >> 4);
          for (int s = s_begin; s < s_begin + 2; ++s) {
            const uint32_t a = f_row + s * 128;
            const uint32_t a2_addr = smem_u32(smem_a2) + as * kA2Row + s * 2048;
            const uint32_t a2 = dlo(a2_addr, zero_addr - a2_addr);
            const uint32_t d = tmem_base + s * (3 * kC);
            if (leader) {
              if (steady) {
                const uint32_t b = b_lo0 + t0 * kC;
#pragma unroll
                for (int k = 0; k < 6; ++k) umma2(d, a + (k >> 1) + (k & 1) * kP, b + k * kB, desc_hi, idesc96, 1u);
                umma2(d, a2, b + 12 * kB, desc_hi, idesc96, 1u);
                umma2(d, a2, b + 13 * kB, desc_hi, idesc96, 1u);
              } else {
                for (int t = 0; t < 3; ++t) {
                  const int j = i - 2 + t;
                  if (j < 0 || j >= it.rows) continue;
                  const uint32_t dj = d + ((go + j) % 3) * kC;
                  const uint32_t b = b_lo0 + t * kC;
#pragma unroll
                  for (int k = 0; k < 6; ++k) umma2(dj, a + (k >> 1) + (k & 1) * kP, b + k * kB, desc_hi, idesc32, 1u);
                  umma2(dj, a2, b + 12 * kB, desc_hi, idesc32, 1u);
                  umma2(dj, a2, b + 13 * kB, desc_hi, idesc32, 1u);
                }
              }
              if (i >= 2) commit_to(bar_slot_full + (s * 3 + r0) * 8);
            }
            __syncwarp();
          }
          if (leader) {
            commit_to(bar_f_empty + fs * 8);
            commit_to(bar_a2_empty + as * 8);
          }
          __syncwarp();
          if (++fs == kFStages) { fs = 0; fph ^= 1; }
          if (++as == kA2Stages) { as = 0; aph ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 4-7): one warpgroup drains all four strips ================
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sc = p.scale_h_inv;
    int go = 0;
    for (RowIter it(p); it.next(); go += it.rows) {
      for (int j = 0; j < it.rows; ++j) {
        const int r = go + j;
        const int slot = r % 3;
        const uint32_t par = (r / 3) & 1;
        for (int s = 0; s < kStrips; ++s) {
          wait_relaxed(&slot_full[s * 3 + slot], par, p.wait_epi_ns);
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
            o0 = fmaf(act_fn<ACT>(fmaf(v[c], sc, p.bias_h[c]), p.alpha_h), p.w2[c], o0);
            o1 = fmaf(act_fn<ACT>(fmaf(v[c + 1], sc, p.bias_h[c + 1]), p.alpha_h), p.w2[c + 1], o1);
            o2 = fmaf(act_fn<ACT>(fmaf(v[c + 2], sc, p.bias_h[c + 2]), p.alpha_h), p.w2[c + 2], o2);
            o3 = fmaf(act_fn<ACT>(fmaf(v[c + 3], sc, p.bias_h[c + 3]), p.alpha_h), p.w2[c + 3], o3);
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

bool fused_x3_ok(int H, int W, int lr_h, int lr_w, int cin_t, int cout_t, int k_t, int cmid, int ksz) {
  return W == kW && lr_w == kCells && k_t == kUp && H == lr_h * kUp && cin_t == 32 && cout_t == kC && cmid == kC && ksz == 3;
}

size_t fused_x3_hw_elems() { return (size_t)kHwBytes / 2; }
size_t fused_x3_wt_elems() { return (size_t)kUp * 8 * 2 * kWtPart / 2; }

// head weights w [3][3][33][32] scaled by `scale` -> [14 K steps][2 planes][160 columns][8]; column n = blk * 32 + co with
// ky = 2 - blk % 3.  Steps 0-5: hi parts of (kx, 16-channel slice); 6-11: their lo parts; 12: DEM step 1 (k 0-2 and k 3-5 =
// hi(Wdem[ky][kx]): multiply dem_hi and dem_lo); 13: DEM step 2 (k 0-2 = lo(Wdem[ky][kx]): multiplies dem_hi).
void fused_x3_pack_head(const float* w, float scale, uint16_t* dst) {
  const int cin_real = 33;
  const size_t step = (size_t)kHwStep / 2;
  for (int kx = 0; kx < 3; ++kx)
    for (int j = 0; j < 2; ++j)
      for (int pl = 0; pl < 2; ++pl)
        for (int n = 0; n < kHwBlocks * kC; ++n)
          for (int e = 0; e < 8; ++e) {
            const int ky = 2 - (n / kC) % 3, co = n % kC, ci = j * 16 + pl * 8 + e;
            const size_t pos = (size_t)(2 * kx + j) * step + ((size_t)pl * kHwBlocks * kC + n) * 8 + e;
            split_weight(w[(((size_t)ky * 3 + kx) * cin_real + ci) * kC + co] * scale, dst[pos], dst[pos + 6 * step]);
          }
  for (int n = 0; n < kHwBlocks * kC; ++n)
    for (int e = 0; e < 6; ++e) {
      const int ky = 2 - (n / kC) % 3, co = n % kC;
      uint16_t hi, lo;
      split_weight(w[(((size_t)ky * 3 + (e % 3)) * cin_real + 32) * kC + co] * scale, hi, lo);
      dst[12 * step + (size_t)n * 8 + e] = hi;
      if (e < 3) dst[13 * step + (size_t)n * 8 + e] = lo;
    }
}

// convT weights w [16 ky][16 kx][32 ci][32 co] scaled by `scale` ->
// [ky][4 blocks][2 K halves][2 parts][2 K planes][128 rows = (kx % 4) * 32 + co][8 ci]
void fused_x3_pack_convt(const float* w, float scale, uint16_t* dst) {
  const size_t part = (size_t)kWtPart / 2;
  for (int ky = 0; ky < kUp; ++ky)
    for (int b = 0; b < 4; ++b)
      for (int kh = 0; kh < 2; ++kh)
        for (int pl = 0; pl < 2; ++pl)
          for (int row = 0; row < 128; ++row)
            for (int e = 0; e < 8; ++e) {
              const int kx = 4 * b + row / kC, co = row % kC, ci = kh * 16 + pl * 8 + e;
              const size_t pos = (((size_t)ky * 4 + b) * 2 + kh) * 2 * part + ((size_t)pl * 128 + row) * 8 + e;
              split_weight(w[(((size_t)ky * kUp + kx) * 32 + ci) * kC + co] * scale, dst[pos], dst[pos + part]);
            }
}

void launch_fused_x3(const __nv_bfloat16* lr, long long lr_plane, const __nv_bfloat16* wt_pack, const float* bias_t, float scale_t_inv,
                     int act_t, float alpha_t, const __nv_bfloat16* hw_pack, const float* bias_h, float scale_h_inv, const float* w2,
                     const float* b2, int act_h, float alpha_h, const float* dem, const DemSource& src, float* pred_m, float* pred_norm,
                     int n_img, int H, float max_depth, float denom, int n_sms, unsigned* flags, cudaStream_t s) {
  X3Params p{};
  p.H = H;
  p.N = n_img;
  p.total_rows = (long long)n_img * H;
  p.alpha_t = alpha_t;
  p.alpha_h = alpha_h;
  p.scale_t_inv = scale_t_inv;
  p.scale_h_inv = scale_h_inv;
  // FSR_X3_WAIT_NS="epilogue,builders,dem" overrides the suspend-time hints (A/B measurements)
  p.wait_epi_ns = 400u;
  p.wait_build_ns = 200u;
  p.wait_dem_ns = 400u;
  if (const char* e = getenv("FSR_X3_WAIT_NS")) {
    unsigned a = 0, b = 0, c = 0;
    if (sscanf(e, "%u,%u,%u", &a, &b, &c) == 3 && a && b && c) {
      p.wait_epi_ns = a;
      p.wait_build_ns = b;
      p.wait_dem_ns = c;
    }
  }
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
    p.bias_h[c] = bias_h ? bias_h[c] : 0.0f;
    p.w2[c] = w2[c];
  }
  p.b2 = b2 ? b2[0] : 0.0f;
  // L as TMA source: one LR row of 32 cells, one K half (2 channel planes), both parts -> [part][2 planes][32 cells][8]
  CUtensorMap mL = make_cp8_wide_tensor_map(lr, kCells, H / kUp, n_img, 4, lr_plane, kCells, 1, 1, 2, 2);
  const int grid = p.total_rows < n_sms ? (int)p.total_rows : n_sms;
  auto go = [&](auto kernel) {
    FSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    kernel<<<grid, kThreads, kSmemBytes, s>>>(mL, p);
  };
  auto by_act_t = [&](auto ah) {
    constexpr int AH = decltype(ah)::value;
    if (act_t == FSR_ACT_RELU) go(fused_hr_x3_kernel<AH, FSR_ACT_RELU>);
    else if (act_t == FSR_ACT_LEAKY) go(fused_hr_x3_kernel<AH, FSR_ACT_LEAKY>);
    else go(fused_hr_x3_kernel<AH, FSR_ACT_NONE>);
  };
  if (act_h == FSR_ACT_RELU) by_act_t(std::integral_constant<int, FSR_ACT_RELU>{});
  else if (act_h == FSR_ACT_LEAKY) by_act_t(std::integral_constant<int, FSR_ACT_LEAKY>{});
  else by_act_t(std::integral_constant<int, FSR_ACT_NONE>{});
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
