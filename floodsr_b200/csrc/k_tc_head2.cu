// Fused head of ResUNet_16x_DEM on tcgen05 (tensor-core precisions): conv3x3 over concat(features, dem_hr)
// -> activation -> conv1x1 -> invert_depth_log1p.  ~81 % of the FLOPs behind `session.run`
// (floodsr/engine/ort.py:193) plus the invert_depth_log1p_np call after it (ort.py:196, preprocessing.py:154-164).
//
// Persistent kernel, one CTA per SM, work item = 128-pixel-wide strip of RB output rows of one tile.
// For every *input* row r of a strip one TMEM accumulator slot D'_r[x, (ky, co)] (N = 3*32 = 96 columns) is
// produced from that row alone:
//     D'_r[x, (ky, co)] = sum_{kx, ci} F[r, x + kx - 1, ci] * W[ky, kx, ci, co]          (6 MMAs, K = 3 x 32)
//                       + sum_{kx}     dem[r, x + kx - 1]   * Wdem[ky, kx, co]  (+ bias) (1 MMA,  K = 16)
// The kx taps are start-address offsets into the row's halo buffer (no-swizzle CP8 operand: 16 B per pixel per
// 8-channel plane), so each feature row is fetched once (TMA, zero-filled outside the tile = 'same' padding).
// The 1-channel DEM operand is built in shared memory by a dedicated warp as (hi, lo) 16-bit pairs, so the DEM
// term keeps ~fp32 input precision; its K slot 6 is a constant 1 that carries the bias.
// An output row then only needs same-lane TMEM reads:
//     conv[y, x, co] = D'_{y-1}[x, (0, co)] + D'_y[x, (1, co)] + D'_{y+1}[x, (2, co)]
// (carried as register partials so every slot is read once) followed by activation, the 1x1 projection, the log1p inversion and coalesced fp32 stores.
// Operand traffic per FLOP is ~2x lower than a direct N = 32 formulation (A tiles are shared by 96 columns).
//
// Warp roles (256 threads): 0 TMA producer | 1 MMA issuer | 2 DEM prefetcher | 3 DEM-operand builder | 4-7 epilogue.
// All hand-offs are mbarriers; accumulator slots form a ring of 5, each read once by the epilogue.
#include <stdlib.h>

#include "fsr_engine.cuh"
#include "tc_common.cuh"

namespace fsr {

using namespace tc;

CUtensorMap make_cp8_tensor_map(const void* base, int W, int H, int N, int chunks, long long plane, int bw, int bh, int bn, int kc);
CUtensorMap make_f32_tensor_map_3d(const void* base, int W, int H, int N, int bw);

namespace {

constexpr int kCmid = 32;
constexpr int kN = 3 * kCmid;                     // accumulator columns per slot
constexpr int kSlots = 5;
constexpr int kStages = 8;
constexpr int kRowPx = 130;                       // 128 + left/right halo pixel
constexpr int kPlaneBytes = kRowPx * 16;          // 2080
constexpr int kRowBytes = 4 * kPlaneBytes;        // 32 feature channels = 4 planes
constexpr int kWBytes = 3 * 2 * (2 * kN * 16);    // (kx, k-slice) x [2 planes][96][8]
constexpr int kW2Bytes = 2 * kN * 16;             // DEM/bias operand [2 planes][96][8]
constexpr int kA2Bytes = 128 * 16;                // DEM operand plane per stage
constexpr int kDemBytes = 640;                    // per-stage staging (128-byte aligned)
constexpr int kThreads = 256;
constexpr int kSmemBytes = 160 * 1024;            // > half an SM's shared memory: exactly one CTA per SM (TMEM is exclusive)

struct Head2Params {
  int H, W, N;          // HR tile extent and tiles in this launch
  int rb;               // output rows per work item
  int n_items;
  int act;
  float alpha;
  int half;             // 16-bit format: 0 bf16, 1 fp16
  float max_depth, denom;
  const __nv_bfloat16* wpack;   // feature weights [kx][kslice][2][96][8] followed by the DEM/bias operand [2][96][8]
  const float* dem;     // [N][H][W] normalised DEM
  float* pred_m;        // [N][H][W]
  float* pred_norm;     // [N][H][W] or nullptr
  float w2[kCmid];      // 1x1 projection
  float b2;
  int mode;             // debug switches
  long long* stats;     // [16] per-role wait cycles of CTA 0 (debug)
  unsigned* dbg;        // mapped host word: which wait timed out (debug builds of the pipeline)
};

#define TWAIT(cnt, ...)                  \
  do {                                    \
    long long _t0 = clock64();            \
    wait_tag(__VA_ARGS__);                \
    cnt += clock64() - _t0;               \
  } while (0)

__device__ __forceinline__ void wait_tag(uint64_t* bar, uint32_t parity, unsigned* dbg, unsigned tag) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      if (dbg) {
        atomicCAS(dbg, 0u, tag);
        __threadfence_system();
      }
      __trap();
    }
  }
}

__device__ __forceinline__ uint16_t to16(float v, int half) {
  if (half) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float from16(uint16_t u, int half) {
  if (half) return __half2float(__ushort_as_half(u));
  return __bfloat162float(__ushort_as_bfloat16(u));
}

__global__ void __launch_bounds__(kThreads, 1)
head2_tc_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ Head2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem_w = smem_raw;                                   // 18432 B
  uint8_t* smem_w2 = smem_w + kWBytes;                          // 3072 B
  uint8_t* smem_rows = smem_w2 + kW2Bytes;                      // kStages x 8320 B (128-B aligned)
  uint8_t* smem_a2 = smem_rows + kStages * kRowBytes;           // kStages x 2048 B
  uint8_t* smem_zero = smem_a2 + kStages * kA2Bytes;            // 2048 B of zeros (upper K plane of the DEM operand)
  uint8_t* smem_dem = smem_zero + kA2Bytes;                     // kStages x 640 B fp32 DEM halo rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dem + kStages * kDemBytes);
  uint64_t* w_full = bars;
  uint64_t* row_full = bars + 1;                 // [kStages]  TMA -> MMA
  uint64_t* a2_full = row_full + kStages;        // [kStages]  DEM builder -> MMA
  uint64_t* row_empty = a2_full + kStages;       // [kStages]  MMA -> producers
  uint64_t* slot_full = row_empty + kStages;     // [kSlots]   MMA -> epilogue
  uint64_t* slot_empty = slot_full + kSlots;     // [kSlots]   epilogue -> MMA (12 arrivals: 3 reading rows x 4 warps)
  uint64_t* dem_full = slot_empty + kSlots;      // [kStages]  DEM prefetcher (32 cp.async arrivals) -> builder
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dem_full + kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int segs = p.W / 128;
  const int rblocks = p.H / p.rb;
  const int n_in = p.rb + 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmF);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int i = 0; i < kStages; ++i) {
        mbar_init(&row_full[i], 1);
        mbar_init(&a2_full[i], 1);
        mbar_init(&row_empty[i], 1);
        mbar_init(&dem_full[i], 32);
      }
      for (int i = 0; i < kSlots; ++i) {
        mbar_init(&slot_full[i], 1);
        mbar_init(&slot_empty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp == 2) {
    for (int i = lane; i < kA2Bytes / 16; i += 32) reinterpret_cast<uint4*>(smem_zero)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one halo row per input row ===================
    if (lane == 0) {
      mbar_expect_tx(w_full, kWBytes + kW2Bytes);
      bulk_load_1d(smem_w, p.wpack, kWBytes + kW2Bytes, w_full);
      int g = 0;  // running input-row counter of this CTA
      long long c_w0 = 0, c_t0 = clock64();
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int t = item;
        const int xs = t % segs;
        t /= segs;
        const int rbk = t % rblocks;
        const int img = t / rblocks;
        for (int i = 0; i < n_in; ++i, ++g) {
          const int s = g % kStages;
          TWAIT(c_w0, &row_empty[s], ((g / kStages) & 1) ^ 1, p.dbg, 0x100000u | g);
          if (p.mode & 4) {
            mbar_arrive(&row_full[s]);
          } else {
            mbar_expect_tx(&row_full[s], kRowBytes);
            tma_load_5d(smem_rows + s * kRowBytes, &tmF, &row_full[s], 0, xs * 128 - 1, rbk * p.rb - 1 + i, img, 0);
          }
        }
      }
      if (p.stats && blockIdx.x == 0) { p.stats[0] = c_w0; p.stats[1] = clock64() - c_t0; p.stats[15] = g; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 6 feature MMAs + 1 DEM/bias MMA (N = 96) per input row ============
    // The loop is warp-uniform (all lanes wait, one elected lane issues) and every descriptor is a precomputed
    // base plus a 16-byte-unit offset: a divergent single-lane loop that rebuilds descriptors costs ~130
    // cycles per MMA, more than twice the MMA itself (56 cycles at N = 96).
    const uint32_t idesc = idesc_16(128, kN, p.half);
    wait_tag(w_full, 0, p.dbg, 0x200000u);
    uint64_t db[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) db[k] = smem_desc_kmajor(smem_u32(smem_w) + k * (2 * kN * 16), kN * 16, 128);
    const uint64_t db2 = smem_desc_kmajor(smem_u32(smem_w2), kN * 16, 128);
    const uint64_t da_base = smem_desc_kmajor(smem_u32(smem_rows), kPlaneBytes, 128);
    const uint32_t zero_addr = smem_u32(smem_zero);
    const uint32_t a2_base = smem_u32(smem_a2);
    long long c_m0 = 0, c_m1 = 0, c_m2 = 0, c_mt = clock64();
    int g = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int i = 0; i < n_in; ++i, ++g) {
        const int s = g % kStages;
        const int slot = g % kSlots;
        TWAIT(c_m0, &slot_empty[slot], ((g / kSlots) & 1) ^ 1, p.dbg, 0x300000u | g);
        TWAIT(c_m1, &row_full[s], (g / kStages) & 1, p.dbg, 0x400000u | g);
        if (!(p.mode & 1)) TWAIT(c_m2, &a2_full[s], (g / kStages) & 1, p.dbg, 0x500000u | g);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + slot * kN;
        const uint64_t da_row = da_base + (uint64_t)((s * kRowBytes) >> 4);
        const uint32_t a2_addr = a2_base + s * kA2Bytes;
        const uint64_t da2 = smem_desc_kmajor(a2_addr, zero_addr - a2_addr, 128);
        if (elect_one()) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            if (p.mode & 64) break;
#pragma unroll
            for (int j = 0; j < 2; ++j)
              umma_bf16(d_addr, da_row + (uint64_t)((j * 2 * kPlaneBytes + kx * 16) >> 4), db[kx * 2 + j], idesc, (kx | j) ? 1u : 0u);
          }
          if (!(p.mode & 1)) umma_bf16(d_addr, da2, db2, idesc, 1u);
          umma_commit(&row_empty[s]);
          umma_commit(&slot_full[slot]);
        }
        __syncwarp();
      }
    }
    if (p.stats && blockIdx.x == 0 && lane == 0) { p.stats[2] = c_m0; p.stats[3] = c_m1; p.stats[4] = c_m2; p.stats[5] = clock64() - c_mt; }
  } else if (warp == 2 && !(p.mode & 16)) {
    // ===================== DEM prefetcher: fp32 halo rows -> smem, up to kStages rows ahead =================
    // 4-byte cp.async (zero-filled outside the tile); completion is signalled straight to the builder's mbarrier,
    // so this warp never waits for memory.
    long long c_p0 = 0, c_pt = clock64();
    int g = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int t = item;
      const int xs = t % segs;
      t /= segs;
      const int rbk = t % rblocks;
      const int img = t / rblocks;
      for (int i = 0; i < n_in; ++i, ++g) {
        const int s = g % kStages;
        const int y = rbk * p.rb - 1 + i;
        const bool yok = y >= 0 && y < p.H;
        const float* row = p.dem + ((size_t)img * p.H + (yok ? y : 0)) * p.W;
        const uint32_t dst = smem_u32(smem_dem + s * kDemBytes);
        TWAIT(c_p0, &row_empty[s], ((g / kStages) & 1) ^ 1, p.dbg, 0x800000u | g);
        for (int k = lane; k < kRowPx; k += 32) {
          const int x = xs * 128 - 1 + k;
          const bool ok = yok && x >= 0 && x < p.W;
          const float* src = row + (ok ? x : 0);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + k * 4), "l"(src), "r"(ok ? 4 : 0) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&dem_full[s])) : "memory");
      }
    }
    if (p.stats && blockIdx.x == 0 && lane == 0) { p.stats[6] = c_p0; p.stats[7] = clock64() - c_pt; }
  } else if (warp == 3 && !(p.mode & 16)) {
    // ===================== DEM operand builder: [128 px][hi(-1,0,+1), lo(-1,0,+1), 1, 0] ====================
    long long c_b0 = 0, c_bt = clock64();
    int g = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int i = 0; i < n_in; ++i, ++g) {
        const int s = g % kStages;
        TWAIT(c_b0, &dem_full[s], (g / kStages) & 1, p.dbg, 0x600000u | g);
        const float* drow = reinterpret_cast<const float*>(smem_dem + s * kDemBytes) + lane * 4;
        float d[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) d[k] = drow[k];
        uint16_t hi[6], lo[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          hi[k] = to16(d[k], p.half);
          lo[k] = to16(d[k] - from16(hi[k], p.half), p.half);
        }
        const uint16_t one = to16(1.0f, p.half);
        uint4* dst = reinterpret_cast<uint4*>(smem_a2 + s * kA2Bytes) + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = (uint32_t)hi[q] | ((uint32_t)hi[q + 1] << 16);
          v.y = (uint32_t)hi[q + 2] | ((uint32_t)lo[q] << 16);
          v.z = (uint32_t)lo[q + 1] | ((uint32_t)lo[q + 2] << 16);
          v.w = (uint32_t)one;
          dst[q] = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[s]);
      }
    }
    if (p.stats && blockIdx.x == 0 && lane == 0) { p.stats[8] = c_b0; p.stats[9] = clock64() - c_bt; }
  } else if (warp >= 4) {
    // ===================== epilogue: one warpgroup, every accumulator slot is read exactly once ============
    // Input row i contributes its ky=0 / 1 / 2 column groups to output rows i, i-1, i-2 (item-local).  The two
    // unfinished output rows are carried as register partials (pa: needs ky=2 next, pb: needs ky=1 then ky=2),
    // so a slot is released right after its single read and the MMA warp can run kSlots-1 rows ahead.
    const int q = warp & 3;            // TMEM lane quarter
    const int m = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    long long c_e0 = 0, c_et = clock64();
    int g = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int t = item;
      const int xs = t % segs;
      t /= segs;
      const int rbk = t % rblocks;
      const int img = t / rblocks;
      const int x = xs * 128 + m;
      float pa[kCmid], pb[kCmid];
      for (int i = 0; i < n_in; ++i, ++g) {
        const int slot = g % kSlots;
        TWAIT(c_e0, &slot_full[slot], (g / kSlots) & 1, p.dbg, 0x700000u | g);
        tc_fence_after();
        const uint32_t taddr = lane_addr + slot * kN;
        float out = p.b2;
        {
          float v[32];
          tmem_ld32(taddr + 2 * kCmid, v);  // ky = 2 -> completes output row i-2
          tmem_ld_wait();
          if (i >= 2 && !(p.mode & 2)) {
#pragma unroll
            for (int c = 0; c < kCmid; ++c) out = fmaf(apply_act(pa[c] + v[c], p.act, p.alpha), p.w2[c], out);
          }
        }
        {
          float v[32];
          tmem_ld32(taddr + kCmid, v);      // ky = 1 -> second contribution of output row i-1
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < kCmid; ++c) pa[c] = pb[c] + v[c];
        }
        tmem_ld32(taddr, pb);               // ky = 0 -> first contribution of output row i
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[slot]);
        if (i >= 2 && !(p.mode & 2)) {
          const int y = rbk * p.rb + i - 2;
          const size_t off = ((size_t)img * p.H + y) * p.W + x;
          if (p.pred_norm) p.pred_norm[off] = out;
          const float yn = fminf(fmaxf(out, 0.0f), 1.0f);
          p.pred_m[off] = fminf(fmaxf(expm1f(__fmul_rn(yn, p.denom)), 0.0f), p.max_depth);
        }
      }
    }
    if (p.stats && blockIdx.x == 0 && warp == 4 && lane == 0) { p.stats[10] = c_e0; p.stats[11] = clock64() - c_et; }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Packed operands of the head (host side, built once per engine):
//   [kx][k-slice j][plane pl][n = ky*32 + co][8]  feature weights W[ky][kx][ci = j*16 + pl*8 + e][co]
//   [plane][n][8]                                  DEM/bias operand: k 0-2 = Wdem[ky][kx], k 3-5 = the same (lo part),
//                                                  k 6 = bias[co] for ky == 1, everything else 0
size_t head2_pack_elems() { return (size_t)(kWBytes + kW2Bytes) / 2; }

void head2_pack(const float* w /* [3][3][33][32] */, const float* bias, uint16_t* dst, uint16_t (*cvt)(float)) {
  const int cin_real = 33;
  size_t pos = 0;
  for (int kx = 0; kx < 3; ++kx)
    for (int j = 0; j < 2; ++j)
      for (int pl = 0; pl < 2; ++pl)
        for (int n = 0; n < kN; ++n)
          for (int e = 0; e < 8; ++e, ++pos) {
            const int ky = n / kCmid, co = n % kCmid, ci = j * 16 + pl * 8 + e;
            dst[pos] = cvt(w[(((size_t)ky * 3 + kx) * cin_real + ci) * kCmid + co]);
          }
  for (int pl = 0; pl < 2; ++pl)
    for (int n = 0; n < kN; ++n)
      for (int e = 0; e < 8; ++e, ++pos) {
        const int ky = n / kCmid, co = n % kCmid;
        float v = 0.0f;
        if (pl == 0 && e < 6) v = w[(((size_t)ky * 3 + (e % 3)) * cin_real + 32) * kCmid + co];
        if (pl == 0 && e == 6 && ky == 1 && bias) v = bias[co];
        dst[pos] = cvt(v);
      }
}

void launch_head2_tc(const __nv_bfloat16* feat, long long plane, const __nv_bfloat16* wpack, const float* w2, const float* b2,
                     const float* dem, float* pred_m, float* pred_norm, int n_img, int H, int W, int cin, int cmid, int ksz,
                     int act, float alpha, float max_depth, float denom, int half, int n_sms, cudaStream_t s) {
  FSR_REQUIRE(cin == 32 && cmid == kCmid && ksz == 3, "head tensor-core path is specialised for 32 -> 32 channels, 3x3");
  FSR_REQUIRE(W % 128 == 0 && H % 32 == 0, "head tensor-core path needs W % 128 == 0 and H % 32 == 0");
  Head2Params p{};
  p.H = H; p.W = W; p.N = n_img;
  p.rb = 32;
  p.n_items = n_img * (W / 128) * (H / p.rb);
  p.act = act;
  p.alpha = alpha;
  p.half = half;
  p.max_depth = max_depth;
  p.denom = denom;
  p.wpack = wpack;
  p.dem = dem;
  p.pred_m = pred_m;
  p.pred_norm = pred_norm;
  for (int c = 0; c < kCmid; ++c) p.w2[c] = w2[c];
  p.b2 = b2 ? b2[0] : 0.0f;
  p.mode = getenv("FSR_HEAD_MODE") ? atoi(getenv("FSR_HEAD_MODE")) : 0;
  static unsigned* dbg_host = nullptr;
  static unsigned* dbg_dev = nullptr;
  if (!dbg_host) {
    FSR_CUDA(cudaHostAlloc(&dbg_host, sizeof(unsigned), cudaHostAllocMapped));
    *dbg_host = 0;
    FSR_CUDA(cudaHostGetDevicePointer(&dbg_dev, dbg_host, 0));
    static unsigned** keep = &dbg_host;
    atexit([]() {
      if (**keep) fprintf(stderr, "[floodsr_b200] head2 pipeline wait timed out: tag 0x%x\n", **keep);
    });
  }
  if (*dbg_host) fprintf(stderr, "[floodsr_b200] head2 pipeline wait timed out earlier: tag 0x%x\n", *dbg_host);
  p.dbg = dbg_dev;
  static long long* d_stats = nullptr;
  static int stat_calls = 0;
  if (getenv("FSR_HEAD_STATS")) {
    if (!d_stats) { FSR_CUDA(cudaMalloc(&d_stats, 16 * sizeof(long long))); FSR_CUDA(cudaMemset(d_stats, 0, 16 * sizeof(long long))); }
    p.stats = d_stats;
  }
  CUtensorMap mF = make_cp8_tensor_map(feat, W, H, n_img, cin / 8, plane, kRowPx, 1, 1, cin / 8);
  static bool attr = false;
  if (!attr) { FSR_CUDA(cudaFuncSetAttribute(head2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes)); attr = true; }
  const int grid = p.n_items < n_sms ? p.n_items : n_sms;
  head2_tc_kernel<<<grid, kThreads, kSmemBytes, s>>>(mF, p);
  FSR_LAUNCH_CHECK();
  if (p.stats && ++stat_calls == 20) {
    long long h[16];
    FSR_CUDA(cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[head2 stats, CTA 0, %lld rows] producer: wait row_empty %lld of %lld | mma: slot_empty %lld row_full %lld a2_full %lld of %lld | "
            "demprefetch: row_empty %lld of %lld | builder: dem_full %lld of %lld | epilogue: slot_full %lld of %lld cycles\n",
            h[15], h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10], h[11]);
  }
}

}  // namespace fsr
