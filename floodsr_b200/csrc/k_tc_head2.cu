// Fused head of ResUNet_16x_DEM on tcgen05 (tensor-core precisions): conv3x3 over concat(features, dem_hr)
// -> activation -> conv1x1 -> invert_depth_log1p.  ~81 % of the FLOPs behind `session.run`
// (floodsr/engine/ort.py:193) plus the invert_depth_log1p_np call after it (ort.py:196, preprocessing.py:154-164).
//
// Persistent kernel, one CTA per SM running TWO independent row pipelines (8 warps each).  The launch's output
// rows (128-pixel-wide strips of every tile, flattened) are split into 2 * gridDim.x equal contiguous ranges, one
// per pipeline, so every SM gets the same number of rows; a range is walked as "items" = runs of rows inside one
// strip.  A single MMA-issuing warp cannot keep the tensor core busy at N = 96 (its per-row barrier polls and
// descriptor arithmetic take longer than the queued MMAs), so two issuers feed it alternately, each into its own
// half of TMEM; MMAs of one pipeline stay ordered, which keeps the result deterministic.
//
// Per pipeline, TMEM holds a ring of 8 output-row accumulators D_y[x, co] (128 lanes x 32 fp32 columns each; both
// rings together = all 512 columns).  An *input* row r feeds the three output rows y = r+1, r, r-1 (ky = 0, 1, 2) with ONE N = 96 MMA per
// K step, because their accumulators are adjacent in the ring and the weight operand is stored
// [ky = 2 | ky = 1 | ky = 0] along N:
//     D_{r-1} | D_r | D_{r+1}  +=  F[r, x + kx - 1, ci-slice] * [W(2,kx) | W(1,kx) | W(0,kx)]     (6 K steps)
//                              +=  demop[r, x]                * [Wd(2)   | Wd(1)   | Wd(0)]       (1 K step)
// The kx taps are start-address offsets into the row's halo buffer (no-swizzle CP8 operand: 16 B per pixel per
// 8-channel plane), so each feature row is fetched once (TMA, zero-filled outside the tile = 'same' padding) and
// the 3x3 sum happens inside the tensor core's fp32 accumulators.  Every MMA accumulates: the epilogue re-zeroes an
// accumulator (tcgen05.st) right after reading it; ring wrap-around splits a row's MMAs in two.
// The 1-channel DEM operand is built in shared memory by a dedicated warp as (hi, lo) 16-bit pairs so the DEM
// term keeps ~fp32 input precision; its K slots 6/7 are constant 1 and carry the bias as a (hi, lo) pair.
// When input row y+1 has been issued, D_y is complete: the epilogue reads its 32 columns once, applies the
// activation, the 1x1 projection and the log1p inversion in fp32 and stores coalesced rows.
//
// Warp roles per pipeline: 0 TMA producer | 1 MMA issuer | 2 DEM prefetcher | 3 DEM-operand builder | 4-7 epilogue
// (one warp per TMEM lane quarter).  All hand-offs are mbarriers.
#include <stdlib.h>

#include "fsr_engine.cuh"
#include "tc_common.cuh"

namespace fsr {

using namespace tc;

CUtensorMap make_cp8_tensor_map(const void* base, int W, int H, int N, int chunks, long long plane, int bw, int bh, int bn, int kc);

namespace {

constexpr int kCmid = 32;
constexpr int kNfull = 3 * kCmid;                 // widest MMA: three adjacent output-row accumulators
constexpr int kPipes = 2;                         // independent row pipelines per CTA (each with its own MMA-issuing warp)
constexpr int kSlots = 8;                         // accumulator ring per pipeline: 8 x 32 fp32 columns = half the TMEM
constexpr int kStages = 9;                        // input rows in flight per pipeline (18 per SM, ~150 KB: HBM latency)
constexpr int kRowPx = 130;                       // 128 + left/right halo pixel
constexpr int kPlaneBytes = kRowPx * 16;          // 2080
constexpr int kRowBytes = 4 * kPlaneBytes;        // 32 feature channels = 4 planes
constexpr int kWStep = 2 * kNfull * 16;           // one K step of the weight operand: [2 planes][96][8]
constexpr int kWBytes = 6 * kWStep;               // (kx, k-slice)
constexpr int kW2Bytes = kWStep;                  // DEM/bias operand
constexpr int kA2Bytes = 128 * 16;                // DEM operand plane per stage
constexpr int kDemBytes = 640;                    // per-stage fp32 DEM halo row (128-byte aligned)
constexpr int kPipeThreads = 256;                 // producer, MMA, DEM prefetch, DEM builder, 4 epilogue warps
constexpr int kThreads = kPipes * kPipeThreads;
constexpr int kSmemBytes = kWBytes + kW2Bytes + kPipes * kStages * (kRowBytes + kA2Bytes + kDemBytes) + kA2Bytes + 2048;  // one CTA per SM

struct HeadParams {
  int H, W, N;          // HR tile extent and tiles in this launch
  long long total_rows; // N * (W / 128) * H strip rows
  float alpha;
  int half;             // 16-bit format: 0 bf16, 1 fp16
  float max_depth, denom;
  const __nv_bfloat16* wpack;   // feature weights [kx][kslice][2][96][8] followed by the DEM/bias operand [2][96][8]
  const float* dem;     // [N][H][W] normalised DEM
  float* pred_m;        // [N][H][W]
  float* pred_norm;     // [N][H][W] or nullptr
  float w2[kCmid];      // 1x1 projection
  float b2;
  long long* stats;     // optional [16] wait-cycle counters of CTA 0 (FSR_HEAD_STATS=1)
};

// mbarrier wait that adds the cycles spent to `cnt` (pipeline diagnostics; `cnt` is dead code when unused)
// mbarrier wait for roles that normally wait long (epilogue, producers): try_wait with a suspend-time hint, so the
// polling does not compete with the tensor core for shared-memory bandwidth.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
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
        : "r"(smem_u32(bar)), "r"(parity), "r"(2000u)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 22)) __trap();
  }
}

template <bool STATS>
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, long long& cnt) {
  if (STATS) {
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    cnt += clock64() - t0;
  } else {
    mbar_wait(bar, parity);
  }
}

// Walks the items of this CTA's row range; every warp role iterates the same sequence.
struct ItemIter {
  long long r, r_end;
  int H, segs;
  int img, xs, y0, rows;
  __device__ ItemIter(const HeadParams& p, int pipe) : H(p.H), segs(p.W / 128) {
    const long long vb = (long long)blockIdx.x * kPipes + pipe, nvb = (long long)gridDim.x * kPipes;
    r = p.total_rows * vb / nvb;
    r_end = p.total_rows * (vb + 1) / nvb;
  }
  __device__ bool next() {
    if (r >= r_end) return false;
    const long long strip = r / H;
    y0 = (int)(r - strip * H);
    const long long left = r_end - r;
    rows = (H - y0) < left ? (H - y0) : (int)left;
    img = (int)(strip / segs);
    xs = (int)(strip % segs);
    r += rows;
    return true;
  }
};

__device__ __forceinline__ uint16_t to16(float v, int half) {
  if (half) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float from16(uint16_t u, int half) {
  if (half) return __half2float(__ushort_as_half(u));
  return __bfloat162float(__ushort_as_bfloat16(u));
}

// Accumulating MMA with the 64-bit shared-memory descriptors given as 32-bit halves, so the issuing warp only does
// 32-bit adds per instruction (descriptor layout: tc_common.cuh, smem_desc_kmajor).
__device__ __forceinline__ void umma_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .b64 da, db;\n"
      ".reg .pred p;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.eq.u32 p, %4, %4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) { return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
// 32 lanes x 32 columns of zeros -> TMEM (re-arms an accumulator so that every MMA can accumulate)
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
      ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int ACT>
__device__ __forceinline__ float act_fn(float v, float alpha) {
  if (ACT == FSR_ACT_RELU) return fmaxf(v, 0.0f);
  if (ACT == FSR_ACT_LEAKY) return v > 0.0f ? v : v * alpha;
  return v;
}

template <int ACT, bool STATS>
__global__ void __launch_bounds__(kThreads, 1)
head_tc_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ HeadParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp_abs = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = warp_abs >> 3;     // pipeline of this warp
  const int warp = warp_abs & 7;   // role inside the pipeline
  uint8_t* smem_w = smem_raw;                                   // 18432 B (shared by both pipelines)
  uint8_t* smem_w2 = smem_w + kWBytes;                          // 3072 B
  uint8_t* smem_rows_all = smem_w2 + kW2Bytes;                  // kPipes x kStages x 8320 B (128-B aligned)
  uint8_t* smem_a2_all = smem_rows_all + kPipes * kStages * kRowBytes;   // kPipes x kStages x 2048 B
  uint8_t* smem_zero = smem_a2_all + kPipes * kStages * kA2Bytes;        // 2048 B of zeros: upper K plane of every DEM operand
  uint8_t* smem_dem_all = smem_zero + kA2Bytes;                 // kPipes x kStages x 640 B fp32 DEM halo rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_dem_all + kPipes * kStages * kDemBytes);
  constexpr int kBarsPerPipe = 3 * kStages + 2 * kSlots;
  uint64_t* w_full = bars;
  uint64_t* pbars = bars + 1 + v * kBarsPerPipe;
  uint64_t* row_full = pbars;                    // [kStages]  TMA bytes + DEM builder arrival -> MMA
  uint64_t* row_empty = row_full + kStages;      // [kStages]  MMA -> producers
  uint64_t* dem_full = row_empty + kStages;      // [kStages]  DEM prefetcher (32 cp.async arrivals) -> builder
  uint64_t* slot_full = dem_full + kStages;      // [kSlots]   MMA -> epilogue
  uint64_t* slot_empty = slot_full + kSlots;     // [kSlots]   epilogue (4 warps) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + kPipes * kBarsPerPipe);
  uint8_t* smem_rows = smem_rows_all + v * kStages * kRowBytes;
  uint8_t* smem_a2 = smem_a2_all + v * kStages * kA2Bytes;
  uint8_t* smem_dem = smem_dem_all + v * kStages * kDemBytes;

  if (warp_abs == 0 && lane == 0) tma_prefetch_desc(&tmF);
  if (warp_abs == 1) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int pp = 0; pp < kPipes; ++pp) {
        uint64_t* b = bars + 1 + pp * kBarsPerPipe;
        for (int i = 0; i < kStages; ++i) {
          mbar_init(&b[i], 2);                       // row_full
          mbar_init(&b[kStages + i], 1);             // row_empty
          mbar_init(&b[2 * kStages + i], 32);        // dem_full
        }
        for (int i = 0; i < kSlots; ++i) {
          mbar_init(&b[3 * kStages + i], 1);         // slot_full
          mbar_init(&b[3 * kStages + kSlots + i], 4);  // slot_empty
        }
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp_abs == 2) {
    for (int i = lane; i < kA2Bytes / 16; i += 32) reinterpret_cast<uint4*>(smem_zero)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_alloc_base = *tmem_slot;
  const uint32_t tmem_base = tmem_alloc_base + v * (kSlots * kCmid);  // this pipeline's half of the accumulator columns
  if (warp >= 4) {
    // every MMA accumulates: start from zeroed accumulators (the epilogue re-zeroes a slot after reading it)
    for (int sl = 0; sl < kSlots; ++sl) tmem_zero32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + sl * kCmid);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one halo row per input row ===================
    if (lane == 0) {
      if (v == 0) {
        mbar_expect_tx(w_full, kWBytes + kW2Bytes);
        bulk_load_1d(smem_w, p.wpack, kWBytes + kW2Bytes, w_full);
      }
      int g = 0, s = 0;  // running input-row counter of this CTA and its smem stage
      uint32_t ph = 1;
      for (ItemIter it(p, v); it.next();) {
        const int n_in = it.rows + 2;
        for (int i = 0; i < n_in; ++i, ++g) {
          mbar_wait_relaxed(&row_empty[s], ph);
          mbar_expect_tx(&row_full[s], kRowBytes);
          tma_load_5d(smem_rows + s * kRowBytes, &tmF, &row_full[s], 0, it.xs * 128 - 1, it.y0 - 1 + i, it.img, 0);
          if (++s == kStages) { s = 0; ph ^= 1; }
        }
      }
      if (STATS && p.stats && blockIdx.x == 0 && v == 0) p.stats[2] = g;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================================================================
    // One warp feeds the tensor core, so its instruction count per row is what limits the kernel once the epilogue
    // is light: the loop is warp-uniform, descriptors are 32-bit bases plus immediates, ring indices and barrier
    // phases are carried incrementally, and only the tcgen05 instructions themselves are issued by one lane.
    const uint32_t idesc0 = idesc_16(128, 0, p.half);  // + (N >> 3) << 17 with N = 32 * accumulators
    const uint32_t idesc96 = idesc0 + (3u << 19);
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1
    mbar_wait(w_full, 0);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem_w), kNfull * 16);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem_rows), kPlaneBytes);
    // DEM operand of stage s: plane 0 at a2 + s * 2 KB, plane 1 = the shared zero plane, i.e. LBO shrinks as s grows:
    // lo(s) = lo(0) + s * (128 - (128 << 16))
    const uint32_t a2_lo0 = desc_lo(smem_u32(smem_a2), smem_u32(smem_zero) - smem_u32(smem_a2));
    constexpr uint32_t kA2Step = (kA2Bytes >> 4) - ((kA2Bytes >> 4) << 16);
    const uint32_t bar_full0 = smem_u32(row_full), bar_empty0 = smem_u32(row_empty);
    const uint32_t bar_slot_full = smem_u32(slot_full), bar_slot_empty = smem_u32(slot_empty);
    const bool leader = elect_one();
    constexpr uint32_t kB = kWStep >> 4, kP = (2 * kPlaneBytes) >> 4;
    // per-stage values, carried incrementally
    int s = 0;
    uint32_t sph = 0, a = a_lo0, a2 = a2_lo0, bar_full = bar_full0, bar_empty = bar_empty0;
    int go = 0;           // output rows of earlier items
    long long c_se = 0, c_rf = 0, t0 = STATS ? clock64() : 0;
    for (ItemIter it(p, v); it.next(); go += it.rows) {
      const int n_in = it.rows + 2;
      for (int i = 0; i < n_in; ++i) {
        const int R = go + i;  // ring index of the output row this input row opens (ky = 0)
        if (i < it.rows)       // its slot must have been read (and re-zeroed) by the epilogue
          mbar_wait_t<STATS>(&slot_empty[R & (kSlots - 1)], ((R / kSlots) & 1) ^ 1, c_se);
        mbar_wait_t<STATS>(&row_full[s], sph, c_rf);
        tc_fence_after();
        const int sa = (R - 2) & (kSlots - 1);
        if (i >= 2 && i < it.rows && sa <= kSlots - 3) {
          // steady state: three adjacent accumulators, one N = 96 MMA per K step
          const uint32_t d = tmem_base + sa * kCmid;
          if (leader) {
            umma_acc(d, a, b_lo0, desc_hi, idesc96);
            umma_acc(d, a + kP, b_lo0 + kB, desc_hi, idesc96);
            umma_acc(d, a + 1, b_lo0 + 2 * kB, desc_hi, idesc96);
            umma_acc(d, a + kP + 1, b_lo0 + 3 * kB, desc_hi, idesc96);
            umma_acc(d, a + 2, b_lo0 + 4 * kB, desc_hi, idesc96);
            umma_acc(d, a + kP + 2, b_lo0 + 5 * kB, desc_hi, idesc96);
            umma_acc(d, a2, b_lo0 + 6 * kB, desc_hi, idesc96);
            umma_commit_addr(bar_empty);
            umma_commit_addr(bar_slot_full + sa * 8);
          }
        } else {
          // item borders (fewer than three target rows) and ring wrap-around (two MMAs per K step)
          const int jlo = i - 2 > 0 ? i - 2 : 0;
          const int jhi = i < it.rows - 1 ? i : it.rows - 1;
          const int n = jhi - jlo + 1;
          const int s1 = (go + jlo) & (kSlots - 1);
          const int n1 = n < kSlots - s1 ? n : kSlots - s1;
          const uint32_t d1 = tmem_base + s1 * kCmid;
          const uint32_t id1 = idesc0 + ((uint32_t)n1 << 19);
          const uint32_t b1 = b_lo0 + (uint32_t)(jlo - (i - 2)) * kCmid;  // weight rows (16 B each) of ky = i - jlo first
          if (leader) {
            umma_acc(d1, a, b1, desc_hi, id1);
            umma_acc(d1, a + kP, b1 + kB, desc_hi, id1);
            umma_acc(d1, a + 1, b1 + 2 * kB, desc_hi, id1);
            umma_acc(d1, a + kP + 1, b1 + 3 * kB, desc_hi, id1);
            umma_acc(d1, a + 2, b1 + 4 * kB, desc_hi, id1);
            umma_acc(d1, a + kP + 2, b1 + 5 * kB, desc_hi, id1);
            umma_acc(d1, a2, b1 + 6 * kB, desc_hi, id1);
            if (n1 < n) {
              const uint32_t id2 = idesc0 + ((uint32_t)(n - n1) << 19);
              const uint32_t b2 = b1 + n1 * kCmid;
              umma_acc(tmem_base, a, b2, desc_hi, id2);
              umma_acc(tmem_base, a + kP, b2 + kB, desc_hi, id2);
              umma_acc(tmem_base, a + 1, b2 + 2 * kB, desc_hi, id2);
              umma_acc(tmem_base, a + kP + 1, b2 + 3 * kB, desc_hi, id2);
              umma_acc(tmem_base, a + 2, b2 + 4 * kB, desc_hi, id2);
              umma_acc(tmem_base, a + kP + 2, b2 + 5 * kB, desc_hi, id2);
              umma_acc(tmem_base, a2, b2 + 6 * kB, desc_hi, id2);
            }
            umma_commit_addr(bar_empty);
            if (i >= 2) umma_commit_addr(bar_slot_full + sa * 8);
          }
        }
        __syncwarp();
        a += kRowBytes >> 4; a2 += kA2Step; bar_full += 8; bar_empty += 8;
        if (++s == kStages) { s = 0; sph ^= 1; a = a_lo0; a2 = a2_lo0; bar_full = bar_full0; bar_empty = bar_empty0; }
      }
    }
    (void)bar_full; (void)bar_slot_empty;
    if (STATS && p.stats && blockIdx.x == 0 && v == 0 && lane == 0) { p.stats[3] = c_se; p.stats[4] = c_rf; p.stats[5] = 0; p.stats[6] = clock64() - t0; }
  } else if (warp == 2) {
    // ===================== DEM prefetcher: fp32 halo rows -> smem, up to kStages rows ahead =================
    // 4-byte cp.async (zero-filled outside the tile); completion is signalled straight to the builder's mbarrier,
    // so this warp never waits for memory.
    int s = 0;
    uint32_t ph = 1;
    for (ItemIter it(p, v); it.next();) {
      const int n_in = it.rows + 2;
      for (int i = 0; i < n_in; ++i) {
        const int y = it.y0 - 1 + i;
        const bool yok = y >= 0 && y < p.H;
        const float* row = p.dem + ((size_t)it.img * p.H + (yok ? y : 0)) * p.W;
        const uint32_t dst = smem_u32(smem_dem + s * kDemBytes);
        mbar_wait_relaxed(&row_empty[s], ph);
        for (int k = lane; k < kRowPx; k += 32) {
          const int x = it.xs * 128 - 1 + k;
          const bool ok = yok && x >= 0 && x < p.W;
          const float* src = row + (ok ? x : 0);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + k * 4), "l"(src), "r"(ok ? 4 : 0) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&dem_full[s])) : "memory");
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===================== DEM operand builder: [128 px][hi(-1,0,+1), lo(-1,0,+1), 1, 1] ====================
    const uint32_t ones = (uint32_t)to16(1.0f, p.half) * 0x10001u;
    int s = 0;
    uint32_t ph = 0;
    for (ItemIter it(p, v); it.next();) {
      const int n_in = it.rows + 2;
      for (int i = 0; i < n_in; ++i) {
        mbar_wait_relaxed(&dem_full[s], ph);
        const float* drow = reinterpret_cast<const float*>(smem_dem + s * kDemBytes) + lane * 4;
        uint16_t hi[6], lo[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const float d = drow[k];
          hi[k] = to16(d, p.half);
          lo[k] = to16(d - from16(hi[k], p.half), p.half);
        }
        uint4* dst = reinterpret_cast<uint4*>(smem_a2 + s * kA2Bytes) + lane * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = (uint32_t)hi[q] | ((uint32_t)hi[q + 1] << 16);
          v.y = (uint32_t)hi[q + 2] | ((uint32_t)lo[q] << 16);
          v.z = (uint32_t)lo[q + 1] | ((uint32_t)lo[q + 2] << 16);
          v.w = ones;
          dst[q] = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&row_full[s]);
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: each finished accumulator is read exactly once ========================
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float w2r[kCmid];
#pragma unroll
    for (int c = 0; c < kCmid; ++c) w2r[c] = p.w2[c];
    int go = 0;
    long long c_sf = 0, t0 = STATS ? clock64() : 0;
    for (ItemIter it(p, v); it.next(); go += it.rows) {
      const int x = it.xs * 128 + m;
      for (int j = 0; j < it.rows; ++j) {
        const int r = go + j;
        const int slot = r & (kSlots - 1);
        if (STATS) mbar_wait_t<STATS>(&slot_full[slot], (r / kSlots) & 1, c_sf); else mbar_wait_relaxed(&slot_full[slot], (r / kSlots) & 1);
        tc_fence_after();
        float v[kCmid];
        tmem_ld32(lane_addr + slot * kCmid, v);
        tmem_ld_wait();
        tmem_zero32(lane_addr + slot * kCmid);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[slot]);
        float o0 = p.b2, o1 = 0.0f, o2 = 0.0f, o3 = 0.0f;
#pragma unroll
        for (int c = 0; c < kCmid; c += 4) {
          o0 = fmaf(act_fn<ACT>(v[c], p.alpha), w2r[c], o0);
          o1 = fmaf(act_fn<ACT>(v[c + 1], p.alpha), w2r[c + 1], o1);
          o2 = fmaf(act_fn<ACT>(v[c + 2], p.alpha), w2r[c + 2], o2);
          o3 = fmaf(act_fn<ACT>(v[c + 3], p.alpha), w2r[c + 3], o3);
        }
        const float out = (o0 + o1) + (o2 + o3);
        const size_t off = ((size_t)it.img * p.H + (it.y0 + j)) * p.W + x;
        if (p.pred_norm) p.pred_norm[off] = out;
        const float yn = fminf(fmaxf(out, 0.0f), 1.0f);
        p.pred_m[off] = fminf(fmaxf(expm1f(__fmul_rn(yn, p.denom)), 0.0f), p.max_depth);
      }
    }
    if (p.stats && blockIdx.x == 0 && warp == 4 && lane == 0) { p.stats[7] = c_sf; p.stats[8] = clock64() - t0; }
  }
  __syncthreads();
  if (warp_abs == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_alloc_base, 512);
  }
}

}  // namespace

// Packed operands of the head (host side, built once per engine); column n = b*32 + co with ky = 2 - b:
//   [kx][k-slice j][plane pl][n][8]  feature weights W[ky][kx][ci = j*16 + pl*8 + e][co]
//   [plane][n][8]                    DEM/bias operand: k 0-2 = Wdem[ky][kx], k 3-5 = the same (they multiply the lo
//                                    parts), k 6/7 = (hi, lo) of bias[co] for ky == 1, everything else 0
size_t head2_pack_elems() { return (size_t)(kWBytes + kW2Bytes) / 2; }

void head2_pack(const float* w /* [3][3][33][32] */, const float* bias, uint16_t* dst, uint16_t (*cvt)(float), float (*back)(uint16_t)) {
  const int cin_real = 33;
  size_t pos = 0;
  for (int kx = 0; kx < 3; ++kx)
    for (int j = 0; j < 2; ++j)
      for (int pl = 0; pl < 2; ++pl)
        for (int n = 0; n < kNfull; ++n)
          for (int e = 0; e < 8; ++e, ++pos) {
            const int ky = 2 - n / kCmid, co = n % kCmid, ci = j * 16 + pl * 8 + e;
            dst[pos] = cvt(w[(((size_t)ky * 3 + kx) * cin_real + ci) * kCmid + co]);
          }
  for (int pl = 0; pl < 2; ++pl)
    for (int n = 0; n < kNfull; ++n)
      for (int e = 0; e < 8; ++e, ++pos) {
        const int ky = 2 - n / kCmid, co = n % kCmid;
        uint16_t v = cvt(0.0f);
        if (pl == 0 && e < 6) v = cvt(w[(((size_t)ky * 3 + (e % 3)) * cin_real + 32) * kCmid + co]);
        if (pl == 0 && ky == 1 && bias) {
          const uint16_t bh = cvt(bias[co]);
          if (e == 6) v = bh;
          if (e == 7) v = cvt(bias[co] - back(bh));
        }
        dst[pos] = v;
      }
}

void launch_head2_tc(const __nv_bfloat16* feat, long long plane, const __nv_bfloat16* wpack, const float* w2, const float* b2,
                     const float* dem, float* pred_m, float* pred_norm, int n_img, int H, int W, int cin, int cmid, int ksz,
                     int act, float alpha, float max_depth, float denom, int half, int n_sms, cudaStream_t s) {
  FSR_REQUIRE(cin == 32 && cmid == kCmid && ksz == 3, "head tensor-core path is specialised for 32 -> 32 channels, 3x3");
  FSR_REQUIRE(W % 128 == 0 && H >= 1, "head tensor-core path needs W % 128 == 0");
  HeadParams p{};
  p.H = H; p.W = W; p.N = n_img;
  p.total_rows = (long long)n_img * (W / 128) * H;
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
  CUtensorMap mF = make_cp8_tensor_map(feat, W, H, n_img, cin / 8, plane, kRowPx, 1, 1, cin / 8);
  static long long* d_stats = nullptr;
  static int stat_calls = 0;
  if (getenv("FSR_HEAD_STATS")) {
    if (!d_stats) FSR_CUDA(cudaMalloc(&d_stats, 16 * sizeof(long long)));
    p.stats = d_stats;
  }
  const int grid = p.total_rows < n_sms ? (int)p.total_rows : n_sms;
  auto go = [&](auto kernel) {
    FSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    kernel<<<grid, kThreads, kSmemBytes, s>>>(mF, p);
  };
  if (p.stats) {
    if (act == FSR_ACT_RELU) go(head_tc_kernel<FSR_ACT_RELU, true>);
    else if (act == FSR_ACT_LEAKY) go(head_tc_kernel<FSR_ACT_LEAKY, true>);
    else go(head_tc_kernel<FSR_ACT_NONE, true>);
  } else {
    if (act == FSR_ACT_RELU) go(head_tc_kernel<FSR_ACT_RELU, false>);
    else if (act == FSR_ACT_LEAKY) go(head_tc_kernel<FSR_ACT_LEAKY, false>);
    else go(head_tc_kernel<FSR_ACT_NONE, false>);
  }
  FSR_LAUNCH_CHECK();
  if (p.stats && ++stat_calls == 8) {
    long long h[16];
    FSR_CUDA(cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "[head stats, CTA 0, %lld input rows] mma: slot_empty %lld row_full %lld of %lld | epilogue grp0: slot_full %lld of %lld cycles\n",
            h[2], h[3], h[4], h[6], h[7], h[8]);
  }
}

}  // namespace fsr
