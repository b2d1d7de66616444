// tcgen05 kernels for the high-resolution end of the network (FSR_PREC_BF16): the 16x transposed
// convolution and the fused head (conv3x3 over concat(features, dem_hr) -> act -> conv1x1 -> log1p inversion).
// Together they are ~90 % of the FLOPs behind `session.run` (floodsr/engine/ort.py:193) and the
// invert_depth_log1p_np call that follows it (ort.py:196, preprocessing.py:154-164).
//
// convT (kernel == stride == k): GEMM  M = LR pixels, K = cin, N = k*k*cout, N-tile 256 = 256/cout
//   adjacent kx positions of one ky.  A is loaded once per CTA, weights stream through a 2-stage ring,
//   accumulators are double-buffered in TMEM (2 x 256 columns); the epilogue writes the HR feature map
//   in CP8 layout, 128 contiguous bytes per thread.
//
// head: each CTA owns a 128-pixel-wide strip of RB output rows.  For every *input* row r of the strip one
//   accumulator slot D'_r[x, (ky, co)] (N = 3*cmid columns) is produced from that row alone:
//       D'_r[x, (ky, co)] = sum_{kx, ci} F[r, x + kx - 1, ci] * W[ky, kx, ci, co]
//   the kx taps are start-address offsets into the row's halo buffer (no-swizzle CP8 operand, 16 B per pixel
//   per plane).  Output row y then only needs same-lane TMEM reads:
//       conv[y, x, co] = D'_{y-1}[x, (0, co)] + D'_y[x, (1, co)] + D'_{y+1}[x, (2, co)]
//   so every feature row is read from L2 once and is used by 3*cmid output columns per MMA (N = 96 instead
//   of 32: shared-memory operand traffic per FLOP drops ~2x).  The 1-channel DEM term, bias, activation, the
//   1x1 projection and the log1p inversion run in the epilogue on CUDA cores in fp32.
#include "fsr_engine.cuh"
#include "tc_common.cuh"

namespace fsr {

using namespace tc;

CUtensorMap make_cp8_tensor_map(const void* base, int W, int H, int N, int chunks, long long plane, int bw, int bh, int bn, int kc);
void conv_tc_tile_box(int H, int W, int& bw, int& bh, int& bn);

namespace {

// =============================================================================================================
// transposed convolution
// =============================================================================================================

constexpr int kCtThreads = 192;
constexpr int kCtBN = 256;

struct ConvTParams {
  int Hin, Win, N;       // LR extent, images in this launch
  int bw, bh, bn;        // LR pixel box (128 pixels)
  int tiles_x, tiles_y;
  int k;                 // kernel == stride
  int cin, cout;         // cin % 16 == 0, 256 % cout == 0
  int n_tiles_total;     // k*k*cout / 256
  int n_tiles_per_cta;
  int act;
  float alpha;
  long long plane_out;   // HR pixels per CP8 plane of the output
  const __nv_bfloat16* wpack;  // [n_tile][cin/8][256][8]
  const float* bias;           // [cout] or nullptr
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(kCtThreads, 1)
convt_tc_kernel(const __grid_constant__ CUtensorMap tmA, const ConvTParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int kc = p.cin / 8;
  const uint32_t a_bytes = (uint32_t)kc * 128u * 16u;
  const uint32_t b_bytes = (uint32_t)kc * kCtBN * 16u;
  uint8_t* smem_a = smem_raw;                       // up to 8 planes x 2 KB
  uint8_t* smem_b = smem_raw + 8 * 2048;            // 2 stages x up to 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + 2 * 8 * kCtBN * 16);
  uint64_t* a_full = bars;
  uint64_t* b_full = bars + 1;     // [2]
  uint64_t* b_empty = bars + 3;    // [2]
  uint64_t* acc_full = bars + 5;   // [2]
  uint64_t* acc_empty = bars + 7;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int t = blockIdx.x;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int tn = t / p.tiles_y;
  const int nt0 = blockIdx.y * p.n_tiles_per_cta;
  const int nt1 = min(nt0 + p.n_tiles_per_cta, p.n_tiles_total);
  const int n_my = nt1 - nt0;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmA);
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(a_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&b_full[i], 1);
        mbar_init(&b_empty[i], 1);
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(a_full, a_bytes);
      tma_load_5d(smem_a, &tmA, a_full, 0, tx * p.bw, ty * p.bh, tn * p.bn, 0);
      for (int i = 0; i < n_my; ++i) {
        const int s = i & 1;
        mbar_wait(&b_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&b_full[s], b_bytes);
        bulk_load_1d(smem_b + (size_t)s * (8 * kCtBN * 16), p.wpack + (size_t)(nt0 + i) * ((size_t)kc * kCtBN * 8), b_bytes, &b_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(128, kCtBN);
      mbar_wait(a_full, 0);
      const uint32_t a_addr = smem_u32(smem_a);
      for (int i = 0; i < n_my; ++i) {
        const int s = i & 1;
        mbar_wait(&acc_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_wait(&b_full[s], (i >> 1) & 1);
        tc_fence_after();
        const uint32_t b_addr = smem_u32(smem_b + (size_t)s * (8 * kCtBN * 16));
        for (int j = 0; j < kc / 2; ++j) {
          const uint64_t da = smem_desc_kmajor(a_addr + j * 2 * 2048, 2048, 128);
          const uint64_t db = smem_desc_kmajor(b_addr + j * 2 * (kCtBN * 16), kCtBN * 16, 128);
          umma_bf16(tmem_base + s * kCtBN, da, db, idesc, j > 0 ? 1u : 0u);
        }
        umma_commit(&b_empty[s]);
        umma_commit(&acc_full[s]);
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int lx = m % p.bw;
    const int ly = (m / p.bw) % p.bh;
    const int ln = m / (p.bw * p.bh);
    const int n_img = tn * p.bn + ln;
    const bool valid = n_img < p.N;
    const int Hout = p.Hin * p.k, Wout = p.Win * p.k;
    const int yin = ty * p.bh + ly, xin = tx * p.bw + lx;
    const int kx_per_tile = kCtBN / p.cout;
    const int tiles_per_ky = p.k / kx_per_tile;
    const int cchunks = p.cout / 8;
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1;
      const int nt = nt0 + i;
      const int ky = nt / tiles_per_ky;
      const int kx0 = (nt % tiles_per_ky) * kx_per_tile;
      mbar_wait(&acc_full[s], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + s * kCtBN;
      const long long row_pix = ((long long)n_img * Hout + (yin * p.k + ky)) * Wout + (long long)xin * p.k + kx0;
#pragma unroll 1
      for (int c8 = 0; c8 < cchunks; ++c8) {
        float b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (p.bias) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c8 * 8));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c8 * 8 + 4));
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
        }
        __nv_bfloat16* dst = p.out + ((long long)c8 * p.plane_out + row_pix) * 8;
#pragma unroll 1
        for (int kx = 0; kx < kx_per_tile; ++kx) {
          float v[8];
          tmem_ld8(taddr + kx * p.cout + c8 * 8, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = apply_act(v[e] + b[e], p.act, p.alpha);
            uint4 o;
            o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(dst + (long long)kx * 8) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[s]);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =============================================================================================================
// fused head
// =============================================================================================================

constexpr int kHeadCmid = 32;                     // mid channels (N = 3 * 32 = 96 accumulator columns per slot)
constexpr int kHeadN = 3 * kHeadCmid;
constexpr int kHeadSlots = 5;
constexpr int kHeadRowStages = 6;
constexpr int kHeadRowPx = 130;                   // 128 + left/right halo pixel
constexpr int kHeadPlaneBytes = kHeadRowPx * 16;  // 2080
constexpr int kHeadRowBytes = 4 * kHeadPlaneBytes;  // 32 channels = 4 planes
constexpr int kHeadWBytes = 3 * 2 * (2 * kHeadN * 16);  // (kx, k-slice) x [2 planes][96][8] bf16 = 18432
constexpr int kHeadThreads = 64 + 256;            // producer warp, MMA warp, 8 epilogue warps

struct HeadConsts {
  float wdem[9][kHeadCmid];  // DEM-channel taps of the 3x3 conv [ky*3+kx][co]
  float bias[kHeadCmid];
  float w2[kHeadCmid];       // 1x1 projection
  float b2;
};

struct HeadParams {
  int H, W, N;           // HR tile extent (512 x 512) and images in this launch
  int rb;                // output rows per CTA
  int act;
  float alpha;
  float max_depth, denom;
  const __nv_bfloat16* wpack;  // [kx][kslice][2][96][8]
  const float* dem;      // [N][H][W] normalised DEM
  float* pred_m;         // [N][H][W]
  float* pred_norm;      // [N][H][W] or nullptr
  HeadConsts c;
};

__global__ void __launch_bounds__(kHeadThreads, 1)
head_tc_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ HeadParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem_w = smem_raw;                                  // 18 KB packed weights (resident)
  uint8_t* smem_rows = smem_raw + kHeadWBytes;                 // kHeadRowStages x 8320 B (each 128-B aligned: 8320 = 65*128)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_rows + kHeadRowStages * kHeadRowBytes);
  uint64_t* w_full = bars;
  uint64_t* row_full = bars + 1;                                   // [kHeadRowStages]
  uint64_t* row_empty = row_full + kHeadRowStages;                 // [kHeadRowStages]
  uint64_t* slot_full = row_empty + kHeadRowStages;                // [kHeadSlots]
  uint64_t* slot_empty = slot_full + kHeadSlots;                   // [kHeadSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(slot_empty + kHeadSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA -> (image, x segment, row block)
  const int segs = p.W / 128;
  const int rblocks = p.H / p.rb;
  int t = blockIdx.x;
  const int xs = t % segs;
  t /= segs;
  const int rbk = t % rblocks;
  const int img = t / rblocks;
  const int x0 = xs * 128;
  const int ya = rbk * p.rb;
  const int n_in = p.rb + 2;   // input rows ya-1 .. ya+rb
  const int n_out = p.rb;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmF);
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int i = 0; i < kHeadRowStages; ++i) {
        mbar_init(&row_full[i], 1);
        mbar_init(&row_empty[i], 1);
      }
      for (int i = 0; i < kHeadSlots; ++i) {
        mbar_init(&slot_full[i], 1);
        // an accumulator slot is read by the three output rows around its input row (4 warps each)
        mbar_init(&slot_empty[i], 12);
      }
      fence_barrier_init();
      // the strip's first two input rows have only one / two reading output rows: pre-arrive for the rest
      mbar_arrive_n(&slot_empty[0], 8);
      mbar_arrive_n(&slot_empty[1], 4);
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: weights once, then one halo row per input row =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, kHeadWBytes);
      bulk_load_1d(smem_w, p.wpack, kHeadWBytes, w_full);
      for (int i = 0; i < n_in; ++i) {
        const int s = i % kHeadRowStages;
        mbar_wait(&row_empty[s], ((i / kHeadRowStages) & 1) ^ 1);
        mbar_expect_tx(&row_full[s], kHeadRowBytes);
        // rows -1 and H, and pixels -1 and W, are outside the tensor: TMA zero-fills them ('same' padding)
        tma_load_5d(smem_rows + s * kHeadRowBytes, &tmF, &row_full[s], 0, x0 - 1, ya - 1 + i, img, 0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 6 MMAs (3 kx taps x 2 K slices), N = 96, per input row ==========
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(128, kHeadN);
      mbar_wait(w_full, 0);
      const uint32_t w_addr = smem_u32(smem_w);
      for (int i = 0; i < n_in; ++i) {
        const int s = i % kHeadRowStages;
        const int slot = i % kHeadSlots;
        mbar_wait(&slot_empty[slot], ((i / kHeadSlots) & 1) ^ 1);
        mbar_wait(&row_full[s], (i / kHeadRowStages) & 1);
        tc_fence_after();
        const uint32_t r_addr = smem_u32(smem_rows + s * kHeadRowBytes);
        const uint32_t d_addr = tmem_base + slot * kHeadN;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t da = smem_desc_kmajor(r_addr + j * 2 * kHeadPlaneBytes + kx * 16, kHeadPlaneBytes, 128);
            const uint64_t db = smem_desc_kmajor(w_addr + (kx * 2 + j) * (2 * kHeadN * 16), kHeadN * 16, 128);
            umma_bf16(d_addr, da, db, idesc, (kx | j) ? 1u : 0u);
          }
        }
        umma_commit(&row_empty[s]);
        umma_commit(&slot_full[slot]);
      }
    }
  } else {
    // ===================== epilogue: two warpgroups alternate output rows =====================
    const int ew = warp - 2;        // 0..7
    const int wg = ew >> 2;         // warpgroup 0/1
    const int q = warp & 3;         // TMEM lane quarter
    const int m = q * 32 + lane;
    const int x = x0 + m;
    const float* dem_img = p.dem + (size_t)img * p.H * p.W;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int o = wg; o < n_out; o += 2) {
      const int y = ya + o;
      // DEM-channel term + bias while the MMAs for this row are still in flight
      float acc[kHeadCmid];
#pragma unroll
      for (int c = 0; c < kHeadCmid; ++c) acc[c] = p.c.bias[c];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = x + kx - 1;
          const float d = (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) ? __ldg(dem_img + (size_t)yy * p.W + xx) : 0.0f;
#pragma unroll
          for (int c = 0; c < kHeadCmid; ++c) acc[c] = fmaf(d, p.c.wdem[ky * 3 + kx][c], acc[c]);
        }
      }
      // same-lane sums over the three input rows' accumulator slots
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int i = o + ky;  // input-row index
        const int slot = i % kHeadSlots;
        mbar_wait(&slot_full[slot], (i / kHeadSlots) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(lane_addr + slot * kHeadN + ky * kHeadCmid, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < kHeadCmid; ++c) acc[c] += v[c];
      }
      // this row is done with its three slots; a slot returns to the MMA warp once all its readers are
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) mbar_arrive(&slot_empty[(o + ky) % kHeadSlots]);
      }
      float out = p.c.b2;
#pragma unroll
      for (int c = 0; c < kHeadCmid; ++c) out = fmaf(apply_act(acc[c], p.act, p.alpha), p.c.w2[c], out);
      const size_t off = ((size_t)img * p.H + y) * p.W + x;
      if (p.pred_norm) p.pred_norm[off] = out;
      const float yn = fminf(fmaxf(out, 0.0f), 1.0f);
      p.pred_m[off] = fminf(fmaxf(expm1f(__fmul_rn(yn, p.denom)), 0.0f), p.max_depth);
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// ---- host-side launchers ---------------------------------------------------------------------------------------

size_t convt_smem_bytes() { return 8 * 2048 + 2 * 8 * kCtBN * 16 + 10 * sizeof(uint64_t) + 16; }
size_t head_smem_bytes() {
  return kHeadWBytes + kHeadRowStages * kHeadRowBytes + (1 + 2 * kHeadRowStages + 2 * kHeadSlots) * sizeof(uint64_t) + 16;
}

void launch_convt_tc(const __nv_bfloat16* src, long long plane_in, const __nv_bfloat16* wpack, const float* bias,
                     __nv_bfloat16* dst, long long plane_out, int n_img, int Hin, int Win, int cin, int cout, int k, int act,
                     float alpha, cudaStream_t s) {
  FSR_REQUIRE(cin % 16 == 0 && cin <= 64, "convT tensor-core path needs cin in {16, 32, 48, 64}");
  FSR_REQUIRE(cout % 8 == 0 && kCtBN % cout == 0 && k % (kCtBN / cout) == 0, "convT tensor-core path: unsupported cout / kernel size");
  ConvTParams p{};
  p.Hin = Hin; p.Win = Win; p.N = n_img;
  conv_tc_tile_box(Hin, Win, p.bw, p.bh, p.bn);
  p.tiles_x = Win / p.bw;
  p.tiles_y = Hin / p.bh;
  p.k = k;
  p.cin = cin;
  p.cout = cout;
  p.n_tiles_total = k * k * cout / kCtBN;
  p.act = act;
  p.alpha = alpha;
  p.plane_out = plane_out;
  p.wpack = wpack;
  p.bias = bias;
  p.out = dst;
  const int m_tiles = p.tiles_x * p.tiles_y * ceil_div(n_img, p.bn);
  // enough CTAs to fill the 148 SMs about twice; each CTA keeps its A tile and loops over its N tiles
  int groups = ceil_div(2 * 148, m_tiles);
  groups = groups < 1 ? 1 : (groups > p.n_tiles_total ? p.n_tiles_total : groups);
  p.n_tiles_per_cta = ceil_div(p.n_tiles_total, groups);
  groups = ceil_div(p.n_tiles_total, p.n_tiles_per_cta);
  CUtensorMap mA = make_cp8_tensor_map(src, Win, Hin, n_img, cin / 8, plane_in, p.bw, p.bh, p.bn, cin / 8);
  static bool attr = false;
  if (!attr) { FSR_CUDA(cudaFuncSetAttribute(convt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)convt_smem_bytes())); attr = true; }
  dim3 grid((unsigned)m_tiles, (unsigned)groups);
  convt_tc_kernel<<<grid, kCtThreads, convt_smem_bytes(), s>>>(mA, p);
  FSR_LAUNCH_CHECK();
}

// feat: CP8 [4][n][H][W][8]; wpack: packed 3x3 weights of the 32 feature channels; hc: fp32 epilogue constants
void launch_head_tc(const __nv_bfloat16* feat, long long plane, const __nv_bfloat16* wpack, const float* wdem, const float* bias,
                    const float* w2, const float* b2, const float* dem, float* pred_m, float* pred_norm, int n_img, int H, int W,
                    int cin, int cmid, int ksz, int act, float alpha, float max_depth, float denom, cudaStream_t s) {
  FSR_REQUIRE(cin == 32 && cmid == kHeadCmid && ksz == 3, "head tensor-core path is specialised for 32 -> 32 channels, 3x3");
  FSR_REQUIRE(W % 128 == 0 && H % 32 == 0, "head tensor-core path needs W % 128 == 0 and H % 32 == 0");
  HeadParams p{};
  p.H = H; p.W = W; p.N = n_img;
  p.rb = 32;
  p.act = act;
  p.alpha = alpha;
  p.max_depth = max_depth;
  p.denom = denom;
  p.wpack = wpack;
  p.dem = dem;
  p.pred_m = pred_m;
  p.pred_norm = pred_norm;
  for (int t = 0; t < 9; ++t)
    for (int c = 0; c < kHeadCmid; ++c) p.c.wdem[t][c] = wdem[t * kHeadCmid + c];
  for (int c = 0; c < kHeadCmid; ++c) {
    p.c.bias[c] = bias ? bias[c] : 0.0f;
    p.c.w2[c] = w2[c];
  }
  p.c.b2 = b2 ? b2[0] : 0.0f;
  CUtensorMap mF = make_cp8_tensor_map(feat, W, H, n_img, cin / 8, plane, kHeadRowPx, 1, 1, cin / 8);
  static bool attr = false;
  if (!attr) { FSR_CUDA(cudaFuncSetAttribute(head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)head_smem_bytes())); attr = true; }
  const int ctas = n_img * (W / 128) * (H / p.rb);
  head_tc_kernel<<<ctas, kHeadThreads, head_smem_bytes(), s>>>(mF, p);
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
