// tcgen05 kernels for the high-resolution end of the network (FSR_PREC_FP16, graphs the fused kernel does not cover): the 16x transposed
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
  int half;              // 16-bit format: 0 bf16, 1 fp16
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
    // MMA issuer: warp-uniform loop, one elected lane issues
    const uint32_t idesc = idesc_16(128, kCtBN, p.half);
    mbar_wait(a_full, 0);
    const uint64_t da_base = smem_desc_kmajor(smem_u32(smem_a), 2048, 128);
    const uint64_t db_base = smem_desc_kmajor(smem_u32(smem_b), kCtBN * 16, 128);
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1;
      mbar_wait(&acc_empty[s], ((i >> 1) & 1) ^ 1);
      mbar_wait(&b_full[s], (i >> 1) & 1);
      tc_fence_after();
      const uint64_t db = db_base + (uint64_t)((s * (8 * kCtBN * 16)) >> 4);
      if (elect_one()) {
        for (int j = 0; j < kc / 2; ++j)
          umma_bf16(tmem_base + s * kCtBN, da_base + (uint64_t)((j * 2 * 2048) >> 4), db + (uint64_t)((j * 2 * (kCtBN * 16)) >> 4), idesc,
                    j > 0 ? 1u : 0u);
        umma_commit(&b_empty[s]);
        umma_commit(&acc_full[s]);
      }
      __syncwarp();
    }
  } else {
    // epilogue: thread = one LR pixel; per (ky, 8-kx group) it owns 128 contiguous bytes in each 8-channel plane of
    // the HR row, written as 256-bit stores (two adjacent kx per store: every 32-byte sector is written whole).
    // Specialised for cout == 32 (one kx = 32 accumulator columns).
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
    float bias[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bias[c] = p.bias ? __ldg(p.bias + c) : 0.0f;
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
      for (int kp = 0; kp < kCtBN / 64; ++kp) {
        float v0[32], v1[32];
        tmem_ld32(taddr + kp * 64, v0);
        tmem_ld32(taddr + kp * 64 + 32, v1);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            float a[8], b[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              a[e] = apply_act(v0[c8 * 8 + e] + bias[c8 * 8 + e], p.act, p.alpha);
              b[e] = apply_act(v1[c8 * 8 + e] + bias[c8 * 8 + e], p.act, p.alpha);
            }
            const uint4 lo = pack_x8(a, p.half), hi = pack_x8(b, p.half);
            __nv_bfloat16* dst = p.out + ((long long)c8 * p.plane_out + row_pix + kp * 2) * 8;
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(lo.x), "r"(lo.y), "r"(lo.z),
                         "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                         : "memory");
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

}  // namespace

// ---- host-side launchers ---------------------------------------------------------------------------------------

size_t convt_smem_bytes() { return 8 * 2048 + 2 * 8 * kCtBN * 16 + 10 * sizeof(uint64_t) + 16; }

void launch_convt_tc(const __nv_bfloat16* src, long long plane_in, const __nv_bfloat16* wpack, const float* bias,
                     __nv_bfloat16* dst, long long plane_out, int n_img, int Hin, int Win, int cin, int cout, int k, int act,
                     float alpha, int half, cudaStream_t s) {
  FSR_REQUIRE(cin % 16 == 0 && cin <= 64, "convT tensor-core path needs cin in {16, 32, 48, 64}");
  FSR_REQUIRE(cout == 32 && k % (kCtBN / cout) == 0, "convT tensor-core path is specialised for 32 output channels and kernel sizes that are multiples of 8");
  ConvTParams p{};
  p.half = half;
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
  int groups = ceil_div(2 * current_sm_count(), m_tiles);
  groups = groups < 1 ? 1 : (groups > p.n_tiles_total ? p.n_tiles_total : groups);
  p.n_tiles_per_cta = ceil_div(p.n_tiles_total, groups);
  groups = ceil_div(p.n_tiles_total, p.n_tiles_per_cta);
  CUtensorMap mA = make_cp8_tensor_map(src, Win, Hin, n_img, cin / 8, plane_in, p.bw, p.bh, p.bn, cin / 8);
  static bool attr[64] = {false};
  if (first_on_device(attr)) FSR_CUDA(cudaFuncSetAttribute(convt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)convt_smem_bytes()));
  dim3 grid((unsigned)m_tiles, (unsigned)groups);
  convt_tc_kernel<<<grid, kCtThreads, convt_smem_bytes(), s>>>(mA, p);
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
