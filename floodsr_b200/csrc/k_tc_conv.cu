// tcgen05 implicit-GEMM convolution for the low-resolution backbone (FSR_PREC_FP16, and FSR_PREC_FP32 with split operands) and the CP8 helper kernels.
//
// Replaces ONNX Runtime's Conv kernels behind `session.run` (floodsr/engine/ort.py:193) for the ResUNet
// encoder/decoder layers.  GEMM view: M = 128 output pixels (a bw x bh x bn box of the batched NHW grid),
// N = BN output channels, K = taps x input channels over concat(src0, src1).
//
//   warp 0   TMA producer: per (tap, K-stage) one 5-D tiled TMA load of the shifted activation box
//            (out-of-bounds coordinates are zero-filled = 'same' padding) + one bulk copy of packed weights
//   warp 1   MMA issuer: tcgen05.mma kind::f16, accumulators in TMEM (BN fp32 columns)
//   warps 2-5 epilogue: tcgen05.ld -> +bias (+residual) -> activation -> bf16 -> 16-byte CP8 stores
//
// Activations live in "CP8" layout [C/8][N][H][W][8] bf16 (see tc_common.cuh), so both UMMA operands use the
// no-swizzle K-major canonical layout and a tap is just a TMA coordinate offset.
#include "fsr_engine.cuh"
#include "tc_common.cuh"

namespace fsr {

using namespace tc;

namespace {

constexpr int kMaxStages = 8;
constexpr int kConvPairs = 2;                           // M tiles one CTA of conv_tc_kernel may compute (sharing the weight slices)
template <int PAIR>
constexpr int kConvThreads = 32 + 32 * PAIR + 128;  // producer warp, one MMA warp per M tile, 4 epilogue warps

// Programmatic dependent launch: consecutive convolution layers are launched with the stream-serialisation attribute, so a
// layer's CTAs may start (barrier set-up, TMEM allocation, constant weight loads) while the previous layer drains;
// griddep_wait() returns once the previous kernel has completed and its writes are visible.  Nothing before the wait may
// read activations or write global memory.  Both instructions are no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, int threads, size_t smem, cudaStream_t s, Args... args) {
  static const bool off = getenv("FSR_NO_PDL") != nullptr;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = off ? 0 : 1;
  FSR_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
}

// MMA with the 64-bit shared-memory descriptors given as 32-bit halves: the issuing lane then only does 32-bit adds per
// instruction (rebuilding 64-bit descriptors per MMA costs more issue time than a narrow MMA takes to execute)
__device__ __forceinline__ void umma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t acc) {
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
__device__ __forceinline__ uint32_t desc_lo32(uint32_t smem_addr, uint32_t lbo_bytes) { return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }
constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1

// Activations are stored as fp16 (pairs): a value beyond 65504 would become Inf, turn into NaN in the next layer and be
// swallowed by the next fmaxf-ReLU -- the result would be silently wrong.  Every convolution epilogue therefore tests what it
// is about to store; the engine turns the flag into an error (the fp32 reference would have carried on with finite numbers).
__device__ __forceinline__ void flag_unstorable(const float (&v)[8], unsigned* flags) {
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 8; ++i) bad |= !(fabsf(v[i]) <= 65504.0f);
  if (bad) atomicOr(flags, FSR_FLAG_PRED_NONFINITE);
}

struct ConvTcParams {
  int H, W, N;             // per-image extent and number of images in this launch
  int bw, bh, bn;          // M-tile box (bw*bh*bn == 128)
  int tiles_x, tiles_y;    // tiles per image row / column
  int ksz;                 // 1 or 3
  int kc;                  // 8-channel chunks per K-stage (even, <= 8)
  int s0, s1;              // K-stages per tap coming from src0 / src1
  int cout;                // real output channels (multiple of 8)
  int act;
  float alpha;
  int half;                // 16-bit format: 0 bf16, 1 fp16
  int stages;              // TMA ring depth (as many as fit: deep layers are L2-latency bound per K step)
  int pair;                // M tiles per CTA (1 or 2)
  int im;                  // 1: image-major tensors ("IM8" [C/8][H*W][N][8], maps of <= 64 pixels): an M tile is 128 images at
                           // one output pixel, each tap is ONE contiguous 2 KB run per channel chunk, out-of-range taps are skipped
  int ncap;                // images per pixel plane of IM8 tensors (allocation capacity)
  int up0;                 // im only: src0 is the level BELOW (H/2 x W/2) and the graph's 2x nearest upsampling is folded into this
                           // convolution: the 3 x 3 taps of an output pixel hit 2 x 2 pixels of the smaller map, so the taps that
                           // share a source pixel are summed on the host (weights per output parity, `wpack` layout below) and
                           // src0 costs 4 instead of 9 K steps per channel stage; the upsampled tensor is never written
  int parts;               // 1: plain 16-bit operands; 2: split fp16 (hi, lo) operands, three MMAs per product (tc_common.cuh)
  int nacc;                // parts == 2: the K steps of the hi*hi products rotate over `nacc` TMEM accumulators and the small
                           // hi*lo / lo*hi products go to one more, all summed in fp32 registers by the epilogue: the tensor core
                           // truncates when it adds into an accumulator, so error grows with the length of an accumulation chain
  float out_scale;         // parts == 2: the packed weights carry a power-of-two factor, undone here
  long long lo_off;        // parts == 2: elements between the hi and the lo tensor of out / res
  long long plane;         // pixels per CP8 plane of the output/residual tensors (N_capacity * H * W)
  const __nv_bfloat16* wpack;  // [n_tile][tap][stage][part][kc][BN][8]; up0: [n_tile][parity 2 x 2][source pixel 2 x 2][s0 stages][..]
                               // for src0, followed by [n_tile][tap][s1 stages][..] for src1
  const float* bias;           // [cout] or nullptr
  const __nv_bfloat16* res;    // CP8 residual or nullptr
  __nv_bfloat16* out;          // CP8 output
  unsigned* flags;             // FSR_FLAG_PRED_NONFINITE: an output does not fit fp16 (|v| > 65504, Inf or NaN)
};

template <int BN, int PAIR, int PARTS>
__global__ void __launch_bounds__(kConvThreads<PAIR>, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  griddep_launch_dependents();
  // stages are sized for this layer's K slice (kc 8-channel planes), so narrow layers fit several CTAs per SM and
  // their TMA latencies overlap.  With PAIR == 2 a CTA computes TWO M tiles (adjacent image blocks at the same
  // output position, so the same taps and weights): every weight slice fetched from L2 feeds two MMAs, one per
  // accumulator, each issued by its own warp.
  const int kABytesMax = PARTS * p.kc * 128 * 16;
  const int kBBytesMax = PARTS * p.kc * BN * 16;
  const int a_stage = PAIR * kABytesMax;
  uint8_t* smem_a = smem_raw;                                   // p.stages x pair x a_bytes
  uint8_t* smem_b = smem_raw + p.stages * a_stage;                // p.stages x b_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + p.stages * kBBytesMax);
  uint64_t* full = bars;                // [p.stages]
  uint64_t* empty = bars + p.stages;     // [p.stages]
  uint64_t* accum_full = bars + 2 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = p.ksz * p.ksz;
  const int stages_per_tap = p.s0 + p.s1;
  const uint32_t a_bytes = (uint32_t)kABytesMax;  // one TMA box: [part][kc][128 px][16 B]
  const uint32_t b_bytes = (uint32_t)kBBytesMax;

  // tile coordinates (IM8: tx = output pixel x, ty = output pixel y, tn = tile of 128 images)
  int t = blockIdx.x;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  const int tn0 = (t / p.tiles_y) * PAIR;
  const int n_tile = blockIdx.y;
  auto tap_ok = [&](int tap) {
    if (!p.im) return true;
    const int pad = p.ksz / 2;
    const int y = ty + tap / p.ksz - pad, x = tx + tap % p.ksz - pad;
    return y >= 0 && y < p.H && x >= 0 && x < p.W;
  };
  int n_taps_ok = 0;
  for (int tap = 0; tap < taps; ++tap) n_taps_ok += tap_ok(tap) ? 1 : 0;
  // up0: pixel (a, b) of the 2 x 2 block of the smaller map that the taps of output pixel (ty, tx) reach; -1 = outside the map
  // (exactly when all taps that map to it fall into the zero padding)
  auto fold_pix = [&](int q) {
    const int r = (ty >> 1) + (q >> 1) - (1 - (ty & 1)), c = (tx >> 1) + (q & 1) - (1 - (tx & 1));
    return (r >= 0 && r < (p.H >> 1) && c >= 0 && c < (p.W >> 1)) ? r * (p.W >> 1) + c : -1;
  };
  int n_fold_ok = 0;
  if (p.up0)
    for (int q = 0; q < 4; ++q) n_fold_ok += fold_pix(q) >= 0 ? 1 : 0;
  const int n_iters_cta = p.up0 ? n_fold_ok * p.s0 + n_taps_ok * p.s1 : n_taps_ok * stages_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.s1 > 0) tma_prefetch_desc(&tmA1);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], PAIR);
      }
      mbar_init(accum_full, PAIR);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, (uint32_t)(PAIR * BN * (PARTS == 2 ? p.nacc + 1 : 1)));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();  // the previous layer's output (and the buffers it still reads) are safe from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int pad = p.ksz / 2;
      int it = 0;
      const size_t stage_elems = (size_t)PARTS * p.kc * BN * 8;
      if (p.up0) {
        // src0 through the folded upsampling: one K run per reachable pixel of the smaller map, weights of this pixel's parity
        const int par = (ty & 1) * 2 + (tx & 1);
        for (int q = 0; q < 4; ++q) {
          const int pix = fold_pix(q);
          if (pix < 0) continue;
          for (int st = 0; st < p.s0; ++st, ++it) {
            const int s = it % p.stages;
            const uint32_t ph = (it / p.stages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], PAIR * a_bytes + b_bytes);
            for (int h = 0; h < PAIR; ++h)
              tma_load_5d(smem_a + s * a_stage + h * kABytesMax, &tmA0, &full[s], (tn0 + h) * 256, pix, st * p.kc, 0, 0);
            bulk_load_1d(smem_b + s * kBBytesMax, p.wpack + ((((size_t)n_tile * 4 + par) * 4 + q) * p.s0 + st) * stage_elems, b_bytes, &full[s]);
          }
        }
      }
      const __nv_bfloat16* wtaps = p.up0 ? p.wpack + (size_t)gridDim.y * 16 * p.s0 * stage_elems : p.wpack;
      const int st_begin = p.up0 ? p.s0 : 0, st_per_tap = p.up0 ? p.s1 : stages_per_tap;
      for (int tap = 0; tap < taps; ++tap) {
        if (!tap_ok(tap)) continue;
        const int dy = tap / p.ksz - pad, dx = tap % p.ksz - pad;
        for (int st = st_begin; st < stages_per_tap; ++st, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_expect_tx(&full[s], PAIR * a_bytes + b_bytes);
          const bool second = st >= p.s0;
          const CUtensorMap* tm = second ? &tmA1 : &tmA0;
          const int chunk0 = (second ? st - p.s0 : st) * p.kc;
          for (int h = 0; h < PAIR; ++h) {  // a block past the last image reads as zeros (TMA bounds fill)
            uint8_t* dst = smem_a + s * a_stage + h * kABytesMax;
            if (p.im) tma_load_5d(dst, tm, &full[s], (tn0 + h) * 256, (ty + dy) * p.W + tx + dx, chunk0, 0, 0);
            else tma_load_5d(dst, tm, &full[s], 2 * (tx * p.bw + dx), ty * p.bh + dy, (tn0 + h) * p.bn, chunk0, 0);
          }
          const __nv_bfloat16* wsrc = wtaps + ((size_t)(n_tile * taps + tap) * st_per_tap + (st - st_begin)) * stage_elems;
          bulk_load_1d(smem_b + s * kBBytesMax, wsrc, b_bytes, &full[s]);
        }
      }
    }
  } else if (warp <= PAIR) {
    // ===================== MMA issuers: warp 1 + h feeds accumulator h (warp-uniform loop, one elected lane issues) =====================
    const int h = warp - 1;
    {
      const uint32_t idesc = idesc_16(128, BN, p.half);
      const uint32_t a_lo0 = desc_lo32(smem_u32(smem_a + h * kABytesMax), 128 * 16);
      const uint32_t b_lo0 = desc_lo32(smem_u32(smem_b), BN * 16);
      const uint32_t a_step = (uint32_t)a_stage >> 4, b_step = (uint32_t)kBBytesMax >> 4;
      const uint32_t d = tmem_base + h * BN * (PARTS == 2 ? p.nacc + 1 : 1);  // split mode: accumulators [d, d + (nacc + 1) * BN)
      const int kpairs = p.kc / 2;
      const uint32_t a_part = (uint32_t)(p.kc * 128 * 16) >> 4, b_part = (uint32_t)(p.kc * BN * 16) >> 4;
      const bool leader = elect_one();
      int s = 0;
      uint32_t ph = 0, a_lo = a_lo0, b_lo = b_lo0;
      for (int it = 0; it < n_iters_cta; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (leader) {
          // one MMA consumes two 8-channel planes: LBO = plane stride, SBO = 8 rows x 16 B
          if (PARTS == 2) {
            // stage `it` adds its hi*hi products to accumulator it % nacc; + a_lo * b_hi + a_hi * b_lo go to accumulator nacc
            const uint32_t dm = d + (uint32_t)(it % p.nacc) * BN, ds = d + (uint32_t)p.nacc * BN;
            for (int j = 0; j < kpairs; ++j) {
              const uint32_t aj = a_lo + j * ((2 * 128 * 16) >> 4), bj = b_lo + j * ((2 * BN * 16) >> 4);
              umma_lo(dm, aj, bj, kDescHi, idesc, (it >= p.nacc || j > 0) ? 1u : 0u);
              umma_lo(ds, aj + a_part, bj, kDescHi, idesc, (it > 0 || j > 0) ? 1u : 0u);
              umma_lo(ds, aj, bj + b_part, kDescHi, idesc, 1u);
            }
          } else {
            for (int j = 0; j < kpairs; ++j)
              umma_lo(d, a_lo + j * ((2 * 128 * 16) >> 4), b_lo + j * ((2 * BN * 16) >> 4), kDescHi, idesc, (it > 0 || j > 0) ? 1u : 0u);
          }
          umma_commit(&empty[s]);  // frees the smem slot once these MMAs have read it
          if (it == n_iters_cta - 1) umma_commit(accum_full);
        }
        __syncwarp();
        a_lo += a_step;
        b_lo += b_step;
        if (++s == p.stages) { s = 0; ph ^= 1; a_lo = a_lo0; b_lo = b_lo0; }
      }
    }
  } else {
    // ===================== epilogue (4 warps, one per TMEM lane quarter) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;
    const int lx = m % p.bw;
    const int ly = (m / p.bw) % p.bh;
    const int ln = m / (p.bw * p.bh);
    for (int h = 0; h < PAIR; ++h) {
      const int tn = tn0 + h;
      const int n_img = p.im ? tn * 128 + m : tn * p.bn + ln;
      const bool valid = n_img < p.N;
      const long long pix = p.im ? ((long long)(ty * p.W + tx) * p.ncap + n_img)
                                 : ((long long)n_img * p.H + (ty * p.bh + ly)) * p.W + (tx * p.bw + lx);
      // the first tile's residual is fetched while the MMAs still run
      uint4 resv[BN / 8];
      if (p.res && valid && PARTS == 1) {
#pragma unroll
        for (int c = 0; c < BN / 8; ++c) {
          const int co = n_tile * BN + c * 8;
          if (co < p.cout) resv[c] = __ldg(reinterpret_cast<const uint4*>(p.res + ((long long)(co >> 3) * p.plane + pix) * 8));
        }
      }
      if (h == 0) {
        mbar_wait(accum_full, 0);
        tc_fence_after();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + h * BN * (PARTS == 2 ? p.nacc + 1 : 1);
      const int n_extra = PARTS == 2 ? (n_iters_cta < p.nacc ? n_iters_cta : p.nacc) : 0;  // accumulators beyond the first
#pragma unroll
      for (int c = 0; c < BN / 8; ++c) {
        float v[8];
        tmem_ld8(taddr + c * 8, v);
        if (PARTS == 2) {
          // all partial accumulators of this column chunk in flight at once, ONE wait; fixed summation order: main
          // accumulators 1 .. used - 1, then the small-product accumulator (index nacc)
          float u[7][8];
#pragma unroll
          for (int a = 1; a <= 7; ++a)
            if (a <= n_extra) tmem_ld8(taddr + (uint32_t)(a == n_extra ? p.nacc : a) * BN + c * 8, u[a - 1]);
          tmem_ld_wait();
#pragma unroll
          for (int a = 1; a <= 7; ++a)
            if (a <= n_extra) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += u[a - 1][i];
            }
        } else {
          tmem_ld_wait();
        }
        const int co = n_tile * BN + c * 8;
        if (valid && co < p.cout) {
          const long long off = ((long long)(co >> 3) * p.plane + pix) * 8;
          if (PARTS == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] *= p.out_scale;
          }
          if (p.bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + co));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + co + 4));
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
            v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
          }
          if (p.res) {
            float r[8];
            if (PARTS == 2)
              join_x8(__ldg(reinterpret_cast<const uint4*>(p.res + off)), __ldg(reinterpret_cast<const uint4*>(p.res + p.lo_off + off)), r);
            else
              unpack_x8(resv[c], r, p.half);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += r[i];
          }
          if (p.act == FSR_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.0f);
          } else if (p.act == FSR_ACT_LEAKY) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.0f ? v[i] : v[i] * p.alpha;
          }
          flag_unstorable(v, p.flags);
          if (PARTS == 2) {
            uint4 hi, lo;
            split_x8(v, hi, lo);
            *reinterpret_cast<uint4*>(p.out + off) = hi;
            *reinterpret_cast<uint4*>(p.out + p.lo_off + off) = lo;
          } else {
            *reinterpret_cast<uint4*>(p.out + off) = pack_x8(v, p.half);
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)(PAIR * BN * (PARTS == 2 ? p.nacc + 1 : 1)));
  }
}


// ---- persistent row-box variant for the wide, shallow levels ------------------------------------------------------
// For 3x3 layers on 32- and 16-pixel-wide maps whose packed weights fit in shared memory.  One CTA per SM loops over M tiles:
//   * the weights are loaded once per CTA;
//   * pixels are enumerated in a PADDED pitch (image width + 8): an M tile is 128 consecutive positions of that
//     enumeration, and ONE halo box (pitch pixels wide starting at x = -1, all rows the tile touches + 1 above / below) is
//     fetched per (M tile, 32-channel group).  Because the box rows are `pitch` pixels apart in shared memory too, tap
//     (ky, kx) of every position is the box shifted by (ky * pitch + kx) * 16 bytes: nine taps from one fetch instead of
//     nine (or three) fetches.  The price is 25 % (W = 32) / 50 % (W = 16) of positions that are padding and are dropped
//     in the epilogue; the kernel is fetch-bound, not MMA-bound (N <= 64);
//   * accumulators are double-buffered in TMEM, so the epilogue of tile t overlaps the MMAs of tile t + 1, and the
//     per-CTA set-up (TMEM allocation, barrier init, first TMA latency) is paid once per SM instead of once per tile.
constexpr int kRowMaxStages = 8;
constexpr int kRowsEpiSets = 4;                       // epilogue warp sets (one per accumulator buffer)
constexpr int kRowsIssuers = 2;                       // MMA-issuing warps (tiles dealt round-robin)
constexpr int kRowsThreads = 32 + 32 * kRowsIssuers + 128 * kRowsEpiSets;  // producer warp, MMA warps, one epilogue set per accumulator buffer
template <int BN>
constexpr int kRowsTmemCols = kRowsEpiSets * BN <= 64 ? 64 : kRowsEpiSets * BN <= 128 ? 128 : kRowsEpiSets * BN <= 256 ? 256 : 512;

struct ConvRowsParams {
  int H, W, N;
  int pitch;           // W + 8
  unsigned pitch_magic;  // ceil(2^32 / pitch): position / pitch as a multiply-high
  int box_rows;        // rows of the halo box (rows an M tile can touch + 2)
  int tiles_per_img;   // H * pitch / 128
  int n_mtiles;        // N * tiles_per_img
  int kc;              // 8-channel planes per channel group (2 or 4)
  int g0, g1;          // channel groups coming from src0 / src1
  int cout, act;
  float alpha;
  int half;
  int stages;          // halo-box ring depth (as many as fit beside the weights)
  int parts;           // 1: plain 16-bit operands; 2: split fp16 (hi, lo) operands, three MMAs per product
  int sets;            // accumulator buffers / epilogue warp sets in use (<= kRowsEpiSets)
  int nacc;            // parts == 2: hi*hi products rotate over `nacc` accumulators per buffer, the small products use one more
                       // (see ConvTcParams::nacc); a buffer is then (nacc + 1) * BN TMEM columns
  float out_scale;     // parts == 2: power-of-two factor carried by the packed weights, undone in the epilogue
  long long lo_off;    // parts == 2: elements between the hi and the lo tensor of out / res
  unsigned long long* stats;  // FSR_ROWS_STATS: per-CTA clock totals of the pipeline roles (diagnostics)
  long long plane;     // pixels per CP8 plane of the output/residual tensors
  const __nv_bfloat16* wpack;  // [part][tap][group][kc][BN][8]
  int w_bytes;
  const float* bias;
  const __nv_bfloat16* res;
  __nv_bfloat16* out;
  unsigned* flags;     // FSR_FLAG_PRED_NONFINITE: an output does not fit fp16
};

template <int BN, int PARTS>
__global__ void __launch_bounds__(kRowsThreads, 1)
conv_rows_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1, const ConvRowsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  griddep_launch_dependents();
  const int plane_b = p.box_rows * p.pitch * 16;    // one 8-channel plane of a halo box
  const int box_bytes = PARTS * p.kc * plane_b;   // [part][kc][box_rows * pitch px][16 B], one TMA box
  const int stage_bytes = (box_bytes + 2 * p.pitch * 16 + 127) & ~127;  // slack: the last taps of padding positions read past the box
  const int w_round = (p.w_bytes + 1023) & ~1023;
  uint8_t* smem_w = smem_raw;
  uint8_t* smem_a = smem_raw + w_round;              // p.stages boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_a + p.stages * stage_bytes);
  uint64_t* w_full = bars;
  uint64_t* full = bars + 1;                 // [p.stages]
  uint64_t* empty = full + p.stages;         // [p.stages]
  uint64_t* acc_full = empty + p.stages;                // [kRowsEpiSets]
  uint64_t* acc_empty = acc_full + kRowsEpiSets;        // [kRowsEpiSets]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kRowsEpiSets);
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = p.g0 + p.g1;
  const int buf_cols = (PARTS == 2 ? p.nacc + 1 : 1) * BN;  // TMEM columns of one accumulator buffer
  const uint32_t tmem_cols = PARTS == 2 ? (uint32_t)(p.sets * buf_cols) : (uint32_t)kRowsTmemCols<BN>;  // host: a power of two >= 32

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.g1 > 0) tma_prefetch_desc(&tmA1);
  }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(w_full, 1);
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], 1);
      }
      for (int i = 0; i < kRowsEpiSets; ++i) {
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  if (threadIdx.x < BN) s_bias[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.0f;
  // the slack behind every box is only ever multiplied into discarded padding positions, but must not hold NaN/Inf patterns
  for (int i = threadIdx.x; i < p.stages * stage_bytes / 16; i += kRowsThreads) reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the packed weights are constants: their load may start before the previous layer has finished
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(w_full, (uint32_t)p.w_bytes);
    bulk_load_1d(smem_w, p.wpack, (uint32_t)p.w_bytes, w_full);
  }
  griddep_wait();  // the previous layer's output (and the buffers it still reads) are safe from here on

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 1;
      long long st_wait = 0;
      const long long c_start = p.stats ? clock64() : 0;
      int img = (int)(blockIdx.x / p.tiles_per_img), t = (int)(blockIdx.x % p.tiles_per_img);
      const int step_img = (int)(gridDim.x / p.tiles_per_img), step_t = (int)(gridDim.x % p.tiles_per_img);
      for (int mt = blockIdx.x; mt < p.n_mtiles; mt += gridDim.x) {
        const int r0 = (int)__umulhi((unsigned)(t * 128), p.pitch_magic);  // image row of the tile's first position
        for (int g = 0; g < groups; ++g) {
          const long long c_a = p.stats ? clock64() : 0;
          mbar_wait(&empty[s], ph);
          if (p.stats) st_wait += clock64() - c_a;
          mbar_expect_tx(&full[s], (uint32_t)box_bytes);
          const bool second = g >= p.g0;
          const CUtensorMap* tm = second ? &tmA1 : &tmA0;
          const int chunk0 = (second ? g - p.g0 : g) * p.kc;
          tma_load_5d(smem_a + s * stage_bytes, tm, &full[s], -2, r0 - 1, img, chunk0, 0);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        t += step_t;
        img += step_img;
        if (t >= p.tiles_per_img) { t -= p.tiles_per_img; ++img; }
      }
      if (p.stats) {
        p.stats[blockIdx.x * 16 + 0] = (unsigned long long)(clock64() - c_start);
        p.stats[blockIdx.x * 16 + 1] = (unsigned long long)st_wait;
      }
    }
  } else if (warp <= kRowsIssuers) {
    // ===================== MMA issuers: warp 1 + e takes tiles e, e + kRowsIssuers, ...; tile nt uses accumulator buffer nt % kRowsEpiSets =====================
    // One thread issues an N <= 64 MMA more slowly than the tensor pipe retires it (descriptor arithmetic + issue latency),
    // so the tiles are dealt to kRowsIssuers issuing warps whose MMAs interleave in the pipe.
    const int e = warp - 1;
    const uint32_t idesc = idesc_16(128, BN, p.half);
    mbar_wait(w_full, 0);
    const uint32_t a_lo0 = desc_lo32(smem_u32(smem_a), (uint32_t)plane_b);
    const uint32_t b_lo0 = desc_lo32(smem_u32(smem_w), BN * 16);
    const uint32_t st_u = (uint32_t)stage_bytes >> 4, row_u = (uint32_t)p.pitch;  // stage / image-row strides in 16-byte units
    const uint32_t kpl_u = (uint32_t)(2 * plane_b) >> 4;                            // one MMA = two channel planes
    const uint32_t wtap_u = (uint32_t)(groups * p.kc * BN * 16) >> 4, wgrp_u = (uint32_t)(p.kc * BN * 16) >> 4;
    const uint32_t a_part = (uint32_t)(p.kc * plane_b) >> 4, w_part = 9u * wtap_u;  // hi -> lo operand distances
    const int kpairs = p.kc / 2;
    const bool leader = elect_one();
    int nt = e;
    long long w_acc = 0, w_full = 0, n_t = 0;
    const long long c_start = p.stats ? clock64() : 0;
    int t = (int)((blockIdx.x + e * gridDim.x) % p.tiles_per_img);
    const int step_t = (int)((kRowsIssuers * gridDim.x) % p.tiles_per_img);
    int s = e * groups;
    uint32_t ph = 0;
    while (s >= p.stages) { s -= p.stages; ph ^= 1; }
    for (int mt = blockIdx.x + e * gridDim.x; mt < p.n_mtiles; mt += kRowsIssuers * gridDim.x, nt += kRowsIssuers) {
      const int ab = nt % p.sets;
      const uint32_t d = tmem_base + ab * buf_cols;
      const uint32_t c0 = (uint32_t)(t * 128) - __umulhi((unsigned)(t * 128), p.pitch_magic) * (uint32_t)p.pitch;  // column (in the padded pitch) of the tile's first position
      long long c_a = p.stats ? clock64() : 0;
      mbar_wait(&acc_empty[ab], ((nt / p.sets) & 1) ^ 1);
      if (p.stats) { w_acc += clock64() - c_a; ++n_t; }
      for (int g = 0; g < groups; ++g) {
        c_a = p.stats ? clock64() : 0;
        mbar_wait(&full[s], ph);
        if (p.stats) w_full += clock64() - c_a;
        tc_fence_after();
        if (leader) {
          const uint32_t a_s = a_lo0 + (uint32_t)s * st_u + c0;
          const uint32_t b_g = b_lo0 + (uint32_t)g * wgrp_u;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t a_t = a_s + (tap / 3) * row_u + (tap % 3);
            const uint32_t b_t = b_g + tap * wtap_u;
            if (PARTS == 2) {
              // hi*hi products of step q = g * 9 + tap go to accumulator q % nacc (nacc <= 9: every accumulator is written in
              // the first group); + a_lo * w_hi + a_hi * w_lo to accumulator nacc
              const int q = g * 9 + tap;
              const uint32_t dm = d + (uint32_t)(q % p.nacc) * BN, ds = d + (uint32_t)p.nacc * BN;
              umma_lo(dm, a_t, b_t, kDescHi, idesc, q >= p.nacc ? 1u : 0u);
              if (kpairs > 1) umma_lo(dm, a_t + kpl_u, b_t + ((2 * BN * 16) >> 4), kDescHi, idesc, 1u);
              umma_lo(ds, a_t + a_part, b_t, kDescHi, idesc, q > 0 ? 1u : 0u);
              umma_lo(ds, a_t, b_t + w_part, kDescHi, idesc, 1u);
              if (kpairs > 1) {
                umma_lo(ds, a_t + a_part + kpl_u, b_t + ((2 * BN * 16) >> 4), kDescHi, idesc, 1u);
                umma_lo(ds, a_t + kpl_u, b_t + w_part + ((2 * BN * 16) >> 4), kDescHi, idesc, 1u);
              }
            } else {
              umma_lo(d, a_t, b_t, kDescHi, idesc, (g > 0 || tap > 0) ? 1u : 0u);
              if (kpairs > 1) umma_lo(d, a_t + kpl_u, b_t + ((2 * BN * 16) >> 4), kDescHi, idesc, 1u);
            }
          }
          umma_commit(&empty[s]);
          if (g == groups - 1) umma_commit(&acc_full[ab]);
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      // the other issuers' boxes of the ring (shared by all issuers, filled in tile order)
      s += (kRowsIssuers - 1) * groups;
      while (s >= p.stages) { s -= p.stages; ph ^= 1; }
      t += step_t;
      if (t >= p.tiles_per_img) t -= p.tiles_per_img;
    }
    if (p.stats && lane == 0) {
      p.stats[blockIdx.x * 16 + 2 + e * 4] = (unsigned long long)(clock64() - c_start);
      p.stats[blockIdx.x * 16 + 3 + e * 4] = (unsigned long long)w_acc;
      p.stats[blockIdx.x * 16 + 4 + e * 4] = (unsigned long long)w_full;
      p.stats[blockIdx.x * 16 + 5 + e * 4] = (unsigned long long)n_t;
    }
  } else {
    // ===================== epilogue: kRowsEpiSets sets of 4 warps, set e drains accumulator buffer e =====================
    // One warp retires a dependent instruction every few cycles, so a single set of four warps (one per TMEM lane
    // quarter) needs longer per tile than the tile's MMAs; two sets work on alternate tiles.
    const int set = (warp - 1 - kRowsIssuers) >> 2;
    const int q = warp & 3;
    const int m = q * 32 + lane;
    // (image, tile-in-image) advance without divisions; warp sets beyond p.sets have no accumulator buffer and no work
    const int first = set < p.sets ? blockIdx.x + set * gridDim.x : p.n_mtiles, step = p.sets * gridDim.x;
    int img = first / p.tiles_per_img, t = first % p.tiles_per_img;
    const int step_img = step / p.tiles_per_img, step_t = step % p.tiles_per_img;
    int nt = set;
    long long e_wait = 0, e_ld = 0;
    const long long e_start = p.stats ? clock64() : 0;
    for (int mt = first; mt < p.n_mtiles; mt += step, nt += p.sets) {
      const int ab = set;
      const int pos = t * 128 + m;
      const int y = (int)__umulhi((unsigned)pos, p.pitch_magic), x = pos - y * p.pitch;
      const bool valid = x < p.W && y < p.H;   // padding positions are dropped
      const long long pix = ((long long)img * p.H + y) * p.W + x;
      uint4 resv[BN / 8];  // fetched while the tile's MMAs still run
      if (p.res && valid && PARTS == 1) {
#pragma unroll
        for (int c = 0; c < BN / 8; ++c) resv[c] = __ldg(reinterpret_cast<const uint4*>(p.res + ((long long)c * p.plane + pix) * 8));
      }
      const long long c_a = p.stats ? clock64() : 0;
      mbar_wait(&acc_full[ab], (nt / p.sets) & 1);
      if (p.stats) e_wait += clock64() - c_a;
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * buf_cols;
#pragma unroll
      for (int c32 = 0; c32 < BN / 32; ++c32) {
        float v[32];
        tmem_ld32(taddr + c32 * 32, v);
        tmem_ld_wait();
        if (PARTS == 2) {
          // sum of the partial accumulators in a fixed order: main 1 .. nacc - 1, then the small-product accumulator
          for (int a = 1; a <= p.nacc; ++a) {
            float u[32];
            tmem_ld32(taddr + (uint32_t)a * BN + c32 * 32, u);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += u[i];
          }
        }
        if (c32 == BN / 32 - 1) {
          // the accumulator is in registers: hand the buffer back before the arithmetic and the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[ab]);
          if (p.stats) e_ld += clock64() - c_a;
        }
        if (valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int co = c32 * 32 + c * 8;
            const long long off = ((long long)(co >> 3) * p.plane + pix) * 8;
            float o[8];
            {
              const float4 b0 = *reinterpret_cast<const float4*>(s_bias + co), b1 = *reinterpret_cast<const float4*>(s_bias + co + 4);
              const float sc = p.out_scale;  // 1 unless parts == 2
              o[0] = fmaf(v[c * 8], sc, b0.x); o[1] = fmaf(v[c * 8 + 1], sc, b0.y); o[2] = fmaf(v[c * 8 + 2], sc, b0.z); o[3] = fmaf(v[c * 8 + 3], sc, b0.w);
              o[4] = fmaf(v[c * 8 + 4], sc, b1.x); o[5] = fmaf(v[c * 8 + 5], sc, b1.y); o[6] = fmaf(v[c * 8 + 6], sc, b1.z); o[7] = fmaf(v[c * 8 + 7], sc, b1.w);
            }
            if (p.res) {
              float r[8];
              if (PARTS == 2)
                join_x8(__ldg(reinterpret_cast<const uint4*>(p.res + off)), __ldg(reinterpret_cast<const uint4*>(p.res + p.lo_off + off)), r);
              else
                unpack_x8(resv[c32 * 4 + c], r, p.half);
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] += r[i];
            }
            if (p.act == FSR_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = fmaxf(o[i], 0.0f);
            } else if (p.act == FSR_ACT_LEAKY) {
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = o[i] > 0.0f ? o[i] : o[i] * p.alpha;
            }
            flag_unstorable(o, p.flags);
            if (PARTS == 2) {
              uint4 hi, lo;
              split_x8(o, hi, lo);
              *reinterpret_cast<uint4*>(p.out + off) = hi;
              *reinterpret_cast<uint4*>(p.out + p.lo_off + off) = lo;
            } else {
              *reinterpret_cast<uint4*>(p.out + off) = pack_x8(o, p.half);
            }
          }
        }
      }
      t += step_t;
      img += step_img;
      if (t >= p.tiles_per_img) { t -= p.tiles_per_img; ++img; }
    }
    if (p.stats && set == 0 && q == 0 && lane == 0) {
      p.stats[blockIdx.x * 16 + 10] = (unsigned long long)(clock64() - e_start);
      p.stats[blockIdx.x * 16 + 11] = (unsigned long long)e_wait;
      p.stats[blockIdx.x * 16 + 12] = (unsigned long long)e_ld;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ---- CP8 helper kernels (memory-bound, 16-byte vectors) -----------------------------------------------------

// concat of up to two 1-channel fp32 NHWC tensors -> CP8 bf16 with `chunks` 8-channel planes (zero padded)
__global__ void pack_small_kernel(const float* __restrict__ s0, int c0, const float* __restrict__ s1, int c1,
                                  __nv_bfloat16* __restrict__ dst, long long n_pix, long long plane, int chunks, int half, int parts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix * chunks; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % n_pix;
    const int ch = (int)(i / n_pix);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = ch * 8 + k;
      float x = 0.0f;
      if (c < c0) x = __ldg(s0 + pix * c0 + c);
      else if (c < c0 + c1) x = __ldg(s1 + pix * c1 + (c - c0));
      v[k] = x;
    }
    if (parts == 2) {
      uint4 hi, lo;
      split_x8(v, hi, lo);
      *reinterpret_cast<uint4*>(dst + ((long long)ch * plane + pix) * 8) = hi;
      *reinterpret_cast<uint4*>(dst + ((long long)(chunks + ch) * plane + pix) * 8) = lo;
    } else {
      *reinterpret_cast<uint4*>(dst + ((long long)ch * plane + pix) * 8) = pack_x8(v, half);
    }
  }
}

// element offset (in 16-bit elements) of 8-channel chunk `ch`, image `img`, pixel (y, x): CP8 [C/8][N][H][W][8] or, for
// small maps, IM8 [C/8][H*W][N][8]; `ncap` = images per plane as allocated
__device__ __forceinline__ long long chunk_off(int im, int ch, long long img, int y, int x, int H, int W, long long ncap) {
  if (im) return (((long long)ch * H * W + (long long)y * W + x) * ncap + img) * 8;
  return (((long long)ch * ncap + img) * H * W + (long long)y * W + x) * 8;
}

__device__ __forceinline__ uint4 x8_max(const uint4& a, const uint4& b, int half) {
  uint4 r;
  if (half) {
    const __half2* pa = reinterpret_cast<const __half2*>(&a);
    const __half2* pb = reinterpret_cast<const __half2*>(&b);
    __half2* pr = reinterpret_cast<__half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  } else {
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
  }
  return r;
}

// k x k pooling (stride k): one thread per (chunk, output pixel); source and destination may use different layouts
__global__ void pool_cp8_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int chunks, int n_img,
                                int Hin, int Win, int k, int mode, long long ncap_in, long long ncap_out, int half, int im_in, int im_out,
                                int parts, int aux) {
  const int Hout = Hin / k, Wout = Win / k;
  const long long n_out = (long long)n_img * Hout * Wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out * chunks; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i / n_out);
    long long pix = i % n_out;
    int X, Y;
    long long img;
    if (im_out) {  // image fastest: coalesced stores in the destination layout
      img = pix % n_img;
      pix /= n_img;
      X = (int)(pix % Wout);
      Y = (int)(pix / Wout);
    } else {
      X = (int)(pix % Wout);
      const long long t2 = pix / Wout;
      Y = (int)(t2 % Hout);
      img = t2 / Hout;
    }
    const long long o = chunk_off(im_out, ch, img, Y, X, Hout, Wout, ncap_out);
    if (parts == 2) {
      // split tensors: pool the joined values (hi + lo is exact in fp32), then split again
      const long long lo_in = (long long)chunks * ncap_in * Hin * Win * 8, lo_out = (long long)chunks * ncap_out * Hout * Wout * 8;
      if (mode == FSR_POOL_PICK) {
        const long long a = chunk_off(im_in, ch, img, Y * k + aux, X * k + aux, Hin, Win, ncap_in);
        *reinterpret_cast<uint4*>(dst + o) = __ldg(reinterpret_cast<const uint4*>(src + a));
        *reinterpret_cast<uint4*>(dst + lo_out + o) = __ldg(reinterpret_cast<const uint4*>(src + lo_in + a));
        continue;
      }
      float acc[8];
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) {
          const long long a = chunk_off(im_in, ch, img, Y * k + dy, X * k + dx, Hin, Win, ncap_in);
          float f[8];
          join_x8(__ldg(reinterpret_cast<const uint4*>(src + a)), __ldg(reinterpret_cast<const uint4*>(src + lo_in + a)), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = (dy == 0 && dx == 0) ? f[j] : (mode == 0 ? fmaxf(acc[j], f[j]) : acc[j] + f[j]);
        }
      if (mode != 0) {
        const float inv = 1.0f / (float)(k * k);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= inv;
      }
      uint4 hi, lo;
      split_x8(acc, hi, lo);
      *reinterpret_cast<uint4*>(dst + o) = hi;
      *reinterpret_cast<uint4*>(dst + lo_out + o) = lo;
      continue;
    }
    if (mode == FSR_POOL_PICK) {
      *reinterpret_cast<uint4*>(dst + o) = __ldg(reinterpret_cast<const uint4*>(src + chunk_off(im_in, ch, img, Y * k + aux, X * k + aux, Hin, Win, ncap_in)));
    } else if (mode == 0) {
      uint4 acc = __ldg(reinterpret_cast<const uint4*>(src + chunk_off(im_in, ch, img, Y * k, X * k, Hin, Win, ncap_in)));
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx)
          acc = x8_max(acc, __ldg(reinterpret_cast<const uint4*>(src + chunk_off(im_in, ch, img, Y * k + dy, X * k + dx, Hin, Win, ncap_in))), half);
      *reinterpret_cast<uint4*>(dst + o) = acc;
    } else {
      float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) {
          float f[8];
          unpack_x8(__ldg(reinterpret_cast<const uint4*>(src + chunk_off(im_in, ch, img, Y * k + dy, X * k + dx, Hin, Win, ncap_in))), f, half);
#pragma unroll
          for (int j = 0; j < 8; ++j) s[j] += f[j];
        }
      const float inv = 1.0f / (float)(k * k);
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] *= inv;
      *reinterpret_cast<uint4*>(dst + o) = pack_x8(s, half);
    }
  }
}

// `chunks` counts the planes of ONE part; mode FSR_UP_NEAREST copies every part's planes, the bilinear modes interpolate the
// (joined) values in fp32 and store them in the tensor's format again
__global__ void upsample_cp8_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int chunks, int n_img,
                                    int Hin, int Win, int f, long long ncap_in, long long ncap_out, int im_in, int im_out, int mode,
                                    int half, int parts) {
  const int Hout = Hin * f, Wout = Win * f;
  const long long n_out = (long long)n_img * Hout * Wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out * chunks; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i / n_out);
    long long pix = i % n_out;
    int X, Y;
    long long img;
    if (im_out) {
      img = pix % n_img;
      pix /= n_img;
      X = (int)(pix % Wout);
      Y = (int)(pix / Wout);
    } else {
      X = (int)(pix % Wout);
      const long long t2 = pix / Wout;
      Y = (int)(t2 % Hout);
      img = t2 / Hout;
    }
    const long long o = chunk_off(im_out, ch, img, Y, X, Hout, Wout, ncap_out);
    const long long lo_in = (long long)chunks * ncap_in * Hin * Win * 8, lo_out = (long long)chunks * ncap_out * Hout * Wout * 8;
    if (mode == FSR_UP_NEAREST) {
      const long long a = chunk_off(im_in, ch, img, Y / f, X / f, Hin, Win, ncap_in);
      *reinterpret_cast<uint4*>(dst + o) = __ldg(reinterpret_cast<const uint4*>(src + a));
      if (parts == 2) *reinterpret_cast<uint4*>(dst + lo_out + o) = __ldg(reinterpret_cast<const uint4*>(src + lo_in + a));
      continue;
    }
    int y0, y1, x0, x1;
    float wy, wx;
    up_linear_coord(Y, f, Hin, mode, y0, y1, wy);
    up_linear_coord(X, f, Win, mode, x0, x1, wx);
    float q[4][8];
    const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const long long a = chunk_off(im_in, ch, img, ys[t], xs[t], Hin, Win, ncap_in);
      if (parts == 2) join_x8(__ldg(reinterpret_cast<const uint4*>(src + a)), __ldg(reinterpret_cast<const uint4*>(src + lo_in + a)), q[t]);
      else unpack_x8(__ldg(reinterpret_cast<const uint4*>(src + a)), q[t], half);
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float top = q[0][j] + (q[1][j] - q[0][j]) * wx, bot = q[2][j] + (q[3][j] - q[2][j]) * wx;
      r[j] = top + (bot - top) * wy;
    }
    if (parts == 2) {
      uint4 hi, lo;
      split_x8(r, hi, lo);
      *reinterpret_cast<uint4*>(dst + o) = hi;
      *reinterpret_cast<uint4*>(dst + lo_out + o) = lo;
    } else {
      *reinterpret_cast<uint4*>(dst + o) = pack_x8(r, half);
    }
  }
}

__global__ void eltwise_cp8_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                   __nv_bfloat16* __restrict__ dst, long long n_vec, int act, float alpha, float beta, int half, long long lo_vec) {
  // lo_vec != 0: split tensors, the lo part lies lo_vec 16-byte vectors behind the hi part
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    if (lo_vec) join_x8(__ldg(reinterpret_cast<const uint4*>(a) + i), __ldg(reinterpret_cast<const uint4*>(a) + lo_vec + i), x);
    else unpack_x8(__ldg(reinterpret_cast<const uint4*>(a) + i), x, half);
    if (b) {
      if (lo_vec) join_x8(__ldg(reinterpret_cast<const uint4*>(b) + i), __ldg(reinterpret_cast<const uint4*>(b) + lo_vec + i), y);
      else unpack_x8(__ldg(reinterpret_cast<const uint4*>(b) + i), y, half);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] += y[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = apply_act2(x[j], act, alpha, beta);
    if (lo_vec) {
      uint4 hi, lo;
      split_x8(x, hi, lo);
      reinterpret_cast<uint4*>(dst)[i] = hi;
      reinterpret_cast<uint4*>(dst)[lo_vec + i] = lo;
    } else {
      reinterpret_cast<uint4*>(dst)[i] = pack_x8(x, half);
    }
  }
}

// CP8 / IM8 16-bit -> NHWC fp32 (debug / parity reads of intermediate tensors)
__global__ void cp8_to_nhwc_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n_img, int H, int W,
                                   long long ncap, int C, int half, int im, long long lo_off) {
  const long long total = n_img * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long pix = i / C;
    const int x = (int)(pix % W);
    pix /= W;
    const int y = (int)(pix % H);
    const long long img = pix / H;
    const __nv_bfloat16* e = src + chunk_off(im, c >> 3, img, y, x, H, W, ncap) + (c & 7);
    float v = half ? __half2float(*reinterpret_cast<const __half*>(e)) : __bfloat162float(*e);
    if (lo_off) v += __half2float(*reinterpret_cast<const __half*>(e + lo_off));
    dst[i] = v;
  }
}

inline int grid_for(long long total, int threads = 256) {
  long long b = (total + threads - 1) / threads;
  const long long cap = (long long)current_sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <int BN>
size_t conv_smem_bytes(int kc, int stages, int pair) {  // kc = 8-channel planes per stage, all parts counted
  return (size_t)stages * pair * (kc * 128 * 16) + (size_t)stages * (kc * BN * 16) + (2 * stages + 1) * sizeof(uint64_t) + 16;
}
template <int BN>
int conv_stages(int kc, int n_iters, long long n_ctas, int pair, bool one_cta_per_sm = false) {
  // few CTAs (deep, narrow levels): one CTA per SM with a deep ring, each K step is L2-latency bound;
  // many CTAs: keep two or more CTAs per SM so that their prologues and epilogues overlap
  const size_t budget = (one_cta_per_sm || n_ctas <= (pair > 1 ? current_sm_count() : 2 * current_sm_count())) ? 200 * 1024 : 100 * 1024;
  int st = (int)(budget / ((size_t)pair * kc * 128 * 16 + (size_t)kc * BN * 16));
  st = st > kMaxStages ? kMaxStages : st;
  st = st > n_iters ? n_iters : st;
  return st < 2 ? 2 : st;
}

}  // namespace

// ---- host-side: tensor maps and launchers ------------------------------------------------------------------

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  FSR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
  if (!ptr || qres != cudaDriverEntryPointSuccess) throw Error(FSR_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// 5-D tensor map over a CP8 activation tensor: dims (8, W, H, N, C/8), element strides in bytes
// (2, 16, W*16, H*W*16, plane*16); box (8, bw, bh, bn, kc); no swizzle; out-of-bounds elements read as zero.
CUtensorMap make_cp8_tensor_map(const void* base, int W, int H, int N, int chunks, long long plane, int bw, int bh, int bn, int kc) {
  CUtensorMap m;
  cuuint64_t dims[5] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)chunks};
  cuuint64_t strides[4] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)plane * 16};
  cuuint32_t box[5] = {8, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn, (cuuint32_t)kc};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(FSR_E_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

// The same CP8 tensor with the 8-channel vector and the x axis described as ONE inner dimension of 64-bit elements (two per
// pixel): dims (2W, H, N, C/8, 1); box (2 bw, bh, bn, kc, 1); x coordinates are doubled.  The TMA unit works row by row of the
// innermost dimension, so a box row is bw * 16 bytes instead of 16: a 40-pixel halo box is 24 requests instead of 960.
// `parts` == 2: split tensors (hi tensor, then lo tensor `chunks` planes further): the fifth dimension selects the part and
// one box brings both, [part][kc][pixels][16 B].
CUtensorMap make_cp8_wide_tensor_map(const void* base, int W, int H, int N, int chunks, long long plane, int bw, int bh, int bn, int kc,
                                     int parts) {
  CUtensorMap m;
  cuuint64_t dims[5] = {(cuuint64_t)2 * W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)chunks, (cuuint64_t)parts};
  cuuint64_t strides[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)plane * 16, (cuuint64_t)plane * 16 * chunks};
  cuuint32_t box[5] = {(cuuint32_t)2 * bw, (cuuint32_t)bh, (cuuint32_t)bn, (cuuint32_t)kc, (cuuint32_t)parts};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (2 * bw > 256) throw Error(FSR_E_INVALID, "wide CP8 box exceeds 256 elements");
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(FSR_E_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

// 5-D tensor map over an IM8 activation tensor [C/8][H*W][ncap][8], in 64-bit elements: dims (2 ncap, H*W, C/8, 1, 1);
// box (256, 1, kc, 1, 1) = 128 images at one pixel, kc channel chunks: every chunk is one contiguous 2 KB row.
CUtensorMap make_im8_tensor_map(const void* base, long long ncap, int HW, int chunks, int kc, int parts) {
  CUtensorMap m;
  cuuint64_t dims[5] = {(cuuint64_t)2 * ncap, (cuuint64_t)HW, (cuuint64_t)chunks, (cuuint64_t)parts, 1};
  cuuint64_t strides[4] = {(cuuint64_t)ncap * 16, (cuuint64_t)HW * ncap * 16, (cuuint64_t)HW * ncap * 16 * chunks,
                           (cuuint64_t)HW * ncap * 16 * chunks * parts};
  cuuint32_t box[5] = {256, 1, (cuuint32_t)kc, (cuuint32_t)parts, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(FSR_E_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

// 3-D tensor map over a 1-channel fp32 raster stack [N][H][W]: box (bw, 1, 1), zero fill outside.
CUtensorMap make_f32_tensor_map_3d(const void* base, int W, int H, int N, int bw) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  cuuint32_t box[3] = {(cuuint32_t)bw, 1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, N == 1 ? 2 : 3, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(FSR_E_CUDA, "cuTensorMapEncodeTiled (fp32 raster) failed with code " + std::to_string((int)r));
  return m;
}

void conv_tc_tile_box(int H, int W, int& bw, int& bh, int& bn) {
  // 128 output pixels per M tile: as many whole rows / images as needed
  bw = W < 128 ? W : 128;
  bh = (128 / bw) < H ? (128 / bw) : H;
  bn = 128 / (bw * bh);
  if (bw * bh * bn != 128 || W % bw || H % bh) throw Error(FSR_E_UNSUPPORTED, "feature-map size does not tile into 128-pixel boxes");
}

// Split mode: longest accumulation chain (hi*hi MMAs added into one TMEM accumulator) the configuration may produce.  The
// tensor core truncates on every addition into an accumulator: measured ~1e-7 relative error per step.  FSR_X3_CHAIN overrides.
static int split_chain_limit() {
  static const int v = getenv("FSR_X3_CHAIN") ? std::max(1, atoi(getenv("FSR_X3_CHAIN"))) : 48;
  return v;
}
// accumulators for the hi*hi products (1, 3 or 7: with the one for the small products the allocation is a power of two)
static int split_nacc(int hh_steps, int max_nacc) {
  const int lim = split_chain_limit();
  for (int n : {1, 3, 7})
    if (n <= max_nacc && (hh_steps + n - 1) / n <= lim) return n;
  return max_nacc;
}
// output channels per CTA.  Split mode (hh_steps = K / 16 > 0): 128-channel tiles leave room for 3 chain accumulators, 64-channel
// tiles for 7; the wider tile halves the activation traffic per MMA and is taken whenever its chains stay short enough
int conv_tc_bn(int cout, int parts, int hh_steps) {
  if (parts == 1) return cout >= 128 ? 128 : (cout >= 64 ? 64 : 32);
  if (cout >= 128 && (hh_steps + 2) / 3 <= split_chain_limit()) return 128;
  return cout >= 64 ? 64 : 32;
}

void launch_conv_tc(const __nv_bfloat16* src0, int C0, long long plane0, const __nv_bfloat16* src1, int C1, long long plane1,
                    const __nv_bfloat16* wpack, int kc, const float* bias, const __nv_bfloat16* res, __nv_bfloat16* dst,
                    long long plane_out, int n_img, int H, int W, int ksz, int cout, int act, float alpha, int half, int im,
                    int parts, float out_scale, int cpad_out, unsigned* flags, cudaStream_t s, int up0) {
  ConvTcParams p{};
  FSR_REQUIRE(!up0 || (im && ksz == 3 && src1 && H % 2 == 0 && W % 2 == 0), "folded upsampling needs an image-major 3 x 3 layer with a second source");
  p.up0 = up0;
  p.flags = flags;
  p.half = half;
  p.parts = parts;
  p.out_scale = parts == 2 ? out_scale : 1.0f;
  p.lo_off = parts == 2 ? (long long)(cpad_out / 8) * plane_out * 8 : 0;
  p.H = H; p.W = W; p.N = n_img;
  p.im = im;
  p.ncap = (int)(plane_out / ((long long)H * W));
  if (im) {
    p.bw = 1; p.bh = 1; p.bn = 128;
    p.tiles_x = W;
    p.tiles_y = H;
  } else {
    conv_tc_tile_box(H, W, p.bw, p.bh, p.bn);
    p.tiles_x = W / p.bw;
    p.tiles_y = H / p.bh;
  }
  p.ksz = ksz;
  p.kc = kc;
  p.s0 = (C0 / 8) / kc;
  p.s1 = src1 ? (C1 / 8) / kc : 0;
  p.cout = cout;
  p.act = act;
  p.alpha = alpha;
  p.plane = plane_out;
  p.wpack = wpack;
  p.bias = bias;
  p.res = res;
  p.out = dst;
  const int tiles_n = ceil_div(n_img, p.bn);
  CUtensorMap m0 = up0 ? make_im8_tensor_map(src0, plane0 / ((long long)(H / 2) * (W / 2)), (H / 2) * (W / 2), C0 / 8, kc, parts)
                   : im ? make_im8_tensor_map(src0, plane0 / ((long long)H * W), H * W, C0 / 8, kc, parts)
                      : make_cp8_wide_tensor_map(src0, W, H, n_img, C0 / 8, plane0, p.bw, p.bh, p.bn, kc, parts);
  CUtensorMap m1 = !src1 ? m0
                   : im  ? make_im8_tensor_map(src1, plane1 / ((long long)H * W), H * W, C1 / 8, kc, parts)
                         : make_cp8_wide_tensor_map(src1, W, H, n_img, C1 / 8, plane1, p.bw, p.bh, p.bn, kc, parts);
  const int n_iters = up0 ? 4 * p.s0 + 9 * p.s1 : ksz * ksz * (p.s0 + p.s1);
  // split mode: 1, 3 or 7 accumulators for the hi*hi products (+ 1 for the small ones; a power-of-two TMEM allocation), so
  // that one accumulation chain stays short: the tensor core truncates on every addition into an accumulator
  const int hh_steps = n_iters * (kc / 2);
  const int BN = conv_tc_bn(cout, parts, hh_steps);
  p.nacc = parts == 2 ? split_nacc(hh_steps, 512 / BN - 1) : 1;
  // Two image blocks per CTA (every weight slice fetched from L2 feeds two MMAs) where it measured faster on B200: maps of
  // <= 16 pixels with enough CTAs left to fill the SMs.  Larger maps have short K loops and many CTAs; there two
  // co-resident single-tile CTAs (8 epilogue warps per SM instead of 4) win.  env FSR_NO_CONV_PAIR disables it.
  const long long n_single = (long long)p.tiles_x * p.tiles_y * tiles_n * ceil_div(cout, BN);
  // (split mode: no H1-family layer has both a <= 16-pixel map and a tile narrow enough for two M tiles' accumulators in TMEM)
  p.pair = (parts == 1 && tiles_n >= 2 && n_single >= 256 && p.tiles_x * p.tiles_y <= 16 && !getenv("FSR_NO_CONV_PAIR")) ? kConvPairs : 1;
  dim3 grid((unsigned)(p.tiles_x * p.tiles_y * ceil_div(tiles_n, p.pair)), (unsigned)ceil_div(cout, BN));
  if (BN == 128) {
    static bool attr[64] = {false};
    if (first_on_device(attr)) {
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    }
    p.stages = conv_stages<128>(kc * parts, n_iters, (long long)grid.x * grid.y, p.pair, parts == 2 && p.pair * 128 * (p.nacc + 1) > 256);
    if (parts == 2) launch_pdl(conv_tc_kernel<128, 1, 2>, grid, kConvThreads<1>, conv_smem_bytes<128>(kc * parts, p.stages, 1), s, m0, m1, p);
    else if (p.pair == 2) launch_pdl(conv_tc_kernel<128, 2, 1>, grid, kConvThreads<2>, conv_smem_bytes<128>(kc * parts, p.stages, 2), s, m0, m1, p);
    else launch_pdl(conv_tc_kernel<128, 1, 1>, grid, kConvThreads<1>, conv_smem_bytes<128>(kc * parts, p.stages, 1), s, m0, m1, p);
  } else if (BN == 64) {
    static bool attr[64] = {false};
    if (first_on_device(attr)) {
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    }
    p.stages = conv_stages<64>(kc * parts, n_iters, (long long)grid.x * grid.y, p.pair, parts == 2 && p.pair * 64 * (p.nacc + 1) > 256);
    if (parts == 2) launch_pdl(conv_tc_kernel<64, 1, 2>, grid, kConvThreads<1>, conv_smem_bytes<64>(kc * parts, p.stages, 1), s, m0, m1, p);
    else if (p.pair == 2) launch_pdl(conv_tc_kernel<64, 2, 1>, grid, kConvThreads<2>, conv_smem_bytes<64>(kc * parts, p.stages, 2), s, m0, m1, p);
    else launch_pdl(conv_tc_kernel<64, 1, 1>, grid, kConvThreads<1>, conv_smem_bytes<64>(kc * parts, p.stages, 1), s, m0, m1, p);
  } else {
    static bool attr[64] = {false};
    if (first_on_device(attr)) {
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
      FSR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<32, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024));
    }
    p.stages = conv_stages<32>(kc * parts, n_iters, (long long)grid.x * grid.y, p.pair, parts == 2 && p.pair * 32 * (p.nacc + 1) > 256);
    if (parts == 2) launch_pdl(conv_tc_kernel<32, 1, 2>, grid, kConvThreads<1>, conv_smem_bytes<32>(kc * parts, p.stages, 1), s, m0, m1, p);
    else if (p.pair == 2) launch_pdl(conv_tc_kernel<32, 2, 1>, grid, kConvThreads<2>, conv_smem_bytes<32>(kc * parts, p.stages, 2), s, m0, m1, p);
    else launch_pdl(conv_tc_kernel<32, 1, 1>, grid, kConvThreads<1>, conv_smem_bytes<32>(kc * parts, p.stages, 1), s, m0, m1, p);
  }
  FSR_LAUNCH_CHECK();
}


// Row-box variant: 32- or 16-pixel-wide maps whose padded enumeration tiles into whole 128-position M tiles and whose packed
// weights fit in shared memory.
static void conv_rows_geometry(int H, int W, int& pitch, int& box_rows, int& tiles_per_img) {
  pitch = W + 8;
  tiles_per_img = H * pitch / 128;
  int span = 1;  // image rows one M tile can touch
  for (int t = 0; t < tiles_per_img; ++t) span = std::max(span, (t * 128 + 127) / pitch - (t * 128) / pitch + 1);
  box_rows = span + 2;
}

static size_t conv_rows_stage_bytes(int kc, int box_rows, int pitch) {  // kc = planes per box, all parts counted
  return ((size_t)kc * box_rows * pitch * 16 + 2 * pitch * 16 + 127) & ~(size_t)127;
}

bool conv_rows_ok(int H, int W, int ksz, int cout, int C0, int C1, int kc, int parts) {
  if (ksz != 3 || (W != 16 && W != 32) || (cout != 32 && cout != 64)) return false;
  if ((H * (W + 8)) % 128) return false;
  if (kc != 2 && kc != 4) return false;
  int pitch, box_rows, tiles_per_img;
  conv_rows_geometry(H, W, pitch, box_rows, tiles_per_img);
  const size_t w_bytes = (size_t)parts * 9 * ((C0 + C1) / 8) * cout * 16;
  // The two MMA-issuing warps work on consecutive tiles at once and wait on ring positions up to one tile ahead of what the
  // other has consumed: a parity wait is only meaningful within one lap of the ring, so both tiles' boxes must fit in it
  // (a 16-channel-group variant for the wider split-mode layers violated this and read boxes before they had landed).
  const size_t groups = (size_t)((C0 + C1) / 8) / kc;
  const size_t need = std::max<size_t>(3, kRowsIssuers * groups);
  const size_t smem = ((w_bytes + 1023) & ~(size_t)1023) + need * conv_rows_stage_bytes(kc * parts, box_rows, pitch) + 512;
  return need <= (size_t)kRowMaxStages && smem <= 200 * 1024;
}

void launch_conv_rows_tc(const __nv_bfloat16* src0, int C0, long long plane0, const __nv_bfloat16* src1, int C1, long long plane1,
                         const __nv_bfloat16* wpack, int kc, const float* bias, const __nv_bfloat16* res, __nv_bfloat16* dst,
                         long long plane_out, int n_img, int H, int W, int cout, int act, float alpha, int half, int n_sms,
                         int parts, float out_scale, unsigned* flags, cudaStream_t s) {
  ConvRowsParams p{};
  p.flags = flags;
  p.parts = parts;
  // split mode: the same chain policy as conv_tc_kernel (the fewest of 1 / 3 / 7 accumulators that keeps every accumulation
  // chain <= 48 MMAs: these layers have 9 - 18 K steps per output, one chain), then as many accumulator buffers / epilogue
  // sets as TMEM holds: sets x (nacc + 1) x cout columns, a power of two.  (Measured: 2 buffers x 8 accumulators 3.33 ms for
  // the LR convolutions of a step, 4 x 4 3.22 ms, 4 x 2 3.16 ms.)  FSR_ROWS_CFG="nacc,sets" overrides (A/B).
  p.sets = kRowsEpiSets;
  p.nacc = 1;
  if (parts == 2) {
    const int hh_steps = 9 * ((C0 + C1) / 16);
    p.nacc = split_nacc(hh_steps, 512 / cout - 1);
    while (p.sets > 1 && p.sets * (p.nacc + 1) * cout > 512) p.sets /= 2;
    if (const char* e = getenv("FSR_ROWS_CFG")) {
      int a = 0, b = 0;
      if (sscanf(e, "%d,%d", &a, &b) == 2 && (a == 1 || a == 3 || a == 7) && (b == 1 || b == 2 || b == 4) && b * (a + 1) * cout <= 512) {
        p.nacc = a;
        p.sets = b;
      }
    }
  }
  p.out_scale = parts == 2 ? out_scale : 1.0f;
  p.lo_off = parts == 2 ? (long long)(cout / 8) * plane_out * 8 : 0;
  p.H = H; p.W = W; p.N = n_img;
  conv_rows_geometry(H, W, p.pitch, p.box_rows, p.tiles_per_img);
  p.pitch_magic = (unsigned)((0x100000000ull + p.pitch - 1) / p.pitch);
  static unsigned long long* d_stats = nullptr;
  if (getenv("FSR_ROWS_STATS") && !d_stats) FSR_CUDA(cudaMalloc(&d_stats, 148 * 16 * sizeof(unsigned long long)));
  p.stats = d_stats;
  p.n_mtiles = n_img * p.tiles_per_img;
  p.kc = kc;
  p.g0 = (C0 / 8) / kc;
  p.g1 = src1 ? (C1 / 8) / kc : 0;
  p.cout = cout; p.act = act; p.alpha = alpha; p.half = half;
  p.plane = plane_out;
  p.wpack = wpack;
  p.w_bytes = parts * 9 * (p.g0 + p.g1) * kc * cout * 16;
  p.bias = bias; p.res = res; p.out = dst;
  CUtensorMap m0 = make_cp8_wide_tensor_map(src0, W, H, n_img, C0 / 8, plane0, p.pitch, p.box_rows, 1, kc, parts);
  CUtensorMap m1 = src1 ? make_cp8_wide_tensor_map(src1, W, H, n_img, C1 / 8, plane1, p.pitch, p.box_rows, 1, kc, parts) : m0;
  const size_t w_round = ((size_t)(p.w_bytes + 1023) & ~(size_t)1023), stage_bytes = conv_rows_stage_bytes(kc * parts, p.box_rows, p.pitch);
  p.stages = (int)std::min<size_t>(kRowMaxStages, (200 * 1024 - 512 - w_round) / stage_bytes);
  const size_t smem = w_round + (size_t)p.stages * stage_bytes + 512;  // barriers (256 B) + bias (256 B)
  const int grid = p.n_mtiles < n_sms ? p.n_mtiles : n_sms;
  auto go = [&](auto kernel) {
    FSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    launch_pdl(kernel, dim3((unsigned)grid), kRowsThreads, smem, s, m0, m1, p);
  };
  if (cout == 64) {
    if (parts == 2) go(conv_rows_tc_kernel<64, 2>);
    else go(conv_rows_tc_kernel<64, 1>);
  } else {
    if (parts == 2) go(conv_rows_tc_kernel<32, 2>);
    else go(conv_rows_tc_kernel<32, 1>);
  }
  FSR_LAUNCH_CHECK();
  if (d_stats) {
    std::vector<unsigned long long> h(148 * 16);
    FSR_CUDA(cudaStreamSynchronize(s));
    FSR_CUDA(cudaMemcpy(h.data(), d_stats, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const unsigned long long* c = h.data() + 16 * (grid / 2);
    fprintf(stderr, "[rows W=%d cin=%d cout=%d tiles=%d] producer total %llu wait-empty %llu | issuer0 total %llu wait-acc %llu wait-full %llu tiles %llu | issuer1 total %llu wait-acc %llu wait-full %llu | epi total %llu wait-acc-full %llu wait+ld %llu\n",
            W, C0 + C1, cout, p.n_mtiles, c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[10], c[11], c[12]);
  }
}

void launch_pack_small(const float* s0, int c0, const float* s1, int c1, __nv_bfloat16* dst, long long n_pix, long long plane,
                       int chunks, int half, int parts, cudaStream_t s) {
  pack_small_kernel<<<grid_for(n_pix * chunks), 256, 0, s>>>(s0, c0, s1, c1, dst, n_pix, plane, chunks, half, parts);
  FSR_LAUNCH_CHECK();
}

void launch_pool_cp8(const __nv_bfloat16* src, __nv_bfloat16* dst, int chunks, int n_img, int Hin, int Win, int k, int mode,
                     long long plane_in, long long plane_out, int half, int im_in, int im_out, int parts, int aux, cudaStream_t s) {
  const long long ncap_in = plane_in / ((long long)Hin * Win), ncap_out = plane_out / ((long long)(Hin / k) * (Win / k));
  pool_cp8_kernel<<<grid_for((long long)n_img * (Hin / k) * (Win / k) * chunks), 256, 0, s>>>(src, dst, chunks, n_img, Hin, Win, k, mode,
                                                                                                ncap_in, ncap_out, half, im_in, im_out, parts, aux);
  FSR_LAUNCH_CHECK();
}

void launch_upsample_cp8(const __nv_bfloat16* src, __nv_bfloat16* dst, int chunks, int n_img, int Hin, int Win, int f,
                         long long plane_in, long long plane_out, int im_in, int im_out, int mode, int half, int parts, cudaStream_t s) {
  const long long ncap_in = plane_in / ((long long)Hin * Win), ncap_out = plane_out / ((long long)Hin * f * Win * f);
  upsample_cp8_kernel<<<grid_for((long long)n_img * Hin * f * Win * f * chunks), 256, 0, s>>>(src, dst, chunks, n_img, Hin, Win, f,
                                                                                                ncap_in, ncap_out, im_in, im_out, mode, half, parts);
  FSR_LAUNCH_CHECK();
}

void launch_eltwise_cp8(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* dst, long long n_vec, int act, float alpha,
                        float beta, int half, long long lo_vec, cudaStream_t s) {
  eltwise_cp8_kernel<<<grid_for(n_vec), 256, 0, s>>>(a, b, dst, n_vec, act, alpha, beta, half, lo_vec);
  FSR_LAUNCH_CHECK();
}

void launch_cp8_to_nhwc(const __nv_bfloat16* src, float* dst, long long n_img, int H, int W, long long plane, int C, int half, int im,
                        long long lo_off, cudaStream_t s) {
  cp8_to_nhwc_kernel<<<grid_for(n_img * H * W * C), 256, 0, s>>>(src, dst, n_img, H, W, plane / ((long long)H * W), C, half, im, lo_off);
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
