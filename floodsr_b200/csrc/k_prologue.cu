// Per-tile input normalisation on the GPU (memory-bound stage).
//
// Replaces, for every window of the tile loop, the numpy calls of EngineORT.run_tile
// (floodsr/engine/ort.py:151-159): replace_nodata_with_zero (preprocessing.py:167-172), the finite asserts,
// scale_depth_log1p_np (:141-151) and normalize_dem (:97-126) with its tile-local np.nanpercentile.
//
// One CTA per tile.  The tile (hr x hr float32, 1 MiB at hr=512) is streamed from HBM once and then
// re-read from L2 by the later passes:
//   pass A   nodata->0, finite check, clip at 0, min/max (warp-shuffle reductions)
//   pass B.. exact selection of the two order statistics numpy's 'linear' percentile interpolates between:
//            radix select on the float bit patterns with a range-adaptive 11-bit digit (<= 3 passes)
//   pass N   min-max normalisation, written as the network's dem_hr input
// The window is read straight out of the raster with bounds checks, which implements the reference's
// zero padding to whole tiles (ResUNet_16x_DEM.py:215-235) without materialising a padded copy.
#include "fsr_common.cuh"

namespace fsr {

namespace {

constexpr int kThreads = 512;      // two CTAs per SM: one tile's barrier-heavy selection overlaps the other's streaming
constexpr int kBins = 2048;
constexpr int kDigitBits = 11;
constexpr int kBinsPerThread = kBins / kThreads;
constexpr int kWarps = kThreads / 32;
constexpr int kCap = 22528;        // compacted upper-tail keys kept in shared memory (88 KB)
constexpr int kSampleRows = 16;    // every 16th tile row feeds the threshold estimate
constexpr int kUnroll = 4;         // independent loads in flight per thread in the streaming passes
// three-kernel path: per-tile words in the workspace ahead of the compacted keys
constexpr int kSlices = 4;         // CTAs that scan one tile (row slices)
constexpr int kMetaWords = 8;
constexpr int kMetaCount = 0, kMetaMin = 1, kMetaMax = 2, kMetaThr = 3, kMetaStatus = 4;
constexpr unsigned kStatusDone = 1u;

struct SelState {
  unsigned lo, hi;  // inclusive key range still containing the wanted order statistic
  int rank;         // rank of the wanted element among keys in [lo, hi]
};

__device__ __forceinline__ float fix_nodata(float x, const fsr_tile_params& p) {
  if (p.has_dem_nodata) {
    bool hit = (x == p.dem_nodata) || (p.dem_nodata_tol >= 0.0f && fabsf(x - p.dem_nodata) <= p.dem_nodata_tol);
    if (hit) x = 0.0f;
  }
  return x;
}

__device__ __forceinline__ float fix_dem(float x, const fsr_tile_params& p, unsigned& flags) {
  x = fix_nodata(x, p);
  if (!isfinite(x)) {
    flags |= FSR_FLAG_DEM_NONFINITE;
    x = 0.0f;
  }
  return x;
}

// Loads 4 consecutive pixels of tile row r starting at column c (c % 4 == 0), zero beyond the raster.
__device__ __forceinline__ float4 load4(const float* __restrict__ dem, int H, int W, int gy, int gx) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (gy >= H || gx >= W) return v;
  size_t off = (size_t)gy * (size_t)W + (size_t)gx;
  const float* ptr = dem + off;
  if (gx + 3 < W && ((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0)) {
    v = __ldg(reinterpret_cast<const float4*>(ptr));
  } else {
    v.x = __ldg(ptr);
    if (gx + 1 < W) v.y = __ldg(ptr + 1);
    if (gx + 2 < W) v.z = __ldg(ptr + 2);
    if (gx + 3 < W) v.w = __ldg(ptr + 3);
  }
  return v;
}

// key of a DEM value: bit pattern of clip(fix(x), 0, inf) (+0.0f canonicalises -0.0), monotone in the value
__device__ __forceinline__ unsigned dem_key(float x, const fsr_tile_params& p, unsigned& flags) {
  return __float_as_uint(fmaxf(fix_dem(x, p, flags), 0.0f) + 0.0f);
}

// Exclusive prefix over the histogram; publishes (through smem) the bin whose cumulative count first exceeds `rank`
// and the count before it.  All threads call this.
__device__ void find_bin(const int* __restrict__ hist, int rank, int* warp_tot, int* sel_bin, int* sel_before) {
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  int h[kBinsPerThread];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < kBinsPerThread; ++k) {
    h[k] = hist[kBinsPerThread * t + k];
    sum += h[k];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kWarps ? warp_tot[lane] : 0;
    int s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    if (lane < kWarps) warp_tot[lane] = s - w;  // exclusive prefix of warp totals
  }
  __syncthreads();
  int excl = warp_tot[warp] + incl - sum;
#pragma unroll
  for (int k = 0; k < kBinsPerThread; ++k) {
    if (rank >= excl && rank < excl + h[k]) {
      *sel_bin = kBinsPerThread * t + k;
      *sel_before = excl;
    }
    excl += h[k];
  }
  __syncthreads();
}

__device__ __forceinline__ int digit_shift(unsigned lo, unsigned hi) {
  unsigned span = hi - lo;
  int bits = 32 - __clz(span);  // span > 0 here
  return bits > kDigitBits ? bits - kDigitBits : 0;
}

// One radix-select refinement step for both wanted order statistics, after the histograms have been filled.
__device__ void refine(SelState* st, const SelState& s0, const SelState& s1, bool need0, bool need1, bool same, int sh0, int sh1,
                       int (*hist)[kBins], int* warp_tot, int* sel_bin, int* sel_before) {
  const int t = threadIdx.x;
  if (need0) {
    find_bin(hist[0], s0.rank, warp_tot, sel_bin, sel_before);
    if (t == 0) {
      unsigned nlo = s0.lo + ((unsigned)*sel_bin << sh0);
      unsigned width = (sh0 >= 32) ? 0xffffffffu : ((1u << sh0) - 1u);
      unsigned nhi = nlo + width;
      if (nhi > s0.hi || nhi < nlo) nhi = s0.hi;
      st[0] = SelState{nlo, nhi, s0.rank - *sel_before};
    }
    __syncthreads();
  }
  if (need1) {
    const int* h = same ? hist[0] : hist[1];
    const int sh = same ? sh0 : sh1;
    find_bin(h, s1.rank, warp_tot, sel_bin, sel_before);
    if (t == 0) {
      unsigned nlo = s1.lo + ((unsigned)*sel_bin << sh);
      unsigned width = (1u << sh) - 1u;
      unsigned nhi = nlo + width;
      if (nhi > s1.hi || nhi < nlo) nhi = s1.hi;
      st[1] = SelState{nlo, nhi, s1.rank - *sel_before};
    }
    __syncthreads();
  }
}

// Adds one key to the selection histograms.  Keys equal to the range start are counted by ballot (zero padding would
// otherwise serialise hundreds of thousands of same-address shared atomics); must be called by whole warps.
__device__ __forceinline__ void hist_add(unsigned key, bool valid, const SelState& s0, const SelState& s1, bool need0, bool need1,
                                         bool same, int sh0, int sh1, int (*hist)[kBins], int& lo0_hits) {
  if (need0) {
    const bool at_lo = valid && (key == s0.lo);
    lo0_hits += __popc(__ballot_sync(0xffffffffu, at_lo));
    if (valid && !at_lo && key > s0.lo && key <= s0.hi) atomicAdd(&hist[0][(key - s0.lo) >> sh0], 1);
  }
  if (need1 && !same) {
    if (valid && key >= s1.lo && key <= s1.hi) atomicAdd(&hist[1][(key - s1.lo) >> sh1], 1);
  }
}

__global__ void __launch_bounds__(kThreads, 2)
tile_normalize_kernel(const float* __restrict__ dem, const float* __restrict__ depth, TileGrid grid, int tile_base,
                      int T, int TL, int scale, fsr_tile_params p, float* __restrict__ dem_norm,
                      float* __restrict__ depth_norm, float* __restrict__ stats, float* __restrict__ dem_lr,
                      unsigned* __restrict__ flags_out, const unsigned* __restrict__ ws_meta) {
  extern __shared__ unsigned s_keys[];   // [kCap] compacted upper-tail keys
  __shared__ int hist[2][kBins];
  __shared__ int warp_tot[32];
  __shared__ float red_min[32], red_max[32];
  __shared__ float pool_part[4][128];
  __shared__ SelState st[2];
  __shared__ int sel_bin, sel_before;
  __shared__ unsigned s_flags;
  __shared__ int s_count;

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int tile_local = blockIdx.x;
  // second launch after the three-kernel path (launch_tile_normalize): only the tiles that path gave up on are left
  if (ws_meta && ws_meta[(size_t)tile_local * kMetaWords + kMetaStatus] == kStatusDone) return;
  const int2 org = grid.origins[tile_base + tile_local];
  const int vec_per_row = T / 4;
  const int n_vec = T * vec_per_row;
  // tile-uniform fast addressing: power-of-two rows split with a shift, and windows that lie wholly inside the raster on
  // 16-byte boundaries skip the per-vector bounds and alignment tests
  const int vsh = (vec_per_row & (vec_per_row - 1)) == 0 ? __ffs(vec_per_row) - 1 : -1;
  const bool interior = org.x + T <= grid.H && org.y + T <= grid.W && (grid.W & 3) == 0 && (org.y & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(dem) & 15u) == 0;
  const float* tile0 = dem + (size_t)org.x * (size_t)grid.W + (size_t)org.y;
  auto tile_load4 = [&](int i) {
    int r, c;
    if (vsh >= 0) {
      r = i >> vsh;
      c = (i & (vec_per_row - 1)) * 4;
    } else {
      r = i / vec_per_row;
      c = (i - r * vec_per_row) * 4;
    }
    if (interior) return __ldg(reinterpret_cast<const float4*>(tile0 + (size_t)r * (size_t)grid.W + c));
    return load4(dem, grid.H, grid.W, org.x + r, org.y + c);
  };
  const int n_px = T * T;
  unsigned my_flags = 0;
  if (t == 0) {
    s_flags = 0;
    s_count = 0;
  }

  // ---- depth_lr: nodata -> 0, finite check, log1p scaling (preprocessing.py:141-151) ----------------
  for (int i = t; i < TL * TL; i += kThreads) {
    int r = i / TL, c = i - r * TL;
    int ly = org.x / scale + r, lx = org.y / scale + c;
    float x = (ly < grid.Hl && lx < grid.Wl) ? __ldg(depth + (size_t)ly * grid.Wl + lx) : 0.0f;
    if (p.has_depth_nodata) {
      bool hit = (x == p.depth_nodata) || (p.depth_nodata_tol >= 0.0f && fabsf(x - p.depth_nodata) <= p.depth_nodata_tol);
      if (hit) x = 0.0f;
    }
    if (!isfinite(x)) {
      my_flags |= FSR_FLAG_DEPTH_NONFINITE;
      x = 0.0f;
    }
    float d = fminf(fmaxf(x, 0.0f), p.max_depth);
    float s = __fdiv_rn(log1pf(d), p.depth_denom);
    depth_norm[(size_t)tile_local * TL * TL + i] = fminf(fmaxf(s, 0.0f), 1.0f);
  }

  const bool own_stats = !p.has_ref_stats;
  // Upper-tail fast path (default dem_pct_clip = 95): the two order statistics lie among the n_top largest keys.  A
  // threshold estimated from a 1/8 sample lets ONE full pass compact every key >= threshold into shared memory
  // (exact count checked afterwards), where the exact radix select runs without touching the tile again.
  const int n_top = n_px - p.rank_lo;
  const bool try_fast = own_stats && (long long)n_top * 14 / 10 + 2048 <= kCap && T % kSampleRows == 0 &&
                        (T / kSampleRows) * T <= kCap && ((T / kSampleRows) * vec_per_row) % kThreads == 0;
  unsigned thr_key = 0;
  if (try_fast) {
    // one coalesced read of every kSampleRows-th tile row (6 % of the tile) into shared memory; min / max and the
    // histogram of the sample are then formed from there
    const int n_srows = T / kSampleRows, n_samp = n_srows * T;
    float smin = INFINITY, smax = 0.0f;
    unsigned dummy = 0;
    for (int i = t; i < n_srows * vec_per_row; i += kThreads) {
      const int sr = i / vec_per_row, c = (i - sr * vec_per_row) * 4;
      const float4 v = load4(dem, grid.H, grid.W, org.x + sr * kSampleRows + kSampleRows / 2, org.y + c);
      const float e[4] = {v.x, v.y, v.z, v.w};
      unsigned k[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float f = fmaxf(fix_dem(e[j], p, dummy), 0.0f) + 0.0f;
        smin = fminf(smin, f);
        smax = fmaxf(smax, f);
        k[j] = __float_as_uint(f);
      }
      *reinterpret_cast<uint4*>(&s_keys[4 * i]) = make_uint4(k[0], k[1], k[2], k[3]);
    }
    smin = warp_min(smin);
    smax = warp_max(smax);
    if (lane == 0) {
      red_min[warp] = smin;
      red_max[warp] = smax;
    }
    for (int i = t; i < kBins; i += kThreads) hist[0][i] = 0;
    __syncthreads();
    smin = red_min[0];
    smax = red_max[0];
    for (int w = 1; w < kWarps; ++w) {
      smin = fminf(smin, red_min[w]);
      smax = fmaxf(smax, red_max[w]);
    }
    const unsigned klo = __float_as_uint(smin), khi = __float_as_uint(smax);
    if (khi > klo) {
      const int sh = digit_shift(klo, khi);
      for (int i = t; i < n_samp; i += kThreads) atomicAdd(&hist[0][(s_keys[i] - klo) >> sh], 1);
      __syncthreads();
      // rank (from below) of the sample element above which ~1.3 x n_top / kSampleRows samples lie
      const int want_above = (int)(((long long)n_top * 13 / 10 + 512) / kSampleRows);
      const int srank = n_samp - 1 - want_above;
      if (srank > 0) {
        find_bin(hist[0], srank, warp_tot, &sel_bin, &sel_before);
        thr_key = klo + ((unsigned)sel_bin << sh);  // lower edge of that sample bin
      }
    }
    __syncthreads();  // hist / red arrays and the key buffer are reused below
  }

  // ---- pass A: min / max of clip(x, 0, inf), finite checks, compaction of the upper tail ------------
  float vmin = INFINITY, vmax = 0.0f;
  const bool compact = try_fast && thr_key > 0;
  // kUnroll independent 16-byte loads per thread are issued before any is consumed: the pass is HBM-latency bound
  for (int i0 = t; i0 < n_vec; i0 += kThreads * kUnroll) {
    float4 vv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      vv[u] = tile_load4(i0 + u * kThreads);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
    const float4 v = vv[u];
    float e[4];
    e[0] = fmaxf(fix_dem(v.x, p, my_flags), 0.0f);
    e[1] = fmaxf(fix_dem(v.y, p, my_flags), 0.0f);
    e[2] = fmaxf(fix_dem(v.z, p, my_flags), 0.0f);
    e[3] = fmaxf(fix_dem(v.w, p, my_flags), 0.0f);
    vmin = fminf(vmin, fminf(fminf(e[0], e[1]), fminf(e[2], e[3])));
    vmax = fmaxf(vmax, fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])));
    if (compact) {
      unsigned k[4];
      int mine = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        k[j] = __float_as_uint(e[j] + 0.0f);
        mine += k[j] >= thr_key ? 1 : 0;
      }
      // warp-aggregated append: one shared atomic per warp per iteration
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int base = 0;
      if (total > 0) {
        if (lane == 31) base = atomicAdd(&s_count, total);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + incl - mine;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (k[j] >= thr_key) {
            if (pos < kCap) s_keys[pos] = k[j];
            ++pos;
          }
      }
    }
  }
  }
  vmin = warp_min(vmin);
  vmax = warp_max(vmax);
  if (lane == 0) {
    red_min[warp] = vmin;
    red_max[warp] = vmax;
  }
  if (my_flags) atomicOr(&s_flags, my_flags);
  __syncthreads();
  if (warp == 0) {
    float a = warp_min(lane < kWarps ? red_min[lane] : INFINITY), b = warp_max(lane < kWarps ? red_max[lane] : 0.0f);
    if (lane == 0) {
      red_min[0] = a;
      red_max[0] = b;
      st[0] = SelState{__float_as_uint(a + 0.0f), __float_as_uint(b + 0.0f), p.rank_lo};
      st[1] = SelState{__float_as_uint(a + 0.0f), __float_as_uint(b + 0.0f), p.rank_hi};
    }
  }
  __syncthreads();
  vmin = red_min[0];
  vmax = red_max[0];
  // the compacted set is usable when it holds every key of rank >= rank_lo and did not overflow
  const int n_comp = s_count;
  const bool fast = compact && n_comp <= kCap && n_comp >= n_top;
  if (fast && t == 0) {
    const int below = n_px - n_comp;  // keys smaller than the threshold
    st[0] = SelState{thr_key, __float_as_uint(vmax + 0.0f), p.rank_lo - below};
    st[1] = SelState{thr_key, __float_as_uint(vmax + 0.0f), p.rank_hi - below};
  }
  __syncthreads();

  // ---- passes B..: radix select of sorted[rank_lo] and sorted[rank_hi] ------------------------------
  // range-adaptive 11-bit digits; over the compacted keys in shared memory (fast path) or over the tile in L2
  for (int guard = 0; own_stats && guard < 8; ++guard) {
    const SelState s0 = st[0], s1 = st[1];
    const bool need0 = s0.hi > s0.lo, need1 = s1.hi > s1.lo;
    if (!need0 && !need1) break;
    const bool same = (s0.lo == s1.lo) && (s0.hi == s1.hi);
    const int sh0 = need0 ? digit_shift(s0.lo, s0.hi) : 0;
    const int sh1 = need1 ? digit_shift(s1.lo, s1.hi) : 0;
    for (int i = t; i < 2 * kBins; i += kThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    int lo0_hits = 0;
    if (fast) {
      const int n_round = (n_comp + 31) & ~31;
      for (int i = t; i < n_round; i += kThreads) {
        const bool valid = i < n_comp;
        hist_add(valid ? s_keys[i] : 0u, valid, s0, s1, need0, need1, same, sh0, sh1, hist, lo0_hits);
      }
    } else {
      for (int i = t; i < n_vec; i += kThreads) {
        float4 v = tile_load4(i);
        unsigned dummy = 0;
        float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) hist_add(dem_key(e[j], p, dummy), true, s0, s1, need0, need1, same, sh0, sh1, hist, lo0_hits);
      }
    }
    if (need0 && lane == 0 && lo0_hits) atomicAdd(&hist[0][0], lo0_hits);
    __syncthreads();
    refine(st, s0, s1, need0, need1, same, sh0, sh1, hist, warp_tot, &sel_bin, &sel_before);
  }

  // ---- numpy _lerp on the two order statistics (float32, no FMA contraction) -------------------------
  const float a = __uint_as_float(st[0].lo), b = __uint_as_float(st[1].lo);
  const float diff = __fsub_rn(b, a);
  float p_clip = __fadd_rn(a, __fmul_rn(diff, p.gamma));
  if (p.gamma >= 0.5f) p_clip = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, p.gamma)));
  float dem_min = fminf(vmin, p_clip);
  float dem_max = fminf(vmax, p_clip);
  if (!own_stats) {  // caller-supplied stats (preprocessing.py:122-123); pass A only did the finite check
    p_clip = p.ref_p_clip;
    dem_min = p.ref_dem_min;
    dem_max = p.ref_dem_max;
  }
  const double range_d = (double)dem_max - (double)dem_min;  // python-float subtraction in the reference
  bool zero_out = false;
  if (!(range_d > 0.0)) {
    zero_out = true;  // preprocessing.py:71-82: all-zero tile -> zeros, any other flat tile is an error
    if (!(fabs(range_d) <= 1e-8 && fabs((double)dem_min) <= 1e-8) && t == 0) atomicOr(&s_flags, FSR_FLAG_DEM_FLAT_NONZERO);
  }
  const float range_f = (float)range_d;
  if (t == 0) {
    float* s = stats + (size_t)(tile_base + tile_local) * 3;
    s[0] = p_clip;
    s[1] = dem_min;
    s[2] = dem_max;
  }

  // ---- pass N: clip to [0, p_clip], min-max scale, clip to [0, 1] (preprocessing.py:91-94) -----------
  // plus, when dem_lr != nullptr, the scale x scale average pooling of the normalised tile that feeds the network's
  // low-resolution branch (the graph's AveragePool on dem_hr): fixed summation order, no atomics.
  // dem_norm == nullptr: the consumer (the fused high-resolution kernel) normalises the raster window itself from `stats`
  // with this very sequence of operations, so the 1 MiB per tile is neither written here nor read back there
  float4* out = dem_norm ? reinterpret_cast<float4*>(dem_norm + (size_t)tile_local * T * T) : nullptr;
  const bool pool = dem_lr != nullptr && scale == 16 && T == 512 && kUnroll % 4 == 0;  // thread <-> (row offset t / 128, column vector t % 128)
  for (int i0 = t; i0 < n_vec; i0 += kThreads * kUnroll) {
    float4 vv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      vv[u] = tile_load4(i0 + u * kThreads);
    }
    float psum[kUnroll / 4];
#pragma unroll
    for (int h = 0; h < kUnroll / 4; ++h) psum[h] = 0.0f;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float e[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // non-finite values were flagged in pass A (the call fails); fmaxf / fminf map NaN to 0 and +-Inf into the clip range
        float x = fminf(fmaxf(fix_nodata(e[j], p), 0.0f), p_clip);
        float n = __fdiv_rn(__fsub_rn(x, dem_min), range_f);
        e[j] = zero_out ? 0.0f : fminf(fmaxf(n, 0.0f), 1.0f);
      }
      if (out) out[i0 + u * kThreads] = make_float4(e[0], e[1], e[2], e[3]);
      psum[u / 4] += (e[0] + e[1]) + (e[2] + e[3]);
    }
    if (pool) {
      // 4 unrolled steps x 4 tile rows = 16 rows = one row of pooled cells
#pragma unroll
      for (int h = 0; h < kUnroll / 4; ++h) {
        float q = psum[h] + __shfl_xor_sync(0xffffffffu, psum[h], 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        if ((lane & 3) == 0) pool_part[t >> 7][(t & 127) >> 2] = q;   // [row offset][cell]
        __syncthreads();
        if (t < 32) {
          const float sum = (pool_part[0][t] + pool_part[1][t]) + (pool_part[2][t] + pool_part[3][t]);
          dem_lr[(size_t)tile_local * (TL * TL) + ((i0 / (kThreads * kUnroll)) * (kUnroll / 4) + h) * TL + t] = sum * (1.0f / 256.0f);
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  if (t == 0 && s_flags) atomicOr(flags_out, s_flags);
}

// ---- three-kernel path -------------------------------------------------------------------------------------------------
// One CTA per tile spends most of a tile's ~170 us waiting on its own barriers and load latencies, which is what a batch of
// <= 148 windows (one window row of the pipelines) pays in full, and 935 tiles pay as 4 rounds of 296 resident CTAs.  The same
// arithmetic in three grids that fill the machine at any batch size:
//   scan    4 CTAs per tile (row slices): every slice derives the tile's threshold from the 1/16 row sample (same rows, same
//           arithmetic -> same threshold), then streams its rows once: finite checks, min / max (global atomics on the
//           monotone bit patterns) and the upper-tail keys appended to the tile's key list in global memory (L2);
//   select  1 CTA per tile: exact radix select of the two order statistics over that list -> stats; a tile whose list
//           overflowed or missed the tail is left to the one-CTA kernel (status stays 0);
//   apply   1 CTA per 16 tile rows: normalisation + 16 x 16 pooling with the one-CTA kernel's summation order (bit-identical).
// The order of the keys in the list depends on timing; the selected order statistics do not (the selection is exact).
struct TileReader {
  const float* dem;
  const float* tile0;
  int H, W, oy, ox, vec_per_row, vsh;
  bool interior;
  __device__ TileReader(const float* dem_, const TileGrid& grid, int2 org, int T) : dem(dem_), H(grid.H), W(grid.W), oy(org.x), ox(org.y) {
    vec_per_row = T / 4;
    vsh = (vec_per_row & (vec_per_row - 1)) == 0 ? __ffs(vec_per_row) - 1 : -1;
    interior = oy + T <= H && ox + T <= W && (W & 3) == 0 && (ox & 3) == 0 && (reinterpret_cast<uintptr_t>(dem) & 15u) == 0;
    tile0 = dem + (size_t)oy * (size_t)W + (size_t)ox;
  }
  __device__ __forceinline__ float4 load4v(int i) const {  // vector i of the tile, row-major
    int r, c;
    if (vsh >= 0) {
      r = i >> vsh;
      c = (i & (vec_per_row - 1)) * 4;
    } else {
      r = i / vec_per_row;
      c = (i - r * vec_per_row) * 4;
    }
    if (interior) return __ldg(reinterpret_cast<const float4*>(tile0 + (size_t)r * (size_t)W + c));
    return load4(dem, H, W, oy + r, ox + c);
  }
};

__global__ void norm_ws_init_kernel(unsigned* __restrict__ ws_meta, int n_tiles) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tiles) return;
  unsigned* m = ws_meta + (size_t)i * kMetaWords;
  m[kMetaCount] = 0u;
  m[kMetaMin] = 0x7f800000u;  // +Inf
  m[kMetaMax] = 0u;
  m[kMetaThr] = 0u;
  m[kMetaStatus] = 0u;
}

__global__ void __launch_bounds__(kThreads, 2)
tile_scan_kernel(const float* __restrict__ dem, const float* __restrict__ depth, TileGrid grid, int tile_base, int T, int TL, int scale,
                 fsr_tile_params p, float* __restrict__ depth_norm, unsigned* __restrict__ ws_meta, unsigned* __restrict__ ws_keys,
                 unsigned* __restrict__ flags_out) {
  __shared__ int hist[kBins];
  __shared__ int warp_tot[32];
  __shared__ float red_min[32], red_max[32];
  __shared__ int sel_bin, sel_before;
  __shared__ unsigned s_flags;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tile_local = blockIdx.x, slice = blockIdx.y;
  const int2 org = grid.origins[tile_base + tile_local];
  const TileReader rd(dem, grid, org, T);
  const int vec_per_row = T / 4, n_vec = T * vec_per_row, n_px = T * T;
  unsigned my_flags = 0;
  if (t == 0) s_flags = 0;

  if (slice == 0) {
    // ---- depth_lr: nodata -> 0, finite check, log1p scaling (preprocessing.py:141-151) ----------------
    for (int i = t; i < TL * TL; i += kThreads) {
      int r = i / TL, c = i - r * TL;
      int ly = org.x / scale + r, lx = org.y / scale + c;
      float x = (ly < grid.Hl && lx < grid.Wl) ? __ldg(depth + (size_t)ly * grid.Wl + lx) : 0.0f;
      if (p.has_depth_nodata) {
        bool hit = (x == p.depth_nodata) || (p.depth_nodata_tol >= 0.0f && fabsf(x - p.depth_nodata) <= p.depth_nodata_tol);
        if (hit) x = 0.0f;
      }
      if (!isfinite(x)) {
        my_flags |= FSR_FLAG_DEPTH_NONFINITE;
        x = 0.0f;
      }
      float d = fminf(fmaxf(x, 0.0f), p.max_depth);
      float s = __fdiv_rn(log1pf(d), p.depth_denom);
      depth_norm[(size_t)tile_local * TL * TL + i] = fminf(fmaxf(s, 0.0f), 1.0f);
    }
  }

  // ---- threshold from every kSampleRows-th tile row (the one-CTA kernel's estimate, histogrammed straight from L2) ----
  const int n_top = n_px - p.rank_lo;
  const int n_srows = T / kSampleRows, n_samp = n_srows * T;
  unsigned thr_key = 0;
  {
    float smin = INFINITY, smax = 0.0f;
    unsigned dummy = 0;
    for (int i = t; i < n_srows * vec_per_row; i += kThreads) {
      const int sr = i / vec_per_row, c = (i - sr * vec_per_row) * 4;
      const float4 v = load4(dem, grid.H, grid.W, org.x + sr * kSampleRows + kSampleRows / 2, org.y + c);
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float f = fmaxf(fix_dem(e[j], p, dummy), 0.0f) + 0.0f;
        smin = fminf(smin, f);
        smax = fmaxf(smax, f);
      }
    }
    smin = warp_min(smin);
    smax = warp_max(smax);
    if (lane == 0) {
      red_min[warp] = smin;
      red_max[warp] = smax;
    }
    for (int i = t; i < kBins; i += kThreads) hist[i] = 0;
    __syncthreads();
    smin = red_min[0];
    smax = red_max[0];
    for (int w = 1; w < kWarps; ++w) {
      smin = fminf(smin, red_min[w]);
      smax = fmaxf(smax, red_max[w]);
    }
    const unsigned klo = __float_as_uint(smin), khi = __float_as_uint(smax);
    if (khi > klo) {
      const int sh = digit_shift(klo, khi);
      for (int i = t; i < n_srows * vec_per_row; i += kThreads) {
        const int sr = i / vec_per_row, c = (i - sr * vec_per_row) * 4;
        const float4 v = load4(dem, grid.H, grid.W, org.x + sr * kSampleRows + kSampleRows / 2, org.y + c);
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) atomicAdd(&hist[(dem_key(e[j], p, dummy) - klo) >> sh], 1);
      }
      __syncthreads();
      const int want_above = (int)(((long long)n_top * 13 / 10 + 512) / kSampleRows);
      const int srank = n_samp - 1 - want_above;
      if (srank > 0) {
        find_bin(hist, srank, warp_tot, &sel_bin, &sel_before);
        thr_key = klo + ((unsigned)sel_bin << sh);
      }
    }
    __syncthreads();  // red arrays are reused below
  }
  unsigned* meta = ws_meta + (size_t)tile_local * kMetaWords;
  unsigned* keys = ws_keys + (size_t)tile_local * kCap;
  if (slice == 0 && t == 0) meta[kMetaThr] = thr_key;

  // ---- this slice's rows: finite checks, min / max, upper tail -> the tile's key list ------------------
  float vmin = INFINITY, vmax = 0.0f;
  const bool compact = thr_key > 0;
  const int v_begin = (n_vec / kSlices) * slice, v_end = slice == kSlices - 1 ? n_vec : v_begin + n_vec / kSlices;
  for (int i0 = v_begin + t; i0 < v_end; i0 += kThreads * kUnroll) {
    float4 vv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) vv[u] = (i0 + u * kThreads < v_end) ? rd.load4v(i0 + u * kThreads) : make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned k[kUnroll][4];
    int mine = 0;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const bool live = i0 + u * kThreads < v_end;
      const float4 v = vv[u];
      float e[4];
      e[0] = fmaxf(fix_dem(v.x, p, my_flags), 0.0f);
      e[1] = fmaxf(fix_dem(v.y, p, my_flags), 0.0f);
      e[2] = fmaxf(fix_dem(v.z, p, my_flags), 0.0f);
      e[3] = fmaxf(fix_dem(v.w, p, my_flags), 0.0f);
      if (live) {
        vmin = fminf(vmin, fminf(fminf(e[0], e[1]), fminf(e[2], e[3])));
        vmax = fmaxf(vmax, fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        k[u][j] = __float_as_uint(e[j] + 0.0f);
        mine += (live && compact && k[u][j] >= thr_key) ? 1 : 0;
      }
    }
    if (compact) {
      // warp-aggregated append: one global atomic per warp for kUnroll vectors per thread
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      if (total > 0) {
        int base = 0;
        if (lane == 31) base = (int)atomicAdd(&meta[kMetaCount], (unsigned)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + incl - mine;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const bool live = i0 + u * kThreads < v_end;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (live && k[u][j] >= thr_key) {
              if (pos < kCap) keys[pos] = k[u][j];
              ++pos;
            }
        }
      }
    }
  }
  vmin = warp_min(vmin);
  vmax = warp_max(vmax);
  if (lane == 0) {
    red_min[warp] = vmin;
    red_max[warp] = vmax;
  }
  if (my_flags) atomicOr(&s_flags, my_flags);
  __syncthreads();
  if (warp == 0) {
    const float a = warp_min(lane < kWarps ? red_min[lane] : INFINITY), b = warp_max(lane < kWarps ? red_max[lane] : 0.0f);
    if (lane == 0) {
      // values are >= +0: the bit patterns order like the values
      atomicMin(&meta[kMetaMin], __float_as_uint(a + 0.0f));
      atomicMax(&meta[kMetaMax], __float_as_uint(b + 0.0f));
      if (s_flags) atomicOr(flags_out, s_flags);
    }
  }
}

__global__ void __launch_bounds__(kThreads, 3)
tile_select_kernel(int tile_base, int T, fsr_tile_params p, unsigned* __restrict__ ws_meta, const unsigned* __restrict__ ws_keys,
                   float* __restrict__ stats, unsigned* __restrict__ flags_out) {
  __shared__ int hist[2][kBins];
  __shared__ int warp_tot[32];
  __shared__ SelState st[2];
  __shared__ int sel_bin, sel_before;
  const int t = threadIdx.x, lane = t & 31;
  const int tile_local = blockIdx.x;
  unsigned* meta = ws_meta + (size_t)tile_local * kMetaWords;
  const unsigned* keys = ws_keys + (size_t)tile_local * kCap;
  const int n_px = T * T, n_top = n_px - p.rank_lo;
  const int n_comp = (int)meta[kMetaCount];
  const unsigned thr_key = meta[kMetaThr];
  const float vmin = __uint_as_float(meta[kMetaMin]), vmax = __uint_as_float(meta[kMetaMax]);
  // the list is usable when it holds every key of rank >= rank_lo and did not overflow; otherwise the one-CTA kernel
  // (launched next) handles this tile from scratch
  if (!(thr_key > 0 && n_comp <= kCap && n_comp >= n_top)) return;
  if (t == 0) {
    const int below = n_px - n_comp;  // keys smaller than the threshold
    st[0] = SelState{thr_key, __float_as_uint(vmax + 0.0f), p.rank_lo - below};
    st[1] = SelState{thr_key, __float_as_uint(vmax + 0.0f), p.rank_hi - below};
  }
  __syncthreads();
  for (int guard = 0; guard < 8; ++guard) {
    const SelState s0 = st[0], s1 = st[1];
    const bool need0 = s0.hi > s0.lo, need1 = s1.hi > s1.lo;
    if (!need0 && !need1) break;
    const bool same = (s0.lo == s1.lo) && (s0.hi == s1.hi);
    const int sh0 = need0 ? digit_shift(s0.lo, s0.hi) : 0;
    const int sh1 = need1 ? digit_shift(s1.lo, s1.hi) : 0;
    for (int i = t; i < 2 * kBins; i += kThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    int lo0_hits = 0;
    const int n_round = (n_comp + 31) & ~31;
    for (int i = t; i < n_round; i += kThreads) {
      const bool valid = i < n_comp;
      hist_add(valid ? __ldg(keys + i) : 0u, valid, s0, s1, need0, need1, same, sh0, sh1, hist, lo0_hits);
    }
    if (need0 && lane == 0 && lo0_hits) atomicAdd(&hist[0][0], lo0_hits);
    __syncthreads();
    refine(st, s0, s1, need0, need1, same, sh0, sh1, hist, warp_tot, &sel_bin, &sel_before);
  }
  if (t == 0) {
    // numpy _lerp on the two order statistics (float32, no FMA contraction), as in the one-CTA kernel
    const float a = __uint_as_float(st[0].lo), b = __uint_as_float(st[1].lo);
    const float diff = __fsub_rn(b, a);
    float p_clip = __fadd_rn(a, __fmul_rn(diff, p.gamma));
    if (p.gamma >= 0.5f) p_clip = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, p.gamma)));
    const float dem_min = fminf(vmin, p_clip), dem_max = fminf(vmax, p_clip);
    const double range_d = (double)dem_max - (double)dem_min;
    if (!(range_d > 0.0) && !(fabs(range_d) <= 1e-8 && fabs((double)dem_min) <= 1e-8)) atomicOr(flags_out, FSR_FLAG_DEM_FLAT_NONZERO);
    float* s = stats + (size_t)(tile_base + tile_local) * 3;
    s[0] = p_clip;
    s[1] = dem_min;
    s[2] = dem_max;
    meta[kMetaStatus] = kStatusDone;
  }
}

// pass N of the one-CTA kernel for 16 tile rows (one row of pooled cells): same thread <-> pixel mapping, same sums
__global__ void __launch_bounds__(kThreads, 3)
tile_apply_kernel(const float* __restrict__ dem, TileGrid grid, int tile_base, int T, int TL, fsr_tile_params p,
                  const unsigned* __restrict__ ws_meta, const float* __restrict__ stats, float* __restrict__ dem_norm,
                  float* __restrict__ dem_lr) {
  __shared__ float pool_part[4][128];
  const int t = threadIdx.x, lane = t & 31;
  const int tile_local = blockIdx.x;
  if (ws_meta[(size_t)tile_local * kMetaWords + kMetaStatus] != kStatusDone) return;
  const int2 org = grid.origins[tile_base + tile_local];
  const TileReader rd(dem, grid, org, T);
  const float* s = stats + (size_t)(tile_base + tile_local) * 3;
  const float p_clip = s[0], dem_min = s[1], dem_max = s[2];
  const double range_d = (double)dem_max - (double)dem_min;  // python-float subtraction in the reference
  const bool zero_out = !(range_d > 0.0);
  const float range_f = (float)range_d;
  float4* out = dem_norm ? reinterpret_cast<float4*>(dem_norm + (size_t)tile_local * T * T) : nullptr;
  const int i0 = blockIdx.y * (kThreads * kUnroll) + t;
  float4 vv[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) vv[u] = rd.load4v(i0 + u * kThreads);
  float psum[kUnroll / 4];
#pragma unroll
  for (int h = 0; h < kUnroll / 4; ++h) psum[h] = 0.0f;
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    float e[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = fminf(fmaxf(fix_nodata(e[j], p), 0.0f), p_clip);
      float n = __fdiv_rn(__fsub_rn(x, dem_min), range_f);
      e[j] = zero_out ? 0.0f : fminf(fmaxf(n, 0.0f), 1.0f);
    }
    if (out) out[i0 + u * kThreads] = make_float4(e[0], e[1], e[2], e[3]);
    psum[u / 4] += (e[0] + e[1]) + (e[2] + e[3]);
  }
  if (dem_lr) {
#pragma unroll
    for (int h = 0; h < kUnroll / 4; ++h) {
      float q = psum[h] + __shfl_xor_sync(0xffffffffu, psum[h], 1);
      q += __shfl_xor_sync(0xffffffffu, q, 2);
      if ((lane & 3) == 0) pool_part[t >> 7][(t & 127) >> 2] = q;  // [row offset][cell]
      __syncthreads();
      if (t < 32) {
        const float sum = (pool_part[0][t] + pool_part[1][t]) + (pool_part[2][t] + pool_part[3][t]);
        dem_lr[(size_t)tile_local * (TL * TL) + (blockIdx.y * (kUnroll / 4) + h) * TL + t] = sum * (1.0f / 256.0f);
      }
      __syncthreads();
    }
  }
}

// normalize_inputs == 0 (ort.py:163-180): inputs are used as they are; finiteness and [0,1] range are asserted.
__global__ void tile_passthrough_kernel(const float* __restrict__ dem, const float* __restrict__ depth, TileGrid grid,
                                        int tile_base, int T, int TL, int scale, float* __restrict__ dem_norm,
                                        float* __restrict__ depth_norm, unsigned* __restrict__ flags_out) {
  const int tile_local = blockIdx.x;
  const int2 org = grid.origins[tile_base + tile_local];
  unsigned f = 0;
  for (int i = threadIdx.x; i < TL * TL; i += blockDim.x) {
    int r = i / TL, c = i - r * TL;
    int ly = org.x / scale + r, lx = org.y / scale + c;
    float x = (ly < grid.Hl && lx < grid.Wl) ? __ldg(depth + (size_t)ly * grid.Wl + lx) : 0.0f;
    if (!isfinite(x)) f |= FSR_FLAG_DEPTH_NONFINITE;
    else if (x < 0.0f || x > 1.0f) f |= FSR_FLAG_DEPTH_NOT_UNIT;
    depth_norm[(size_t)tile_local * TL * TL + i] = x;
  }
  for (int i = threadIdx.x; i < T * T; i += blockDim.x) {
    int r = i / T, c = i - r * T;
    int gy = org.x + r, gx = org.y + c;
    float x = (gy < grid.H && gx < grid.W) ? __ldg(dem + (size_t)gy * grid.W + gx) : 0.0f;
    if (!isfinite(x)) f |= FSR_FLAG_DEM_NONFINITE;
    else if (x < 0.0f || x > 1.0f) f |= FSR_FLAG_DEM_NOT_UNIT;
    dem_norm[(size_t)tile_local * T * T + i] = x;
  }
  if (f) atomicOr(flags_out, f);
}

}  // namespace

size_t tile_normalize_ws_bytes(int n_tiles) { return (size_t)n_tiles * (kMetaWords + kCap) * sizeof(unsigned); }

void launch_tile_normalize(const float* d_dem, const float* d_depth, const TileGrid& grid, int tile_base, int n_tiles,
                           int T, int TL, int scale, const fsr_tile_params& p, float* d_dem_norm, float* d_depth_norm,
                           float* d_stats, float* d_dem_lr, unsigned* d_flags, cudaStream_t stream, void* d_ws) {
  if (n_tiles <= 0) return;
  FSR_REQUIRE(T % 4 == 0 && T * (T / 4) % kThreads == 0, "hr tile must be a multiple of 64 pixels");
  if (p.normalize_inputs) {
    static bool attr[64] = {false};
    if (first_on_device(attr))
      FSR_CUDA(cudaFuncSetAttribute(tile_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCap * (int)sizeof(unsigned)));
    // three-kernel path (see tile_scan_kernel): the usual geometry and percentile (upper tail small enough for the key list),
    // tile statistics computed here (not supplied by the caller); the one-CTA kernel then runs for the tiles it left open
    const int vec_per_row = T / 4, n_top = T * T - p.rank_lo;
    // (measured on B200: 85 tiles 181 -> 121 us; from ~2 tiles per SM upwards the one-CTA kernel's single pass over HBM wins:
    // 935 tiles 0.71 vs 0.78 ms.  Both paths produce the same bits, so the choice may follow the batch size.)
    const char* force = getenv("FSR_NORM_SPLIT");  // "0": never, "1": always (tests)
    const bool small = force ? atoi(force) != 0 : n_tiles <= 2 * current_sm_count();
    const bool split = d_ws && small && !p.has_ref_stats && T == 512 && scale == 16 && kUnroll % 4 == 0 &&
                       (long long)n_top * 14 / 10 + 2048 <= kCap && T % kSampleRows == 0 && (T / kSampleRows) * T <= kCap &&
                       ((T / kSampleRows) * vec_per_row) % kThreads == 0 && (T * vec_per_row / kSlices) % (kThreads * kUnroll) == 0;
    unsigned* meta = nullptr;
    if (split) {
      meta = reinterpret_cast<unsigned*>(d_ws);
      unsigned* keys = meta + (size_t)n_tiles * kMetaWords;
      norm_ws_init_kernel<<<ceil_div(n_tiles, 256), 256, 0, stream>>>(meta, n_tiles);
      tile_scan_kernel<<<dim3((unsigned)n_tiles, kSlices), kThreads, 0, stream>>>(d_dem, d_depth, grid, tile_base, T, TL, scale, p,
                                                                               d_depth_norm, meta, keys, d_flags);
      tile_select_kernel<<<n_tiles, kThreads, 0, stream>>>(tile_base, T, p, meta, keys, d_stats, d_flags);
      tile_apply_kernel<<<dim3((unsigned)n_tiles, (unsigned)(T * vec_per_row / (kThreads * kUnroll))), kThreads, 0, stream>>>(
          d_dem, grid, tile_base, T, TL, p, meta, d_stats, d_dem_norm, d_dem_lr);
      FSR_LAUNCH_CHECK();
    }
    tile_normalize_kernel<<<n_tiles, kThreads, kCap * sizeof(unsigned), stream>>>(d_dem, d_depth, grid, tile_base, T, TL, scale, p,
                                                                                 d_dem_norm, d_depth_norm, d_stats, d_dem_lr, d_flags, meta);
  } else {
    tile_passthrough_kernel<<<n_tiles, 1024, 0, stream>>>(d_dem, d_depth, grid, tile_base, T, TL, scale, d_dem_norm,
                                                          d_depth_norm, d_flags);
  }
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
