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

constexpr int kThreads = 1024;
constexpr int kBins = 2048;
constexpr int kDigitBits = 11;

struct SelState {
  unsigned lo, hi;  // inclusive key range still containing the wanted order statistic
  int rank;         // rank of the wanted element among keys in [lo, hi]
};

__device__ __forceinline__ float fix_dem(float x, const fsr_tile_params& p, unsigned& flags) {
  if (p.has_dem_nodata) {
    bool hit = (x == p.dem_nodata) || (p.dem_nodata_tol >= 0.0f && fabsf(x - p.dem_nodata) <= p.dem_nodata_tol);
    if (hit) x = 0.0f;
  }
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

// Find, for `rank`, the histogram bin whose cumulative count first exceeds it.  All threads call this.
// Result is published through smem (*sel_bin, *sel_before).
__device__ void find_bin(const int* __restrict__ hist, int rank, int* warp_tot, int* sel_bin, int* sel_before) {
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int h0 = hist[2 * t], h1 = hist[2 * t + 1];
  int incl = h0 + h1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
    int s = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    warp_tot[lane] = s - w;  // exclusive prefix of warp totals
  }
  __syncthreads();
  const int excl = warp_tot[warp] + incl - (h0 + h1);
  if (rank >= excl && rank < excl + h0) {
    *sel_bin = 2 * t;
    *sel_before = excl;
  } else if (rank >= excl + h0 && rank < excl + h0 + h1) {
    *sel_bin = 2 * t + 1;
    *sel_before = excl + h0;
  }
  __syncthreads();
}

__device__ __forceinline__ int digit_shift(unsigned lo, unsigned hi) {
  unsigned span = hi - lo;
  int bits = 32 - __clz(span);  // span > 0 here
  return bits > kDigitBits ? bits - kDigitBits : 0;
}

__global__ void __launch_bounds__(kThreads, 1)
tile_normalize_kernel(const float* __restrict__ dem, const float* __restrict__ depth, TileGrid grid, int tile_base,
                      int T, int TL, int scale, fsr_tile_params p, float* __restrict__ dem_norm,
                      float* __restrict__ depth_norm, float* __restrict__ stats, unsigned* __restrict__ flags_out) {
  __shared__ int hist[2][kBins];
  __shared__ int warp_tot[32];
  __shared__ float red_min[32], red_max[32];
  __shared__ SelState st[2];
  __shared__ int sel_bin, sel_before;
  __shared__ unsigned s_flags;

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int tile_local = blockIdx.x;
  const int2 org = grid.origins[tile_base + tile_local];
  const int vec_per_row = T / 4;
  const int n_vec = T * vec_per_row;
  unsigned my_flags = 0;
  if (t == 0) s_flags = 0;

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

  // ---- pass A: min / max of clip(x, 0, inf) ----------------------------------------------------------
  float vmin = INFINITY, vmax = 0.0f;
  const bool own_stats = !p.has_ref_stats;
  for (int i = t; i < n_vec; i += kThreads) {
    int r = i / vec_per_row, c = (i - r * vec_per_row) * 4;
    float4 v = load4(dem, grid.H, grid.W, org.x + r, org.y + c);
    float a = fmaxf(fix_dem(v.x, p, my_flags), 0.0f), b = fmaxf(fix_dem(v.y, p, my_flags), 0.0f);
    float cc = fmaxf(fix_dem(v.z, p, my_flags), 0.0f), d = fmaxf(fix_dem(v.w, p, my_flags), 0.0f);
    vmin = fminf(vmin, fminf(fminf(a, b), fminf(cc, d)));
    vmax = fmaxf(vmax, fmaxf(fmaxf(a, b), fmaxf(cc, d)));
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
    float a = warp_min(red_min[lane]), b = warp_max(red_max[lane]);
    if (lane == 0) {
      red_min[0] = a;
      red_max[0] = b;
      // +0.0f canonicalises -0.0 so that keys are the plain bit patterns of non-negative floats
      st[0] = SelState{__float_as_uint(a + 0.0f), __float_as_uint(b + 0.0f), p.rank_lo};
      st[1] = SelState{__float_as_uint(a + 0.0f), __float_as_uint(b + 0.0f), p.rank_hi};
    }
  }
  __syncthreads();
  vmin = red_min[0];
  vmax = red_max[0];

  // ---- passes B..: radix select of sorted[rank_lo] and sorted[rank_hi] ------------------------------
  for (int guard = 0; own_stats && guard < 8; ++guard) {
    const SelState s0 = st[0], s1 = st[1];
    const bool need0 = s0.hi > s0.lo, need1 = s1.hi > s1.lo;
    if (!need0 && !need1) break;
    const bool same = (s0.lo == s1.lo) && (s0.hi == s1.hi);
    const int sh0 = need0 ? digit_shift(s0.lo, s0.hi) : 0;
    const int sh1 = need1 ? digit_shift(s1.lo, s1.hi) : 0;
    for (int i = t; i < 2 * kBins; i += kThreads) (&hist[0][0])[i] = 0;
    __syncthreads();
    int lo0_hits = 0;  // keys equal to the range start are counted by ballot: zero padding would otherwise
                       // serialise hundreds of thousands of same-address shared atomics
    for (int i = t; i < n_vec; i += kThreads) {
      int r = i / vec_per_row, c = (i - r * vec_per_row) * 4;
      float4 v = load4(dem, grid.H, grid.W, org.x + r, org.y + c);
      unsigned dummy = 0;
      float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        unsigned key = __float_as_uint(fmaxf(fix_dem(e[j], p, dummy), 0.0f) + 0.0f);
        if (need0) {
          bool at_lo = (key == s0.lo);
          lo0_hits += __popc(__ballot_sync(0xffffffffu, at_lo));
          if (!at_lo && key > s0.lo && key <= s0.hi) atomicAdd(&hist[0][(key - s0.lo) >> sh0], 1);
        }
        if (need1 && !same) {
          if (key >= s1.lo && key <= s1.hi) atomicAdd(&hist[1][(key - s1.lo) >> sh1], 1);
        }
      }
    }
    if (need0 && lane == 0 && lo0_hits) atomicAdd(&hist[0][0], lo0_hits);
    __syncthreads();
    if (need0) {
      find_bin(hist[0], s0.rank, warp_tot, &sel_bin, &sel_before);
      if (t == 0) {
        unsigned nlo = s0.lo + ((unsigned)sel_bin << sh0);
        unsigned width = (sh0 >= 32) ? 0xffffffffu : ((1u << sh0) - 1u);
        unsigned nhi = nlo + width;
        if (nhi > s0.hi || nhi < nlo) nhi = s0.hi;
        st[0] = SelState{nlo, nhi, s0.rank - sel_before};
      }
      __syncthreads();
    }
    if (need1) {
      const int* h = same ? hist[0] : hist[1];
      const int sh = same ? sh0 : sh1;
      find_bin(h, s1.rank, warp_tot, &sel_bin, &sel_before);
      if (t == 0) {
        unsigned nlo = s1.lo + ((unsigned)sel_bin << sh);
        unsigned width = (1u << sh) - 1u;
        unsigned nhi = nlo + width;
        if (nhi > s1.hi || nhi < nlo) nhi = s1.hi;
        st[1] = SelState{nlo, nhi, s1.rank - sel_before};
      }
      __syncthreads();
    }
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
  float4* out = reinterpret_cast<float4*>(dem_norm + (size_t)tile_local * T * T);
  for (int i = t; i < n_vec; i += kThreads) {
    int r = i / vec_per_row, c = (i - r * vec_per_row) * 4;
    float4 v = load4(dem, grid.H, grid.W, org.x + r, org.y + c);
    unsigned dummy = 0;
    float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = fminf(fmaxf(fix_dem(e[j], p, dummy), 0.0f), p_clip);
      float n = __fdiv_rn(__fsub_rn(x, dem_min), range_f);
      e[j] = zero_out ? 0.0f : fminf(fmaxf(n, 0.0f), 1.0f);
    }
    out[i] = make_float4(e[0], e[1], e[2], e[3]);
  }
  __syncthreads();
  if (t == 0 && s_flags) atomicOr(flags_out, s_flags);
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

void launch_tile_normalize(const float* d_dem, const float* d_depth, const TileGrid& grid, int tile_base, int n_tiles,
                           int T, int TL, int scale, const fsr_tile_params& p, float* d_dem_norm, float* d_depth_norm,
                           float* d_stats, unsigned* d_flags, cudaStream_t stream) {
  if (n_tiles <= 0) return;
  FSR_REQUIRE(T % 4 == 0 && T * (T / 4) % kThreads == 0, "hr tile must be a multiple of 64 pixels");
  if (p.normalize_inputs) {
    tile_normalize_kernel<<<n_tiles, kThreads, 0, stream>>>(d_dem, d_depth, grid, tile_base, T, TL, scale, p, d_dem_norm,
                                                            d_depth_norm, d_stats, d_flags);
  } else {
    tile_passthrough_kernel<<<n_tiles, 1024, 0, stream>>>(d_dem, d_depth, grid, tile_base, T, TL, scale, d_dem_norm,
                                                          d_depth_norm, d_flags);
  }
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
