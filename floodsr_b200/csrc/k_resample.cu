// Bilinear change of grid between two north-up rasters in the same CRS (memory-bound gather, SURVEY.md section 8 f#2).
//
// Replaces the two rasterio.warp.reproject(..., resampling=Resampling.bilinear) calls around the tile loop:
// floodsr/preprocessing.py:371-387 (raw DEM grid -> model grid, nodata aware) and
// floodsr/models/ResUNet_16x_DEM.py:554-573 (prediction on the model grid -> raw DEM grid).  The arithmetic those calls
// reach lives in GDAL's warp kernel (alg/gdalwarpkernel.cpp), which is not part of the reference tree; this kernel follows
// the same restatement as oracle/resample_np.py (4-sample formula when neither axis is down-sampled below 0.95, the
// general kernel with the triangle filter widened by 1 / scale otherwise; float64 accumulation; neighbours outside the
// raster or equal to the source nodata are skipped and the sum renormalised) operation for operation, so the two agree
// bit for bit.  Parity with GDAL itself is unpinned.
//
// One thread per destination pixel; every source pixel is read by the few destination pixels around it, so the reads are
// served by L1 / L2 and the kernel moves ~4 B in + 4 B out per destination pixel.
#include "fsr_engine.cuh"

namespace fsr {

namespace {

struct AxisMap {
  double a_dst, c_dst, a_src, c_src;  // source coordinate of destination index i: ((a_dst * (i + 0.5) + c_dst) - c_src) / a_src
  double k;                           // filter scale of the general kernel: min(destination pixels per source pixel, 1)
  int radius;                         // ceil(1 / k)
  int n_src;
};

__device__ __forceinline__ double axis_coord(const AxisMap& m, int i) {
  // explicit rounding at every step: no FMA contraction, identical to the float64 numpy expression
  const double t = __dadd_rn(__dmul_rn(m.a_dst, __dadd_rn((double)i, 0.5)), m.c_dst);
  return __ddiv_rn(__dsub_rn(t, m.c_src), m.a_src);
}

// weight of tap `off` (relative to base = floor(s - 0.5)); 0 when the tap lies outside the raster
__device__ __forceinline__ double tap_weight(const AxisMap& m, bool four, double s, double base, int off) {
  const int idx = (int)base + off;
  if (idx < 0 || idx >= m.n_src) return 0.0;
  if (four) {
    const double ratio = __dsub_rn(1.5, __dsub_rn(s, base));
    return off == 0 ? ratio : __dsub_rn(1.0, ratio);
  }
  const double delta = __dsub_rn(__dsub_rn(s, 0.5), base);
  const double w = __dsub_rn(1.0, fabs(__dmul_rn(__dsub_rn((double)off, delta), m.k)));
  return w > 0.0 ? w : 0.0;
}

__global__ void __launch_bounds__(256)
resample_bilinear_kernel(const float* __restrict__ src, int sw, float* __restrict__ dst, int dh, int dw, AxisMap mx, AxisMap my,
                         int four, int has_nodata, float src_nodata, float fill) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= dw || j >= dh) return;
  const double sx = axis_coord(mx, i), sy = axis_coord(my, j);
  float out = fill;
  if (sx >= 0.0 && sx < (double)mx.n_src && sy >= 0.0 && sy < (double)my.n_src) {
    const double bx = floor(__dsub_rn(sx, 0.5)), by = floor(__dsub_rn(sy, 0.5));
    const int x_lo = four ? 0 : -mx.radius, x_hi = four ? 1 : mx.radius;
    const int y_lo = four ? 0 : -my.radius, y_hi = four ? 1 : my.radius;
    double acc = 0.0, wsum = 0.0;
    for (int oy = y_lo; oy <= y_hi; ++oy) {
      const double wy = tap_weight(my, four, sy, by, oy);
      if (wy == 0.0) continue;
      const float* row = src + (size_t)((int)by + oy) * sw;
      for (int ox = x_lo; ox <= x_hi; ++ox) {
        const double wx = tap_weight(mx, four, sx, bx, ox);
        if (wx == 0.0) continue;
        const float v = __ldg(row + (int)bx + ox);
        if (has_nodata && v == src_nodata) continue;
        const double w = __dmul_rn(wx, wy);
        acc = __dadd_rn(acc, __dmul_rn((double)v, w));
        wsum = __dadd_rn(wsum, w);
      }
    }
    if (wsum >= (four ? 1e-5 : 1e-6)) out = (float)(wsum == 1.0 ? acc : __ddiv_rn(acc, wsum));
  }
  dst[(size_t)j * dw + i] = out;
}

}  // namespace

void launch_resample_bilinear(const float* d_src, int sh, int sw, float* d_dst, int dh, int dw, const fsr_resample_params& p,
                              cudaStream_t s) {
  AxisMap mx{p.x_a_dst, p.x_c_dst, p.x_a_src, p.x_c_src, 1.0, 1, sw};
  AxisMap my{p.y_a_dst, p.y_c_dst, p.y_a_src, p.y_c_src, 1.0, 1, sh};
  const double x_scale = fabs(p.x_a_src / p.x_a_dst), y_scale = fabs(p.y_a_src / p.y_a_dst);
  const int four = x_scale >= 0.95 && y_scale >= 0.95;
  mx.k = x_scale < 1.0 ? x_scale : 1.0;
  my.k = y_scale < 1.0 ? y_scale : 1.0;
  mx.radius = (int)ceil(1.0 / mx.k);
  my.radius = (int)ceil(1.0 / my.k);
  dim3 grid((unsigned)ceil_div(dw, 256), (unsigned)dh);
  resample_bilinear_kernel<<<grid, 256, 0, s>>>(d_src, sw, d_dst, dh, dw, mx, my, four, p.has_src_nodata, p.src_nodata, p.dst_fill);
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
