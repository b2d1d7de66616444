// Overlap-blend mosaic of tile predictions (memory-bound stage).
//
// Replaces the per-window `accum += pred*outer(wy,wx); weight_sum += outer(wy,wx)` read-modify-write loop
// and the final `accum / max(weight_sum, 1e-6)`, crop and clip of
// ModelWorker._run_tiled_model_on_prepared (floodsr/models/ResUNet_16x_DEM.py:335-363, :391), and the
// direct paste of the "hard" method (:297-314).
//
// Gather formulation: one thread owns VEC adjacent output pixels and visits the windows covering them in
// the reference's row-major window order, so the float32 sums are formed in exactly the reference's order
// (bit-exact for identical tile predictions) with no atomics and no weight_sum/accum rasters in HBM.
// For row-band sharding the sum may start from the neighbour band's partial sums (`init`) and may be
// restricted to this band's tile rows.
#include "fsr_engine.cuh"

namespace fsr {

namespace {

__device__ __forceinline__ float edge_weight(const float* __restrict__ ramp, int local, int idx, int n, int T, int overlap) {
  // feather ramp with the scene-edge flattening of ResUNet_16x_DEM.py:344-352
  if (idx == 0 && local < overlap) return 1.0f;
  if (idx == n - 1 && local >= T - overlap) return 1.0f;
  return __ldg(ramp + local);
}

template <int VEC>
__global__ void __launch_bounds__(256)
blend_kernel(const float* __restrict__ tiles, int ty0, int ty1, BlendGeom g, int row0, int n_rows,
             const float* __restrict__ init, int init_rows, int finalize, float max_depth, float* __restrict__ out) {
  const int xv = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const int ry = blockIdx.y;  // row within [0, n_rows)
  if (xv >= g.W || ry >= n_rows) return;
  const int y = row0 + ry;
  const int yf = g.y_first[y], yc = g.y_count[y];
  const int xf = g.x_first[xv], xc = g.x_count[xv];
  const size_t tile_px = (size_t)g.T * g.T;
  float acc[VEC], wsum[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    acc[v] = 0.0f;
    wsum[v] = 0.0f;
  }
  if (init && ry < init_rows) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = __ldg(init + (size_t)ry * g.W + xv + v);
  }
  for (int yi = yf; yi < yf + yc; ++yi) {
    const int ly = y - g.y_starts[yi];
    const float wy = g.ramp ? edge_weight(g.ramp, ly, yi, g.ny, g.T, g.overlap) : 1.0f;
    const bool mine = yi >= ty0 && yi < ty1;
    for (int xi = xf; xi < xf + xc; ++xi) {
      const int lx = xv - g.x_starts[xi];
      float p[VEC];
      if (mine) {
        const float* src = tiles + ((size_t)(yi - ty0) * g.nx + xi) * tile_px + (size_t)ly * g.T + lx;
        if (VEC == 4) {
          float4 q = __ldcs(reinterpret_cast<const float4*>(src));
          p[0] = q.x; p[1 % VEC] = q.y; p[2 % VEC] = q.z; p[3 % VEC] = q.w;
        } else {
          p[0] = __ldcs(src);
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float wx = g.ramp ? edge_weight(g.ramp, lx + v, xi, g.nx, g.T, g.overlap) : 1.0f;
        const float w = __fmul_rn(wy, wx);
        wsum[v] = __fadd_rn(wsum[v], w);
        if (mine) acc[v] = __fadd_rn(acc[v], __fmul_rn(p[v], w));
      }
    }
  }
  float r[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    if (finalize) {
      float q = wsum[v] > 0.0f ? __fdiv_rn(acc[v], fmaxf(wsum[v], 1e-6f)) : 0.0f;
      r[v] = fminf(fmaxf(q, 0.0f), max_depth);
    } else {
      r[v] = acc[v];
    }
  }
  float* dstp = out + (size_t)ry * g.W + xv;
  if (VEC == 4) {
    *reinterpret_cast<float4*>(dstp) = make_float4(r[0], r[1 % VEC], r[2 % VEC], r[3 % VEC]);
  } else {
    dstp[0] = r[0];
  }
}

}  // namespace

void launch_blend(const float* d_tiles, int ty0, int ty1, const BlendGeom& g, int row0, int n_rows, const float* d_init,
                  int init_rows, bool finalize, float max_depth, float* d_out, cudaStream_t s) {
  if (n_rows <= 0) return;
  const bool vec = g.vec_ok && (g.W % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 15u) == 0) &&
                   (!d_init || (reinterpret_cast<uintptr_t>(d_init) & 15u) == 0);
  // gridDim.y is limited to 65535 rows per launch
  for (int r0 = 0; r0 < n_rows; r0 += 65535) {
    const int nr = n_rows - r0 < 65535 ? n_rows - r0 : 65535;
    const float* init = (d_init && r0 < init_rows) ? d_init + (size_t)r0 * g.W : nullptr;
    const int irows = init ? init_rows - r0 : 0;
    float* outp = d_out + (size_t)r0 * g.W;
    if (vec) {
      dim3 grid((unsigned)ceil_div(g.W / 4, 256), (unsigned)nr);
      blend_kernel<4><<<grid, 256, 0, s>>>(d_tiles, ty0, ty1, g, row0 + r0, nr, init, irows, finalize ? 1 : 0, max_depth, outp);
    } else {
      dim3 grid((unsigned)ceil_div(g.W, 256), (unsigned)nr);
      blend_kernel<1><<<grid, 256, 0, s>>>(d_tiles, ty0, ty1, g, row0 + r0, nr, init, irows, finalize ? 1 : 0, max_depth, outp);
    }
    FSR_LAUNCH_CHECK();
  }
}

}  // namespace fsr
