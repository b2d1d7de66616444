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
             const float* __restrict__ init, int init_rows, int finalize, float max_depth, float* __restrict__ out, int x0, int x1) {
  const int xv = x0 + (blockIdx.x * blockDim.x + threadIdx.x) * VEC;  // columns [x0, x1) of the raster
  const int ry = blockIdx.y;  // row within [0, n_rows)
  if (xv >= x1 || ry >= n_rows) return;
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

// Fast variant for the usual geometry (every pixel is covered by at most 2 window rows and 2 window columns, float4 path):
// each thread owns 4 pixels of R consecutive rows; the column-side lookups (window indices, feather weights) are done once
// for all rows, and the up to 4 R tile loads are issued before any is consumed (the kernel is HBM-latency bound).
// The per-pixel accumulation order is the generic kernel's (window rows outer, columns inner): results are bit-identical.
template <int R>
__global__ void __launch_bounds__(256)
blend_fast_kernel(const float* __restrict__ tiles, int ty0, int ty1, BlendGeom g, int row0, int n_rows,
                  const float* __restrict__ init, int init_rows, int finalize, float max_depth, float* __restrict__ out, int x0, int x1) {
  const int xv = x0 + (blockIdx.x * blockDim.x + threadIdx.x) * 4;  // columns [x0, x1) of the raster
  if (xv >= x1) return;
  const size_t tile_px = (size_t)g.T * g.T;
  // column side
  const int xf = g.x_first[xv], xc = g.x_count[xv];
  int lx[2];
  float wx[2][4];
#pragma unroll
  for (int dx = 0; dx < 2; ++dx) {
    const int xi = xf + dx;
    lx[dx] = dx < xc ? xv - g.x_starts[xi] : 0;
    // window origins are multiples of 4 on this path, so the 4 pixels share their covering windows
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(g.wx_tab + (size_t)dx * g.Wpad + xv));
    wx[dx][0] = w4.x; wx[dx][1] = w4.y; wx[dx][2] = w4.z; wx[dx][3] = w4.w;
  }
  // row side + loads; the block walks down the raster in steps of gridDim.y row groups (column-side state is reused)
  for (int rbase = blockIdx.y * R; rbase < n_rows; rbase += gridDim.y * R) {
  float4 q[R][2][2];
  float wy[R][2];
  bool use[R][2];   // window row contributes to the weight sum
  bool mine[R][2];  // ... and its tiles belong to this band
  bool rok[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    rok[r] = rbase + r < n_rows;
    const int y = row0 + rbase + (rok[r] ? r : 0);
    const int yf = g.y_first[y], yc = g.y_count[y];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int yi = yf + dy;
      use[r][dy] = rok[r] && dy < yc;
      const int ly = use[r][dy] ? y - g.y_starts[yi] : 0;
      wy[r][dy] = __ldg(g.wy_tab + (size_t)dy * g.Hpad + y);
      mine[r][dy] = use[r][dy] && yi >= ty0 && yi < ty1;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        q[r][dy][dx] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine[r][dy] && dx < xc)
          q[r][dy][dx] = __ldcs(reinterpret_cast<const float4*>(tiles + ((size_t)(yi - ty0) * g.nx + xf + dx) * tile_px + (size_t)ly * g.T + lx[dx]));
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (!rok[r]) continue;
    const int ry = rbase + r;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, wsum[4] = {0.f, 0.f, 0.f, 0.f};
    if (init && ry < init_rows) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(init + (size_t)ry * g.W + xv));
      acc[0] = a.x; acc[1] = a.y; acc[2] = a.z; acc[3] = a.w;
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      if (!use[r][dy]) continue;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        if (dx >= xc) continue;
        const float p[4] = {q[r][dy][dx].x, q[r][dy][dx].y, q[r][dy][dx].z, q[r][dy][dx].w};
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float w = __fmul_rn(wy[r][dy], wx[dx][v]);
          wsum[v] = __fadd_rn(wsum[v], w);
          if (mine[r][dy]) acc[v] = __fadd_rn(acc[v], __fmul_rn(p[v], w));
        }
      }
    }
    float o[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      if (finalize) {
        // x / 1 == x exactly: pixels covered by one window (or on flattened scene edges) skip the IEEE division
        const float t = wsum[v] == 1.0f ? acc[v] : (wsum[v] > 0.0f ? __fdiv_rn(acc[v], fmaxf(wsum[v], 1e-6f)) : 0.0f);
        o[v] = fminf(fmaxf(t, 0.0f), max_depth);
      } else {
        o[v] = acc[v];
      }
    }
    *reinterpret_cast<float4*>(out + (size_t)ry * g.W + xv) = make_float4(o[0], o[1], o[2], o[3]);
  }
  }
}

}  // namespace

void launch_blend(const float* d_tiles, int ty0, int ty1, const BlendGeom& g, int row0, int n_rows, const float* d_init,
                  int init_rows, bool finalize, float max_depth, float* d_out, cudaStream_t s, int x0, int x1) {
  if (x1 < 0) x1 = g.W;  // columns [x0, x1) of every row (d_out / d_init still point at column 0)
  if (n_rows <= 0 || x1 <= x0) return;
  FSR_REQUIRE(x0 >= 0 && x1 <= g.W, "blend columns outside the raster");
  const bool vec = g.vec_ok && (g.W % 4 == 0) && x0 % 4 == 0 && x1 % 4 == 0 && ((reinterpret_cast<uintptr_t>(d_out) & 15u) == 0) &&
                   (!d_init || (reinterpret_cast<uintptr_t>(d_init) & 15u) == 0);
  const int wx = x1 - x0;
  // gridDim.y is limited to 65535 rows per launch
  for (int r0 = 0; r0 < n_rows; r0 += 65535) {
    const int nr = n_rows - r0 < 65535 ? n_rows - r0 : 65535;
    const float* init = (d_init && r0 < init_rows) ? d_init + (size_t)r0 * g.W : nullptr;
    const int irows = init ? init_rows - r0 : 0;
    float* outp = d_out + (size_t)r0 * g.W;
    if (vec && g.max_cover <= 2) {
      constexpr int R = 2;
      // ~32 blocks per SM, each walking down the rows (measured on B200: 4 -> 0.50 ms, 8 -> 0.39, 32 -> 0.37, one block per
      // row pair -> 0.43); env FSR_BLEND_BLOCKS overrides
      const int gx = ceil_div(wx / 4, 256);
      static const int per_sm = getenv("FSR_BLEND_BLOCKS") ? atoi(getenv("FSR_BLEND_BLOCKS")) : 32;
      const int gy = std::min(ceil_div(nr, R), std::max(1, (current_sm_count() * per_sm) / gx));
      dim3 grid((unsigned)gx, (unsigned)gy);
      blend_fast_kernel<R><<<grid, 256, 0, s>>>(d_tiles, ty0, ty1, g, row0 + r0, nr, init, irows, finalize ? 1 : 0, max_depth, outp, x0, x1);
    } else if (vec) {
      dim3 grid((unsigned)ceil_div(wx / 4, 256), (unsigned)nr);
      blend_kernel<4><<<grid, 256, 0, s>>>(d_tiles, ty0, ty1, g, row0 + r0, nr, init, irows, finalize ? 1 : 0, max_depth, outp, x0, x1);
    } else {
      dim3 grid((unsigned)ceil_div(wx, 256), (unsigned)nr);
      blend_kernel<1><<<grid, 256, 0, s>>>(d_tiles, ty0, ty1, g, row0 + r0, nr, init, irows, finalize ? 1 : 0, max_depth, outp, x0, x1);
    }
    FSR_LAUNCH_CHECK();
  }
}

}  // namespace fsr
