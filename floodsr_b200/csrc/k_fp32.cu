// fp32 CUDA-core layer kernels (FSR_PREC_FP32): the accuracy path of the network forward pass.
//
// These replace ONNX Runtime's CPU kernels behind `session.run` (floodsr/engine/ort.py:193) with plain
// float32 FMA arithmetic so that predictions stay within 1e-4 m of the reference; the throughput path is
// the tcgen05 bf16 backend (k_tc.cu).  All tensors are NHWC float32, batched over tiles.
#include "fsr_common.cuh"

namespace fsr {

namespace {

// ------------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution: M = pixels (n,y,x), N = cout, K = taps * cin over concat(src0, src1).
// 64 x BN block tile, 16-deep K slices staged in shared memory, 4 x (BN/16) register micro-tile.
// ------------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(256)
conv_igemm_fp32_kernel(const float* __restrict__ src0, int C0, const float* __restrict__ src1, int C1,
                       const float* __restrict__ wgt /* [k*k*(C0+C1)][cout] */, const float* __restrict__ bias,
                       const float* __restrict__ res, float* __restrict__ dst, int n_img, int H, int W, int ksz,
                       int cout, int act, float alpha) {
  constexpr int BM = 64, BK = 16, TN = BN / 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int cin = C0 + C1;
  const int K = ksz * ksz * cin;
  const int pad = ksz / 2;
  const long long M = (long long)n_img * H * W;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // ty -> 4 pixels, tx -> TN couts

  // A-load mapping: 16 consecutive k per pixel row (channels are contiguous in NHWC)
  const int a_k = tid & 15;
  const int a_m = tid >> 4;  // 0..15, + 16*j
  int ay[4], ax[4];
  long long aimg[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    long long m = m0 + a_m + 16 * j;
    if (m < M) {
      long long img = m / (H * W);
      int rem = (int)(m - img * (H * W));
      aimg[j] = img;
      ay[j] = rem / W;
      ax[j] = rem - ay[j] * W;
    } else {
      aimg[j] = -1;
      ay[j] = ax[j] = 0;
    }
  }
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // stage A
    {
      int k = k0 + a_k;
      int tap = 0, ci = 0;
      bool kval = k < K;
      if (kval) {
        tap = k / cin;
        ci = k - tap * cin;
      }
      int dy = tap / ksz - pad, dx = tap % ksz - pad;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = 0.0f;
        if (kval && aimg[j] >= 0) {
          int y = ay[j] + dy, x = ax[j] + dx;
          if (y >= 0 && y < H && x >= 0 && x < W) {
            size_t pix = ((size_t)aimg[j] * H + y) * W + x;
            v = ci < C0 ? __ldg(src0 + pix * C0 + ci) : __ldg(src1 + pix * C1 + (ci - C0));
          }
        }
        As[a_k][a_m + 16 * j] = v;
      }
    }
    // stage B: BK x BN, coalesced over cout
    for (int i = tid; i < BK * BN; i += 256) {
      int kk = i / BN, nn = i - kk * BN;
      int k = k0 + kk, n = n0 + nn;
      Bs[kk][nn] = (k < K && n < cout) ? __ldg(wgt + (size_t)k * cout + n) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[TN];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx + 16 * j;
      if (n >= cout) continue;
      float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.0f);
      if (res) v += __ldg(res + (size_t)m * cout + n);
      dst[(size_t)m * cout + n] = apply_act(v, act, alpha);
    }
  }
}

// Transposed convolution with kernel == stride == k: every output pixel depends on exactly one input pixel.
__global__ void convt_fp32_kernel(const float* __restrict__ src, const float* __restrict__ wgt /* [k][k][cin][cout] */,
                                  const float* __restrict__ bias, float* __restrict__ dst, int n_img, int Hin, int Win,
                                  int cin, int cout, int k, int act, float alpha) {
  const int Hout = Hin * k, Wout = Win * k;
  const size_t total = (size_t)n_img * Hout * Wout * cout;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int co = (int)(idx % cout);
    size_t pix = idx / cout;
    int X = (int)(pix % Wout);
    size_t t2 = pix / Wout;
    int Y = (int)(t2 % Hout);
    size_t img = t2 / Hout;
    int yi = Y / k, ky = Y - yi * k, xi = X / k, kx = X - xi * k;
    const float* in = src + ((img * Hin + yi) * Win + xi) * cin;
    const float* w = wgt + ((size_t)(ky * k + kx) * cin) * cout + co;
    float acc = 0.0f;
    for (int ci = 0; ci < cin; ++ci) acc = fmaf(__ldg(in + ci), __ldg(w + (size_t)ci * cout), acc);
    acc += bias ? __ldg(bias + co) : 0.0f;
    dst[idx] = apply_act(acc, act, alpha);
  }
}

__global__ void pool_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_img, int Hin, int Win, int C,
                            int k, int mode, int aux) {
  const int Hout = Hin / k, Wout = Win / k;
  const size_t total = (size_t)n_img * Hout * Wout * C;
  const float inv = 1.0f / (float)(k * k);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    size_t pix = idx / C;
    int X = (int)(pix % Wout);
    size_t t2 = pix / Wout;
    int Y = (int)(t2 % Hout);
    size_t img = t2 / Hout;
    if (mode == FSR_POOL_PICK) {
      dst[idx] = __ldg(src + ((img * Hin + (size_t)Y * k + aux) * Win + (size_t)X * k + aux) * C + c);
      continue;
    }
    float acc = mode == 0 ? -INFINITY : 0.0f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        float v = __ldg(src + ((img * Hin + (size_t)Y * k + dy) * Win + (size_t)X * k + dx) * C + c);
        acc = mode == 0 ? fmaxf(acc, v) : acc + v;
      }
    dst[idx] = mode == 0 ? acc : acc * inv;
  }
}

__global__ void upsample_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_img, int Hin, int Win, int C,
                                int f, int mode) {
  const int Hout = Hin * f, Wout = Win * f;
  const size_t total = (size_t)n_img * Hout * Wout * C;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % C);
    size_t pix = idx / C;
    int X = (int)(pix % Wout);
    size_t t2 = pix / Wout;
    int Y = (int)(t2 % Hout);
    size_t img = t2 / Hout;
    if (mode == FSR_UP_NEAREST) {
      dst[idx] = __ldg(src + ((img * Hin + Y / f) * Win + X / f) * C + c);
      continue;
    }
    int y0, y1, x0, x1;
    float wy, wx;
    up_linear_coord(Y, f, Hin, mode, y0, y1, wy);
    up_linear_coord(X, f, Win, mode, x0, x1, wx);
    const float* im = src + img * Hin * Win * C + c;
    const float a = __ldg(im + ((size_t)y0 * Win + x0) * C), b = __ldg(im + ((size_t)y0 * Win + x1) * C);
    const float d = __ldg(im + ((size_t)y1 * Win + x0) * C), e = __ldg(im + ((size_t)y1 * Win + x1) * C);
    const float top = a + (b - a) * wx, bot = d + (e - d) * wx;
    dst[idx] = top + (bot - top) * wy;
  }
}

__global__ void eltwise_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst,
                               size_t total, int act, float alpha, float beta) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    float v = __ldg(a + idx) + (b ? __ldg(b + idx) : 0.0f);
    dst[idx] = apply_act2(v, act, alpha, beta);
  }
}

// Head tail: 1x1 conv to one linear channel over cmid features per pixel.
__global__ void head_1x1_kernel(const float* __restrict__ feat, const float* __restrict__ w2, const float* __restrict__ b2,
                                float* __restrict__ out, size_t n_pix, int cmid) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (size_t)gridDim.x * blockDim.x) {
    const float* f = feat + p * cmid;
    float acc = 0.0f;
    for (int c = 0; c < cmid; ++c) acc = fmaf(__ldg(f + c), __ldg(w2 + c), acc);
    out[p] = acc + (b2 ? __ldg(b2) : 0.0f);
  }
}

}  // namespace

// invert_depth_log1p_np (preprocessing.py:154-164): clip to [0,1], expm1(y*log1p(D)), clip to [0,D].
__global__ void invert_depth_kernel(const float* __restrict__ pred_norm, float* __restrict__ pred_m, size_t n,
                                    float max_depth, float denom) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float y = fminf(fmaxf(__ldg(pred_norm + i), 0.0f), 1.0f);
    float v = expm1f(__fmul_rn(y, denom));
    pred_m[i] = fminf(fmaxf(v, 0.0f), max_depth);
  }
}

static inline int grid_for(size_t total, int threads = 256) {
  size_t b = (total + threads - 1) / threads;
  const size_t cap = (size_t)current_sm_count() * 32;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

void launch_conv_fp32(const float* src0, int C0, const float* src1, int C1, const float* w, const float* bias,
                      const float* res, float* dst, int n_img, int H, int W, int k, int cout, int act, float alpha,
                      cudaStream_t s) {
  long long M = (long long)n_img * H * W;
  if (cout % 64 == 0) {
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)(cout / 64));
    conv_igemm_fp32_kernel<64><<<grid, 256, 0, s>>>(src0, C0, src1, C1, w, bias, res, dst, n_img, H, W, k, cout, act, alpha);
  } else {
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((cout + 31) / 32));
    conv_igemm_fp32_kernel<32><<<grid, 256, 0, s>>>(src0, C0, src1, C1, w, bias, res, dst, n_img, H, W, k, cout, act, alpha);
  }
  FSR_LAUNCH_CHECK();
}

void launch_convt_fp32(const float* src, const float* w, const float* bias, float* dst, int n_img, int Hin, int Win,
                       int cin, int cout, int k, int act, float alpha, cudaStream_t s) {
  size_t total = (size_t)n_img * Hin * k * Win * k * cout;
  convt_fp32_kernel<<<grid_for(total), 256, 0, s>>>(src, w, bias, dst, n_img, Hin, Win, cin, cout, k, act, alpha);
  FSR_LAUNCH_CHECK();
}

void launch_pool_fp32(const float* src, float* dst, int n_img, int Hin, int Win, int C, int k, int mode, int aux, cudaStream_t s) {
  size_t total = (size_t)n_img * (Hin / k) * (Win / k) * C;
  pool_kernel<<<grid_for(total), 256, 0, s>>>(src, dst, n_img, Hin, Win, C, k, mode, aux);
  FSR_LAUNCH_CHECK();
}

void launch_upsample_fp32(const float* src, float* dst, int n_img, int Hin, int Win, int C, int f, int mode, cudaStream_t s) {
  size_t total = (size_t)n_img * Hin * f * Win * f * C;
  upsample_kernel<<<grid_for(total), 256, 0, s>>>(src, dst, n_img, Hin, Win, C, f, mode);
  FSR_LAUNCH_CHECK();
}

void launch_eltwise_fp32(const float* a, const float* b, float* dst, size_t total, int act, float alpha, float beta, cudaStream_t s) {
  eltwise_kernel<<<grid_for(total), 256, 0, s>>>(a, b, dst, total, act, alpha, beta);
  FSR_LAUNCH_CHECK();
}

void launch_head_1x1_fp32(const float* feat, const float* w2, const float* b2, float* out, size_t n_pix, int cmid,
                          cudaStream_t s) {
  head_1x1_kernel<<<grid_for(n_pix), 256, 0, s>>>(feat, w2, b2, out, n_pix, cmid);
  FSR_LAUNCH_CHECK();
}

void launch_invert_depth(const float* pred_norm, float* pred_m, size_t n, float max_depth, float denom, cudaStream_t s) {
  invert_depth_kernel<<<grid_for(n), 256, 0, s>>>(pred_norm, pred_m, n, max_depth, denom);
  FSR_LAUNCH_CHECK();
}

}  // namespace fsr
