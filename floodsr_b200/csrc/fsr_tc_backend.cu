// Tensor-core backend of the Engine (every precision mode but the SIMT diagnostic one): tensor formats, weight packing and
// op dispatch onto the tcgen05 kernels.  FSR_PREC_FP16 stores activations and weights as one 16-bit tensor;
// FSR_PREC_FP32 (the <= 1e-4 m mode) stores every one of them as a split fp16 pair (hi, lo) and runs three MMAs per product
// (tc_common.cuh: "parts" == 2).
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <numeric>

#include "fsr_engine.cuh"

namespace fsr {

// launchers implemented in k_tc_conv.cu / k_tc_head.cu
int conv_tc_bn(int cout, int parts, int hh_steps);
void launch_conv_tc(const __nv_bfloat16* src0, int C0, long long plane0, const __nv_bfloat16* src1, int C1, long long plane1,
                    const __nv_bfloat16* wpack, int kc, const float* bias, const __nv_bfloat16* res, __nv_bfloat16* dst,
                    long long plane_out, int n_img, int H, int W, int ksz, int cout, int act, float alpha, int half, int im,
                    int parts, float out_scale, int cpad_out, unsigned* flags, cudaStream_t s, int up0 = 0);
bool conv_rows_ok(int H, int W, int ksz, int cout, int C0, int C1, int kc, int parts);
void launch_conv_rows_tc(const __nv_bfloat16* src0, int C0, long long plane0, const __nv_bfloat16* src1, int C1, long long plane1,
                         const __nv_bfloat16* wpack, int kc, const float* bias, const __nv_bfloat16* res, __nv_bfloat16* dst,
                         long long plane_out, int n_img, int H, int W, int cout, int act, float alpha, int half, int n_sms,
                         int parts, float out_scale, unsigned* flags, cudaStream_t s);
void launch_pack_small(const float* s0, int c0, const float* s1, int c1, __nv_bfloat16* dst, long long n_pix, long long plane,
                       int chunks, int half, int parts, cudaStream_t s);
void launch_pool_cp8(const __nv_bfloat16* src, __nv_bfloat16* dst, int chunks, int n_img, int Hin, int Win, int k, int mode,
                     long long plane_in, long long plane_out, int half, int im_in, int im_out, int parts, int aux, cudaStream_t s);
void launch_upsample_cp8(const __nv_bfloat16* src, __nv_bfloat16* dst, int chunks, int n_img, int Hin, int Win, int f,
                         long long plane_in, long long plane_out, int im_in, int im_out, int mode, int half, int parts, cudaStream_t s);
void launch_eltwise_cp8(const __nv_bfloat16* a, const __nv_bfloat16* b, __nv_bfloat16* dst, long long n_vec, int act, float alpha,
                        float beta, int half, long long lo_vec, cudaStream_t s);
void launch_cp8_to_nhwc(const __nv_bfloat16* src, float* dst, long long n_img, int H, int W, long long plane, int C, int half, int im,
                        long long lo_off, cudaStream_t s);
void launch_convt_tc(const __nv_bfloat16* src, long long plane_in, const __nv_bfloat16* wpack, const float* bias,
                     __nv_bfloat16* dst, long long plane_out, int n_img, int Hin, int Win, int cin, int cout, int k, int act,
                     float alpha, int half, cudaStream_t s);
size_t head2_pack_elems();
void head2_pack(const float* w, const float* bias, uint16_t* dst, uint16_t (*cvt)(float), float (*back)(uint16_t));
void launch_head2_tc(const __nv_bfloat16* feat, long long plane, const __nv_bfloat16* wpack, const float* w2, const float* b2,
                     const float* dem, float* pred_m, float* pred_norm, int n_img, int H, int W, int cin, int cmid, int ksz,
                     int act, float alpha, float max_depth, float denom, int half, int n_sms, cudaStream_t s);

bool fused_hr_ok(int H, int W, int lr_h, int lr_w, int cin_t, int cout_t, int k_t, int cmid, int ksz);
size_t fused_hw_elems();
size_t fused_wt_elems();
void fused_pack_head(const float* w, const float* bias, uint16_t* dst, uint16_t (*cvt)(float), float (*back)(uint16_t));
void fused_pack_convt(const float* w, uint16_t* dst, uint16_t (*cvt)(float));
void launch_fused_hr_tc(const __nv_bfloat16* lr, long long lr_plane, const __nv_bfloat16* wt_pack, const float* bias_t, int act_t,
                        float alpha_t, const __nv_bfloat16* hw_pack, const float* w2, const float* b2, int act_h, float alpha_h,
                        const float* dem, const DemSource& src, float* pred_m, float* pred_norm, int n_img, int H, float max_depth,
                        float denom, int half, int n_sms, unsigned* flags, cudaStream_t s);

// split-operand (fp32-tolerance) variant of the fused kernel: k_tc_fused_x3.cu
bool fused_x3_ok(int H, int W, int lr_h, int lr_w, int cin_t, int cout_t, int k_t, int cmid, int ksz);
size_t fused_x3_hw_elems();
size_t fused_x3_wt_elems();
void fused_x3_pack_head(const float* w, float scale, uint16_t* dst);
void fused_x3_pack_convt(const float* w, float scale, uint16_t* dst);
void launch_fused_x3(const __nv_bfloat16* lr, long long lr_plane, const __nv_bfloat16* wt_pack, const float* bias_t, float scale_t_inv,
                     int act_t, float alpha_t, const __nv_bfloat16* hw_pack, const float* bias_h, float scale_h_inv, const float* w2,
                     const float* b2, int act_h, float alpha_h, const float* dem, const DemSource& src, float* pred_m, float* pred_norm,
                     int n_img, int H, float max_depth, float denom, int n_sms, unsigned* flags, cudaStream_t s);

static bool g_pack_half = false;  // 16-bit format used while packing weights (set by tc_prepare)
static inline uint16_t f2bf(float f) {
  uint16_t u;
  if (g_pack_half) {
    __half h = __float2half_rn(f);
    memcpy(&u, &h, 2);
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(f);
    memcpy(&u, &h, 2);
  }
  return u;
}

static inline float bf2f(uint16_t u) {
  if (g_pack_half) {
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
  }
  __nv_bfloat16 h;
  memcpy(&h, &u, 2);
  return __bfloat162float(h);
}

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Split mode: weights are multiplied by a power of two so that the largest one lands in [2^9, 2^10): the lo parts of all
// but vanishing weights are then normal fp16 numbers.  The kernels' epilogues multiply by the inverse (exact).
float split_weight_scale(const float* w, size_t n) {
  float m = 0.f;
  for (size_t i = 0; i < n; ++i) m = std::max(m, std::fabs(w[i]));
  if (!(m > 0.f) || !std::isfinite(m)) return 1.0f;
  int e = 0;
  std::frexp(m, &e);  // m = f * 2^e, f in [0.5, 1)
  return std::ldexp(1.0f, 10 - e);
}
static inline uint16_t f2h(float f) {
  __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static inline float h2f(uint16_t u) {
  __half h;
  memcpy(&h, &u, 2);
  return __half2float(h);
}
// (hi, lo) fp16 pair of a scaled weight
void split_weight(float v, uint16_t& hi, uint16_t& lo) {
  hi = f2h(v);
  lo = f2h(v - h2f(hi));
}

// Decide per-tensor storage and pack every conv-like op's weights into the layouts the kernels stream.
void Engine::tc_prepare(const float* w) {
  parts_ = precision_ == FSR_PREC_FP32 ? 2 : 1;
  g_pack_half = true;  // fp16 storage in every tensor-core mode (the kernels keep a bf16 code path; no mode selects it)
  const int nt = (int)tensors_.size();
  tc_fmt_.assign(nt, 0);
  tc_cpad_.assign(nt, 0);
  tc_im_.assign(nt, 0);
  for (int i = 0; i < nt; ++i) {
    tc_fmt_[i] = tensors_[i].c >= 8 ? 1 : 0;  // 1: CP8 bf16, 0: NHWC fp32 (1-channel rasters)
    tc_cpad_[i] = tc_fmt_[i] ? round_up(tensors_[i].c, 16) : tensors_[i].c;
    // maps of <= 64 pixels are stored image-major (IM8): a conv tap is then one contiguous run per channel chunk instead
    // of hundreds of 32-byte TMA box rows
    tc_im_[i] = (tc_fmt_[i] && !big_[i] && tensors_[i].h * tensors_[i].w <= 64 && !getenv("FSR_NO_IM8")) ? 1 : 0;
  }
  tc_ops_.clear();
  tc_ops_.resize(ops_.size());
  // The tensor-core kernels of the high-resolution end cover the transposed-convolution -> head pair at 32 channels (the
  // shape of the model family's published size); any other high-resolution layer set runs as fp32 FMA kernels on the
  // low-resolution result (run_hr_simt): correct for every graph the lowering accepts, not a throughput path.
  hr_simt_ = parts_ == 2;
  if (parts_ == 1) {
    std::vector<int> hr_ops;
    for (size_t i = 0; i < ops_.size(); ++i)
      if (op_hr_[i]) hr_ops.push_back((int)i);
    bool ok = hr_ops.size() == 2 && ops_[hr_ops[0]].kind == FSR_OP_CONVT && ops_[hr_ops[1]].kind == FSR_OP_HEAD &&
              ops_[hr_ops[1]].src0 == ops_[hr_ops[0]].dst;
    if (ok) {
      const fsr_op& ct = ops_[hr_ops[0]];
      const fsr_op& hd = ops_[hr_ops[1]];
      const int cin = tensors_[ct.src0].c;
      ok = tc_fmt_[ct.src0] && cin % 16 == 0 && cin <= 64 && ct.cout == 32 && ct.k % (256 / ct.cout) == 0 && hd.cout == 32 && hd.k == 3 &&
           hd.src1 >= 0 && !tc_fmt_[hd.src1] && tensors_[hd.src1].c == 1;
    }
    hr_simt_ = !ok;
  }
  // 2x nearest upsampling folded into the convolution that consumes it (conv_tc_kernel, ConvTcParams::up0): image-major
  // levels only (maps of <= 64 pixels), the upsampled tensor must have no other reader.  FSR_NO_FOLD_UP=1 keeps the two ops.
  if (!getenv("FSR_NO_FOLD_UP"))
    for (size_t ui = 0; ui < ops_.size(); ++ui) {
      const fsr_op& up = ops_[ui];
      if (up.kind != FSR_OP_UPSAMPLE || up.mode != FSR_UP_NEAREST || up.k != 2 || op_hr_[ui]) continue;
      if (!tc_fmt_[up.src0] || !tc_im_[up.src0] || !tc_im_[up.dst] || up.dst == hdr_.out_tensor) continue;
      int reader = -1, n_readers = 0;
      for (size_t oi = 0; oi < ops_.size(); ++oi) {
        const fsr_op& o = ops_[oi];
        if (o.src0 == up.dst || o.src1 == up.dst || o.res == up.dst) {
          ++n_readers;
          reader = (int)oi;
        }
      }
      if (n_readers != 1) continue;
      const fsr_op& cv = ops_[reader];
      if (cv.kind != FSR_OP_CONV || cv.k != 3 || cv.src0 != up.dst || cv.src1 < 0 || cv.src1 == up.dst || cv.res == up.dst ||
          !tc_fmt_[cv.src1] || !tc_im_[cv.dst] || op_hr_[reader])
        continue;
      tc_ops_[ui].folded = true;
      tc_ops_[reader].up_src = up.src0;
    }
  for (size_t oi = 0; oi < ops_.size(); ++oi) {
    const fsr_op& op = ops_[oi];
    TcOp& t = tc_ops_[oi];
    if (hr_simt_ && op_hr_[oi]) continue;  // runs on the fp32 kernels (or, in split mode, in the fused split kernel)
    auto need_cp8 = [&](int id, const char* what) {
      if (id >= 0 && !tc_fmt_[id]) throw Error(FSR_E_UNSUPPORTED, std::string("bf16 backend: ") + what + " needs a >= 8-channel tensor");
    };
    if (op.kind == FSR_OP_CONV) {
      const bool s0_small = !tc_fmt_[op.src0];
      const bool s1_small = op.src1 >= 0 && !tc_fmt_[op.src1];
      int C0, C1;
      if (s0_small && (op.src1 < 0 || s1_small)) {
        t.pack_small = true;  // concat of 1-channel rasters -> one zero-padded 16-channel CP8 tensor
        C0 = 16;
        C1 = 0;
        FSR_REQUIRE(tensors_[op.src0].c + (op.src1 >= 0 ? tensors_[op.src1].c : 0) <= 16, "packed input wider than 16 channels");
      } else {
        need_cp8(op.src0, "conv source");
        need_cp8(op.src1, "conv source");
        C0 = tc_cpad_[op.src0];
        C1 = op.src1 >= 0 ? tc_cpad_[op.src1] : 0;
      }
      need_cp8(op.dst, "conv output");
      need_cp8(op.res, "conv residual");
      FSR_REQUIRE(op.cout % 16 == 0, "tensor-core backend: conv output channels must be a multiple of 16");
      int kc = std::gcd(C0 / 8, C1 ? C1 / 8 : C0 / 8);
      const int kc_max = parts_ == 2 ? 4 : 8;  // split stages carry two parts: same bytes per stage
      while (kc > kc_max || (kc % 2)) {
        if (kc % 2) throw Error(FSR_E_UNSUPPORTED, "bf16 backend: conv input channels must be multiples of 16");
        kc /= 2;
      }
      {
        // wide, shallow levels run the persistent row-box kernel, which works on 32-channel groups
        const auto& d = tensors_[op.dst];
        const int kc_rows = kc > 4 ? 4 : kc;
        if (!getenv("FSR_NO_CONV_ROWS") && !tc_im_[op.dst] && conv_rows_ok(d.h, d.w, op.k, op.cout, C0, C1, kc_rows, parts_)) {
          kc = kc_rows;
          t.rows = true;
        }
      }
      t.kc = kc;
      t.C0 = C0;
      t.C1 = C1;
      const int taps = op.k * op.k;
      const int s0 = (C0 / 8) / kc, s1 = C1 ? (C1 / 8) / kc : 0;
      const bool fold = t.up_src >= 0;
      FSR_REQUIRE(!(fold && (t.rows || t.pack_small)), "folded upsampling is conv_tc_kernel's (image-major levels)");
      const int BN = t.rows ? op.cout : conv_tc_bn(op.cout, parts_, (fold ? 4 * s0 + 9 * s1 : taps * (s0 + s1)) * (kc / 2));
      const int n_tiles = ceil_div(op.cout, BN);
      const int real_c0 = t.pack_small ? tensors_[op.src0].c + (op.src1 >= 0 ? tensors_[op.src1].c : 0) : tensors_[op.src0].c;
      const int real_c1 = t.pack_small ? 0 : (op.src1 >= 0 ? tensors_[op.src1].c : 0);
      const int cin_real = real_c0 + real_c1;
      std::vector<uint16_t> pk((size_t)parts_ * n_tiles * (fold ? 16 * s0 + 9 * s1 : taps * (s0 + s1)) * kc * BN * 8, 0);
      const float* wt = w + op.w_off;  // [tap][cin_real][cout]
      const float wscale = parts_ == 2 ? split_weight_scale(wt, (size_t)taps * cin_real * op.cout) : 1.0f;
      t.out_scale = 1.0f / wscale;
      // element strides of the packed layout: conv_tc_kernel streams [n_tile][tap][stage][part][kc][BN][8], the row-box kernel
      // keeps [part][tap][group][kc][BN][8] resident (one output tile)
      const size_t slice = (size_t)kc * BN * 8;
      const size_t part_stride = t.rows ? (size_t)taps * (s0 + s1) * slice : slice;
      const size_t stage_stride = t.rows ? slice : (size_t)parts_ * slice;
      // folded upsampling: src0's taps that share a pixel of the smaller map are summed per output parity.  Along one axis an
      // even output coordinate reaches source offsets {-1: tap -1; 0: taps 0, +1}, an odd one {0: taps -1, 0; +1: tap +1}.
      if (fold) {
        const size_t fold_elems = (size_t)n_tiles * 16 * s0 * stage_stride;
        auto taps_of = [](int parity, int a, int (&d)[2]) {  // taps (0 .. 2) that land on source pixel a (0 / 1) of the pair
          int n = 0;
          for (int k = 0; k < 3; ++k)
            if ((parity == 0 ? (k == 0 ? 0 : 1) : (k == 2 ? 1 : 0)) == a) d[n++] = k;
          return n;
        };
        for (int nt_i = 0; nt_i < n_tiles; ++nt_i)
          for (int par = 0; par < 4; ++par)
            for (int q = 0; q < 4; ++q) {
              int dys[2], dxs[2];
              const int ny_t = taps_of(par >> 1, q >> 1, dys), nx_t = taps_of(par & 1, q & 1, dxs);
              for (int st = 0; st < s0; ++st)
                for (int j = 0; j < kc; ++j)
                  for (int n = 0; n < BN; ++n)
                    for (int e = 0; e < 8; ++e) {
                      const int co = nt_i * BN + n, ci = (st * kc + j) * 8 + e;
                      if (ci >= real_c0 || co >= op.cout) continue;
                      double v = 0.0;
                      for (int a = 0; a < ny_t; ++a)
                        for (int b = 0; b < nx_t; ++b) v += wt[((size_t)(dys[a] * 3 + dxs[b]) * cin_real + ci) * op.cout + co];
                      const size_t pos = ((((size_t)nt_i * 4 + par) * 4 + q) * s0 + st) * stage_stride + ((size_t)j * BN + n) * 8 + e;
                      if (parts_ == 2) split_weight((float)v * wscale, pk[pos], pk[pos + part_stride]);
                      else pk[pos] = f2bf((float)v);
                    }
            }
        for (int nt_i = 0; nt_i < n_tiles; ++nt_i)
          for (int tap = 0; tap < taps; ++tap)
            for (int st = 0; st < s1; ++st)
              for (int j = 0; j < kc; ++j)
                for (int n = 0; n < BN; ++n)
                  for (int e = 0; e < 8; ++e) {
                    const int co = nt_i * BN + n, c = (st * kc + j) * 8 + e;
                    if (c >= real_c1 || co >= op.cout) continue;
                    const size_t pos = fold_elems + (((size_t)nt_i * taps + tap) * s1 + st) * stage_stride + ((size_t)j * BN + n) * 8 + e;
                    const float v = wt[((size_t)tap * cin_real + real_c0 + c) * op.cout + co];
                    if (parts_ == 2) split_weight(v * wscale, pk[pos], pk[pos + part_stride]);
                    else pk[pos] = f2bf(v);
                  }
      }
      for (int nt_i = 0; nt_i < (fold ? 0 : n_tiles); ++nt_i)
        for (int tap = 0; tap < taps; ++tap)
          for (int st = 0; st < s0 + s1; ++st)
            for (int j = 0; j < kc; ++j)
              for (int n = 0; n < BN; ++n)
                for (int e = 0; e < 8; ++e) {
                  const int co = nt_i * BN + n;
                  int ci;  // index into the concatenated real input channels, -1 = padding
                  if (st < s0) {
                    const int c = (st * kc + j) * 8 + e;
                    ci = c < real_c0 ? c : -1;
                  } else {
                    const int c = ((st - s0) * kc + j) * 8 + e;
                    ci = c < real_c1 ? real_c0 + c : -1;
                  }
                  if (ci < 0 || co >= op.cout) continue;
                  const size_t pos = (((size_t)nt_i * taps + tap) * (s0 + s1) + st) * stage_stride + ((size_t)j * BN + n) * 8 + e;
                  const float v = wt[((size_t)tap * cin_real + ci) * op.cout + co];
                  if (parts_ == 2) split_weight(v * wscale, pk[pos], pk[pos + part_stride]);
                  else pk[pos] = f2bf(v);
                }
      t.wpack.ensure(pk.size() * 2);
      FSR_CUDA(cudaMemcpy(t.wpack.p, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    } else if (op.kind == FSR_OP_CONVT) {
      need_cp8(op.src0, "convT source");
      need_cp8(op.dst, "convT output");
      const int cin = tc_cpad_[op.src0], cin_real = tensors_[op.src0].c, cout = op.cout, k = op.k;
      FSR_REQUIRE(cin % 16 == 0 && cin <= 64 && cout == 32 && k % (256 / cout) == 0,
                  "bf16 backend: unsupported transposed-convolution shape");
      const int kc = cin / 8;
      const int n_tiles = k * k * cout / 256;
      const int kx_per_tile = 256 / cout, tiles_per_ky = k / kx_per_tile;
      std::vector<uint16_t> pk((size_t)n_tiles * kc * 256 * 8, 0);
      const float* wt = w + op.w_off;  // [ky][kx][cin_real][cout]
      size_t pos = 0;
      for (int nt_i = 0; nt_i < n_tiles; ++nt_i) {
        const int ky = nt_i / tiles_per_ky, kx0 = (nt_i % tiles_per_ky) * kx_per_tile;
        for (int j = 0; j < kc; ++j)
          for (int n = 0; n < 256; ++n)
            for (int e = 0; e < 8; ++e, ++pos) {
              const int kx = kx0 + n / cout, co = n % cout, ci = j * 8 + e;
              if (ci < cin_real) pk[pos] = f2bf(wt[(((size_t)ky * k + kx) * cin_real + ci) * cout + co]);
            }
      }
      t.wpack.ensure(pk.size() * 2);
      FSR_CUDA(cudaMemcpy(t.wpack.p, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    } else if (op.kind == FSR_OP_HEAD) {
      need_cp8(op.src0, "head feature source");
      FSR_REQUIRE(op.src1 >= 0 && !tc_fmt_[op.src1] && tensors_[op.src1].c == 1, "bf16 backend: head needs a 1-channel second source");
      const int cf = tensors_[op.src0].c, cmid = op.cout;
      FSR_REQUIRE(cf == 32 && cmid == 32 && op.k == 3, "bf16 backend: head is specialised for 32 -> 32 channels, 3x3");
      const float* wt = w + op.w_off;  // [ky][kx][cf + 1][cmid]
      std::vector<uint16_t> pk(head2_pack_elems(), 0);
      head2_pack(wt, op.b_off >= 0 ? w + op.b_off : nullptr, pk.data(), f2bf, bf2f);
      t.wpack.ensure(pk.size() * 2);
      FSR_CUDA(cudaMemcpy(t.wpack.p, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
      const int cin_real = cf + 1;
      t.h_wdem.resize(9 * cmid);
      for (int tap = 0; tap < 9; ++tap)
        for (int co = 0; co < cmid; ++co) t.h_wdem[tap * cmid + co] = wt[((size_t)tap * cin_real + cf) * cmid + co];
      t.h_bias.assign(cmid, 0.f);
      if (op.b_off >= 0) std::copy(w + op.b_off, w + op.b_off + cmid, t.h_bias.begin());
      t.h_w2.assign(w + op.w2_off, w + op.w2_off + cmid);
      t.h_b2 = op.b2_off >= 0 ? w[op.b2_off] : 0.f;
    } else if (op.kind == FSR_OP_POOL) {
      if (tc_fmt_[op.src0] != tc_fmt_[op.dst]) throw Error(FSR_E_UNSUPPORTED, "bf16 backend: pooling changes tensor format");
    } else {
      need_cp8(op.src0, "layer source");
      need_cp8(op.src1, "layer source");
      need_cp8(op.dst, "layer output");
    }
  }
  tc_prepare_fused(w);
}

// The usual high-resolution tail (16x transposed convolution -> fused head) runs as ONE kernel that never writes the
// feature map to HBM (k_tc_fused.cu); anything else falls back to the convT + head kernel pair.
void Engine::tc_prepare_fused(const float* w) {
  fused_ct_ = fused_hd_ = -1;
  if (getenv("FSR_NO_FUSED_HR")) return;
  std::vector<int> hr_ops;
  for (size_t i = 0; i < ops_.size(); ++i)
    if (op_hr_[i]) hr_ops.push_back((int)i);
  if (hr_ops.size() != 2) return;
  const fsr_op& ct = ops_[hr_ops[0]];
  const fsr_op& hd = ops_[hr_ops[1]];
  if (ct.kind != FSR_OP_CONVT || hd.kind != FSR_OP_HEAD || hd.src0 != ct.dst || hd.src1 != 1) return;
  const auto& tl = tensors_[ct.src0];
  const auto& tf = tensors_[ct.dst];
  if (!tc_fmt_[ct.src0] || tc_cpad_[ct.src0] != 32) return;
  if (parts_ == 2) {
    if (getenv("FSR_X3_HR_SIMT") || !fused_x3_ok(tf.h, tf.w, tl.h, tl.w, tl.c, ct.cout, ct.k, hd.cout, hd.k)) return;
    const float* wh = w + hd.w_off;  // [3][3][33][32]
    const float* wc = w + ct.w_off;  // [16][16][32][32]
    const float sh = split_weight_scale(wh, (size_t)hd.k * hd.k * (tf.c + 1) * hd.cout);
    const float st = split_weight_scale(wc, (size_t)ct.k * ct.k * tl.c * ct.cout);
    std::vector<uint16_t> hw(fused_x3_hw_elems(), 0), wt(fused_x3_wt_elems(), 0);
    fused_x3_pack_head(wh, sh, hw.data());
    fused_x3_pack_convt(wc, st, wt.data());
    fused_hw_.ensure(hw.size() * 2);
    fused_wt_.ensure(wt.size() * 2);
    FSR_CUDA(cudaMemcpy(fused_hw_.p, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    FSR_CUDA(cudaMemcpy(fused_wt_.p, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice));
    fused_bias_t_.assign(ct.cout, 0.f);
    if (ct.b_off >= 0) std::copy(w + ct.b_off, w + ct.b_off + ct.cout, fused_bias_t_.begin());
    fused_bias_h_.assign(hd.cout, 0.f);
    if (hd.b_off >= 0) std::copy(w + hd.b_off, w + hd.b_off + hd.cout, fused_bias_h_.begin());
    fused_scale_t_inv_ = 1.0f / st;
    fused_scale_h_inv_ = 1.0f / sh;
    TcOp& th = tc_ops_[hr_ops[1]];
    th.h_w2.assign(w + hd.w2_off, w + hd.w2_off + hd.cout);
    th.h_b2 = hd.b2_off >= 0 ? w[hd.b2_off] : 0.f;
    fused_ct_ = hr_ops[0];
    fused_hd_ = hr_ops[1];
    return;
  }
  if (!fused_hr_ok(tf.h, tf.w, tl.h, tl.w, tl.c, ct.cout, ct.k, hd.cout, hd.k)) return;
  std::vector<uint16_t> hw(fused_hw_elems(), 0), wt(fused_wt_elems(), 0);
  fused_pack_head(w + hd.w_off, hd.b_off >= 0 ? w + hd.b_off : nullptr, hw.data(), f2bf, bf2f);
  fused_pack_convt(w + ct.w_off, wt.data(), f2bf);
  fused_hw_.ensure(hw.size() * 2);
  fused_wt_.ensure(wt.size() * 2);
  FSR_CUDA(cudaMemcpy(fused_hw_.p, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
  FSR_CUDA(cudaMemcpy(fused_wt_.p, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice));
  fused_bias_t_.assign(ct.cout, 0.f);
  if (ct.b_off >= 0) std::copy(w + ct.b_off, w + ct.b_off + ct.cout, fused_bias_t_.begin());
  fused_ct_ = hr_ops[0];
  fused_hd_ = hr_ops[1];
}

// Windows [sub0, sub0 + n) of the batch whose low-resolution result (and normalised inputs) are resident.
void Engine::tc_run_fused(int n, float* d_pred_m, float max_depth, float denom, cudaStream_t s, int sub0) {
  const fsr_op& ct = ops_[fused_ct_];
  const fsr_op& hd = ops_[fused_hd_];
  const TcOp& th = tc_ops_[fused_hd_];
  const auto& tf = tensors_[ct.dst];
  const auto& tl = tensors_[ct.src0];
  const int half = 1;
  const size_t hr_px = (size_t)tf.h * tf.w;
  float* pn = tbase_[hd.dst] ? tbase_[hd.dst] + (size_t)sub0 * hr_px : nullptr;  // nullptr: the caller does not want the normalised prediction
  float* pm = d_pred_m;
  if (!pm) {
    d_tmp_b.ensure((size_t)n * hr_px * sizeof(float));
    pm = d_tmp_b.as<float>();
  }
  // image sub0 of the CP8 tensor L (planes are cap_tiles_ images apart, which the tensor map keeps)
  const __nv_bfloat16* lr = reinterpret_cast<const __nv_bfloat16*>(tbase_[ct.src0]) + (size_t)sub0 * tl.h * tl.w * 8;
  const float* dem = tbase_[1] ? tbase_[1] + (size_t)sub0 * hr_px : nullptr;
  DemSource src = dem_src_;
  if (src.on) {
    src.origins += sub0;
    src.stats += (size_t)sub0 * 3;
  }
  ProfScope scope(prof, PROF_HEAD, s);
  if (parts_ == 2) {
    launch_fused_x3(lr, tc_plane(ct.src0), fused_wt_.as<__nv_bfloat16>(), fused_bias_t_.data(), fused_scale_t_inv_, ct.act, ct.alpha,
                    fused_hw_.as<__nv_bfloat16>(), fused_bias_h_.data(), fused_scale_h_inv_, th.h_w2.data(), &th.h_b2, hd.act, hd.alpha, dem,
                    src, pm, pn, n, tf.h, max_depth, denom, n_sms_, d_flags(), s);
    return;
  }
  launch_fused_hr_tc(lr, tc_plane(ct.src0), fused_wt_.as<__nv_bfloat16>(), fused_bias_t_.data(), ct.act, ct.alpha,
                     fused_hw_.as<__nv_bfloat16>(), th.h_w2.data(), &th.h_b2, hd.act, hd.alpha, dem, src, pm, pn, n, tf.h, max_depth, denom,
                     half, n_sms_, d_flags(), s);
}

void Engine::tc_ensure_arena(int cap) {
  for (size_t i = 0; i < tensors_.size(); ++i) {
    if ((int)i == 0 || (int)i == 1 || (int)i == hdr_.out_tensor) continue;
    if (fused_ct_ >= 0 && (int)i == ops_[fused_ct_].dst) continue;  // the feature map stays on chip
    const auto& t = tensors_[i];
    const size_t tiles = big_[i] ? hr_sub_ : cap;
    const size_t elt = tc_fmt_[i] ? 2 * parts_ : 4;  // split tensors: hi tensor + lo tensor
    if (hr_simt_ && big_[i]) continue;  // high-resolution feature maps of the fp32 pair live in simt_buf_
    tbuf_[i].ensure((size_t)t.h * t.w * tc_cpad_[i] * elt * tiles);
  }
  for (size_t oi = 0; oi < ops_.size(); ++oi)
    if (tc_ops_[oi].pack_small) {
      const auto& d = tensors_[ops_[oi].dst];
      tc_ops_[oi].packbuf.ensure((size_t)d.h * d.w * 16 * 2 * parts_ * cap);
    }
}

// pixels per CP8 plane of tensor `tid` as allocated (capacity, not the live batch)
long long Engine::tc_plane(int tid) const {
  const auto& t = tensors_[tid];
  return (long long)(big_[tid] ? hr_sub_ : cap_tiles_) * t.h * t.w;
}

void Engine::tc_run_ops(bool hr_phase, int n, int sub_start, float* d_pred_m, float max_depth, float denom, cudaStream_t s) {
  for (size_t i = 0; i < ops_.size(); ++i) {
    if ((op_hr_[i] != 0) != hr_phase || (int)i == skip_op_) continue;
    tc_run_one((int)i, n, sub_start, d_pred_m, max_depth, denom, s, n_sms_);
  }
}

// High-resolution phase of one chunk.  When it is the usual pair (transposed convolution -> fused head) the two kernels of
// consecutive sub-chunks overlap: the convT of sub-chunk k+1 (HBM-write bound) runs on its own stream next to the head of
// sub-chunk k (tensor / shared-memory bound), the head's persistent grid being capped so that some SMs are left for it.
// The feature map is double-buffered; events order producer and consumer.
void Engine::tc_run_hr_phase(int n, float* d_pred_m, float max_depth, float denom, cudaStream_t s) {
  if (fused_ct_ >= 0 && skip_op_ != fused_ct_ && skip_op_ != fused_hd_) {
    tc_run_fused(n, d_pred_m, max_depth, denom, s);
    return;
  }
  if (hr_simt_) {
    run_hr_simt(n, d_pred_m, max_depth, denom, s);
    return;
  }
  std::vector<int> hr_ops;
  for (size_t i = 0; i < ops_.size(); ++i)
    if (op_hr_[i] && (int)i != skip_op_) hr_ops.push_back((int)i);
  const bool pair = hr_ops.size() == 2 && ops_[hr_ops[0]].kind == FSR_OP_CONVT && ops_[hr_ops[1]].kind == FSR_OP_HEAD &&
                    ops_[hr_ops[1]].src0 == ops_[hr_ops[0]].dst;
  if (!hr_overlap_ || !pair || n <= hr_sub_) {
    for (int sub = 0; sub < n; sub += hr_sub_) tc_run_ops(true, std::min(hr_sub_, n - sub), sub, d_pred_m, max_depth, denom, s);
    return;
  }
  const int ct = hr_ops[0], hd = hr_ops[1], ftid = ops_[ct].dst;
  if (!s_hd_) {
    int lo = 0, hi = 0;
    FSR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = greatest priority (numerically lowest)
    FSR_CUDA(cudaStreamCreateWithPriority(&s_hd_, cudaStreamNonBlocking, hi));
    FSR_CUDA(cudaStreamCreateWithPriority(&s_ct_, cudaStreamNonBlocking, lo));
    for (auto& e : ev_ring_) FSR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  d_big2_.ensure(tbuf_[ftid].bytes);
  float* bufs[2] = {tbuf_[ftid].as<float>(), d_big2_.as<float>()};
  cudaEvent_t ev_start = ev_ring_[8];
  FSR_CUDA(cudaEventRecord(ev_start, s));
  FSR_CUDA(cudaStreamWaitEvent(s_ct_, ev_start, 0));
  FSR_CUDA(cudaStreamWaitEvent(s_hd_, ev_start, 0));
  int k = 0;
  for (int sub = 0; sub < n; sub += hr_sub_, ++k) {
    const int m = std::min(hr_sub_, n - sub);
    tbase_[ftid] = bufs[k & 1];
    if (k >= 2) FSR_CUDA(cudaStreamWaitEvent(s_ct_, ev_ring_[4 + ((k - 2) & 3)], 0));  // the head that read this buffer is done
    tc_run_one(ct, m, sub, d_pred_m, max_depth, denom, s_ct_, n_sms_);
    FSR_CUDA(cudaEventRecord(ev_ring_[k & 3], s_ct_));
    FSR_CUDA(cudaStreamWaitEvent(s_hd_, ev_ring_[k & 3], 0));
    tc_run_one(hd, m, sub, d_pred_m, max_depth, denom, s_hd_, head_sms_);
    FSR_CUDA(cudaEventRecord(ev_ring_[4 + (k & 3)], s_hd_));
  }
  FSR_CUDA(cudaStreamWaitEvent(s, ev_ring_[4 + ((k - 1) & 3)], 0));
  tbase_[ftid] = bufs[0];
}

// High-resolution layers the tensor-core kernels do not cover (widths other than 32, extra element-wise layers; in split mode also
// FSR_X3_HR_SIMT=1, the A/B switch of the parity tests): the low-resolution result is converted to fp32 NHWC and the layers run
// as plain fp32 FMA kernels (k_fp32.cu), a few tiles at a time.  Correct for every graph the lowering accepts, far from the
// tensor-core roofline.
void Engine::run_hr_simt(int n, float* d_pred_m, float max_depth, float denom, cudaStream_t s) {
  const int sub_tiles = 4;
  simt_buf_.resize(tensors_.size());
  std::vector<float*> saved = tbase_;
  const size_t hr_px = (size_t)hdr_.hr_tile * hdr_.hr_tile;
  size_t headmid = 0;
  for (size_t i = 0; i < ops_.size(); ++i) {
    if (!op_hr_[i] || (int)i == skip_op_) continue;
    const fsr_op& op = ops_[i];
    for (int tid : {op.src0, op.src1, op.res}) {
      if (tid < 0 || big_[tid] || !tc_fmt_[tid] || tbase_[tid] != saved[tid]) continue;
      const auto& t = tensors_[tid];  // low-resolution CP8 pair -> fp32 NHWC
      simt_buf_[tid].ensure((size_t)cap_tiles_ * t.h * t.w * t.c * sizeof(float));
      ProfScope scope(prof, PROF_LR_MISC, s);
      launch_cp8_to_nhwc(reinterpret_cast<const __nv_bfloat16*>(saved[tid]), simt_buf_[tid].as<float>(), n, t.h, t.w, tc_plane(tid), t.c, 1,
                         tc_im_[tid], parts_ == 2 ? (long long)(tc_cpad_[tid] / 8) * tc_plane(tid) * 8 : 0, s);
      tbase_[tid] = simt_buf_[tid].as<float>();
    }
    if (big_[op.dst] && (int)op.dst != hdr_.out_tensor) {
      const auto& t = tensors_[op.dst];
      simt_buf_[op.dst].ensure((size_t)sub_tiles * t.h * t.w * t.c * sizeof(float));
      tbase_[op.dst] = simt_buf_[op.dst].as<float>();
    }
    if (op.kind == FSR_OP_HEAD) headmid = std::max(headmid, hr_px * op.cout * sizeof(float) * sub_tiles);
  }
  if (headmid) d_headmid_.ensure(headmid);
  float* pn = tbase_[hdr_.out_tensor] ? tbase_[hdr_.out_tensor] : d_pred_norm_.as<float>();
  tbase_[hdr_.out_tensor] = pn;
  for (int sub = 0; sub < n; sub += sub_tiles) run_ops(true, std::min(sub_tiles, n - sub), sub, s);
  if (d_pred_m) {
    ProfScope scope(prof, PROF_INVERT, s);
    launch_invert_depth(pn, d_pred_m, (size_t)n * hr_px, max_depth, denom, s);
  }
  tbase_ = saved;
}

void Engine::tc_run_one(int i, int n, int sub_start, float* d_pred_m, float max_depth, float denom, cudaStream_t s, int head_sms) {
  const float* W = d_weights_.as<float>();
  const int half = 1;
  const int parts = parts_;
  // elements between the hi and the lo tensor of a split CP8 tensor
  auto lo_off = [&](int tid) -> long long { return parts == 2 ? (long long)(tc_cpad_[tid] / 8) * tc_plane(tid) * 8 : 0; };
  auto wp = [&](int off) -> const float* { return off >= 0 ? W + off : nullptr; };
  // element pointer of a tensor for the live sub-batch: CP8 planes keep their capacity stride, so the sub-batch
  // offset is a pixel offset inside every plane
  auto cp8 = [&](int tid) -> __nv_bfloat16* {
    if (tid < 0) return nullptr;
    const auto& t = tensors_[tid];
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(tbase_[tid]);
    return big_[tid] ? base : base + (size_t)sub_start * t.h * t.w * 8;
  };
  auto f32 = [&](int tid) -> float* {
    if (tid < 0 || !tbase_[tid]) return nullptr;
    const auto& t = tensors_[tid];
    return tbase_[tid] + (size_t)sub_start * t.h * t.w * t.c;
  };
  {
    const fsr_op& op = ops_[i];
    TcOp& tc = tc_ops_[i];
    const auto& ts = tensors_[op.src0];
    const auto& td = tensors_[op.dst];
    const int cat = op.kind == FSR_OP_HEAD ? PROF_HEAD : op.kind == FSR_OP_CONVT ? PROF_CONVT
                  : op.kind == FSR_OP_CONV ? PROF_LR_CONV : PROF_LR_MISC;
    ProfScope scope(prof, cat, s);
    switch (op.kind) {
      case FSR_OP_CONV: {
        const __nv_bfloat16* s0;
        const __nv_bfloat16* s1 = nullptr;
        long long pl0, pl1 = 0;
        if (tc.pack_small) {
          __nv_bfloat16* pb = tc.packbuf.as<__nv_bfloat16>() + (size_t)sub_start * td.h * td.w * 8;
          const long long plane = (long long)cap_tiles_ * td.h * td.w;
          launch_pack_small(f32(op.src0), ts.c, f32(op.src1), op.src1 >= 0 ? tensors_[op.src1].c : 0, pb, (long long)n * td.h * td.w,
                            plane, 2, half, parts, s);
          s0 = pb;
          pl0 = plane;
        } else {
          const int t0 = tc.up_src >= 0 ? tc.up_src : op.src0;  // folded upsampling: the level below feeds the convolution directly
          s0 = cp8(t0);
          pl0 = tc_plane(t0);
          if (op.src1 >= 0) {
            s1 = cp8(op.src1);
            pl1 = tc_plane(op.src1);
          }
        }
        if (tc.rows)
          launch_conv_rows_tc(s0, tc.C0, pl0, s1, tc.C1, pl1, tc.wpack.as<__nv_bfloat16>(), tc.kc, wp(op.b_off), cp8(op.res),
                              cp8(op.dst), tc_plane(op.dst), n, td.h, td.w, op.cout, op.act, op.alpha, half, n_sms_, parts, tc.out_scale, d_flags(), s);
        else
          launch_conv_tc(s0, tc.C0, pl0, s1, tc.C1, pl1, tc.wpack.as<__nv_bfloat16>(), tc.kc, wp(op.b_off), cp8(op.res), cp8(op.dst),
                         tc_plane(op.dst), n, td.h, td.w, op.k, op.cout, op.act, op.alpha, half, tc_im_[op.dst], parts, tc.out_scale,
                         tc_cpad_[op.dst], d_flags(), s, tc.up_src >= 0 ? 1 : 0);
        break;
      }
      case FSR_OP_POOL:
        if (tc_fmt_[op.src0])
          launch_pool_cp8(cp8(op.src0), cp8(op.dst), tc_cpad_[op.src0] / 8, n, ts.h, ts.w, op.k, op.mode, tc_plane(op.src0),
                          tc_plane(op.dst), half, tc_im_[op.src0], tc_im_[op.dst], parts, op.aux, s);
        else
          launch_pool_fp32(f32(op.src0), f32(op.dst), n, ts.h, ts.w, ts.c, op.k, op.mode, op.aux, s);
        break;
      case FSR_OP_UPSAMPLE:
        if (tc.folded && !force_folded_) break;  // its only reader takes the level below (TcOp::up_src)
        launch_upsample_cp8(cp8(op.src0), cp8(op.dst), tc_cpad_[op.src0] / 8, n, ts.h, ts.w, op.k, tc_plane(op.src0), tc_plane(op.dst),
                            tc_im_[op.src0], tc_im_[op.dst], op.mode, half, parts, s);
        break;
      case FSR_OP_ELTWISE: {
        // planes are strided by capacity: run plane by plane over the live pixels
        // IM8 planes interleave the images: process the whole allocated plane there
        const long long live = tc_im_[op.dst] ? tc_plane(op.dst) : (long long)n * td.h * td.w;
        for (int c8 = 0; c8 < tc_cpad_[op.dst] / 8; ++c8) {
          const size_t o = (size_t)c8 * tc_plane(op.dst) * 8;
          launch_eltwise_cp8(cp8(op.src0) + o, op.src1 >= 0 ? cp8(op.src1) + o : nullptr, cp8(op.dst) + o, live, op.act, op.alpha, op.beta,
                             half, lo_off(op.dst) / 8, s);
        }
        break;
      }
      case FSR_OP_CONVT:
        if (parts == 2) throw Error(FSR_E_UNSUPPORTED, "split mode runs the high-resolution end in the fused kernel or the fp32 pair");
        launch_convt_tc(cp8(op.src0), tc_plane(op.src0), tc.wpack.as<__nv_bfloat16>(), wp(op.b_off), cp8(op.dst), tc_plane(op.dst), n,
                        ts.h, ts.w, tc_cpad_[op.src0], op.cout, op.k, op.act, op.alpha, half, s);
        break;
      case FSR_OP_HEAD: {
        if (parts == 2) throw Error(FSR_E_UNSUPPORTED, "split mode runs the high-resolution end in the fused kernel or the fp32 pair");
        float* pn = f32(op.dst);  // may be nullptr when the caller does not want the normalised prediction
        float* pm = d_pred_m ? d_pred_m + (size_t)sub_start * td.h * td.w : nullptr;
        if (!pm) {
          // forward-only call: metres go to scratch
          d_tmp_b.ensure((size_t)hr_sub_ * td.h * td.w * sizeof(float));
          pm = d_tmp_b.as<float>();
        }
        launch_head2_tc(cp8(op.src0), tc_plane(op.src0), tc.wpack.as<__nv_bfloat16>(), tc.h_w2.data(), &tc.h_b2, f32(op.src1), pm, pn,
                        n, td.h, td.w, ts.c, op.cout, op.k, op.act, op.alpha, max_depth, denom, half, head_sms, s);
        break;
      }
      default:
        throw Error(FSR_E_UNSUPPORTED, "op kind not implemented in the bf16 backend");
    }
  }
}

void Engine::debug_read_tensor(int tid, int n_tiles, float* d_out, cudaStream_t s) {
  FSR_REQUIRE(tid >= 0 && tid < (int)tensors_.size() && tbase_[tid], "tensor is not materialised");
  if (precision_ != FSR_PREC_FP32_SIMT)
    for (size_t oi = 0; oi < ops_.size(); ++oi)
      if (tc_ops_[oi].folded && ops_[oi].dst == tid) {  // never written by the forward pass: produce it now for the reader
        force_folded_ = true;
        tc_run_one((int)oi, n_tiles, 0, nullptr, 1.f, 1.f, s, n_sms_);
        force_folded_ = false;
      }
  const auto& t = tensors_[tid];
  const long long n_pix = (long long)n_tiles * t.h * t.w;
  if (precision_ != FSR_PREC_FP32_SIMT && tc_fmt_[tid]) {
    launch_cp8_to_nhwc(reinterpret_cast<const __nv_bfloat16*>(tbase_[tid]), d_out, n_tiles, t.h, t.w, tc_plane(tid), t.c,
                       1, tc_im_[tid], parts_ == 2 ? (long long)(tc_cpad_[tid] / 8) * tc_plane(tid) * 8 : 0, s);
  } else {
    FSR_CUDA(cudaMemcpyAsync(d_out, tbase_[tid], (size_t)n_pix * t.c * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
}

}  // namespace fsr
