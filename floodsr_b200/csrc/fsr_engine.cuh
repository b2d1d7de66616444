// Engine object behind the C ABI: owns the lowered plan, device weights, activation arenas and the
// raster-level scratch (tile predictions, window geometry tables).
#pragma once
#include <vector>

#include "fsr_common.cuh"

namespace fsr {

struct BlendGeom {
  const int* y_starts;
  const int* x_starts;
  int ny, nx;
  const float* ramp;     // [T] feather weights, nullptr for the hard method
  const int* y_first;    // [Hpad] first window row covering raster row y
  const int* y_count;    // [Hpad] number of window rows covering it
  const int* x_first;    // [Wpad]
  const int* x_count;
  int T, overlap;
  int H, W;              // cropped output extent
  int vec_ok;            // window x-origins are multiples of 4: float4 path allowed
  int max_cover;         // largest number of window rows / columns covering one coordinate
  // per-coordinate tables for the fast kernel (max_cover <= 2): weight of the first / second covering window at a
  // coordinate ([2][pad], same values edge_weight() returns, 1 for the hard method) -- rows and columns
  const float* wy_tab;
  const float* wx_tab;
  int Hpad, Wpad;
};


// Where the fused high-resolution kernels take the normalised DEM from.  on == 0: `tiles` holds the normalised tiles
// [n][T][T] (stage calls, pre-normalised inputs).  on == 1: the kernel reads every tile's window straight out of the raw
// raster and applies normalize_dem_with_stats_np (floodsr/preprocessing.py:61-94) itself from the tile's statistics, with
// the operation sequence of the normalisation kernel (k_prologue.cu, pass N): the normalised tile never exists in HBM.
struct DemSource {
  int on = 0;
  const float* ras = nullptr;     // raw raster rows (row 0 = first row the launch's windows may read)
  const int2* origins = nullptr;  // [n] window origins (y0, x0) relative to `ras`
  const float* stats = nullptr;   // [n][3] p_clip, dem_min, dem_max
  int H = 0, W = 0;               // valid raster extent (pixels beyond it read as 0)
  int has_nodata = 0;
  float nodata = 0.f, nodata_tol = -1.f;
};

// one pixel of normalize_dem_with_stats_np; (p_clip, dem_min, range, zero_out) from dem_norm_spec()
__device__ __forceinline__ float dem_normalise_px(float x, const DemSource& d, float p_clip, float dem_min, float range_f, bool zero_out) {
  if (d.has_nodata) {
    const bool hit = (x == d.nodata) || (d.nodata_tol >= 0.0f && fabsf(x - d.nodata) <= d.nodata_tol);
    if (hit) x = 0.0f;
  }
  x = fminf(fmaxf(x, 0.0f), p_clip);
  const float n = __fdiv_rn(__fsub_rn(x, dem_min), range_f);
  return zero_out ? 0.0f : fminf(fmaxf(n, 0.0f), 1.0f);
}
__device__ __forceinline__ void dem_norm_spec(const float* stats3, float& p_clip, float& dem_min, float& range_f, bool& zero_out) {
  p_clip = stats3[0];
  dem_min = stats3[1];
  const double range_d = (double)stats3[2] - (double)dem_min;  // python-float subtraction in the reference
  zero_out = !(range_d > 0.0);
  range_f = (float)range_d;
}

struct DeviceBuf {
  void* p = nullptr;
  size_t bytes = 0;
  void ensure(size_t n) {
    if (n <= bytes) return;
    if (p) FSR_CUDA(cudaFree(p));
    p = nullptr;
    bytes = 0;
    FSR_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// kernel launchers (k_prologue.cu, k_fp32.cu, k_blend.cu)
void launch_tile_normalize(const float* d_dem, const float* d_depth, const TileGrid& grid, int tile_base, int n_tiles,
                           int T, int TL, int scale, const fsr_tile_params& p, float* d_dem_norm, float* d_depth_norm,
                           float* d_stats, float* d_dem_lr, unsigned* d_flags, cudaStream_t stream, void* d_ws = nullptr);
size_t tile_normalize_ws_bytes(int n_tiles);  // workspace of the three-kernel path (d_ws; nullptr = one CTA per tile)
void launch_conv_fp32(const float* src0, int C0, const float* src1, int C1, const float* w, const float* bias,
                      const float* res, float* dst, int n_img, int H, int W, int k, int cout, int act, float alpha,
                      cudaStream_t s);
void launch_convt_fp32(const float* src, const float* w, const float* bias, float* dst, int n_img, int Hin, int Win,
                       int cin, int cout, int k, int act, float alpha, cudaStream_t s);
void launch_pool_fp32(const float* src, float* dst, int n_img, int Hin, int Win, int C, int k, int mode, int aux, cudaStream_t s);
void launch_upsample_fp32(const float* src, float* dst, int n_img, int Hin, int Win, int C, int f, int mode, cudaStream_t s);
void launch_eltwise_fp32(const float* a, const float* b, float* dst, size_t total, int act, float alpha, float beta, cudaStream_t s);
void launch_head_1x1_fp32(const float* feat, const float* w2, const float* b2, float* out, size_t n_pix, int cmid,
                          cudaStream_t s);
void launch_invert_depth(const float* pred_norm, float* pred_m, size_t n, float max_depth, float denom, cudaStream_t s);
void launch_blend(const float* d_tiles, int ty0, int ty1, const BlendGeom& g, int row0, int n_rows, const float* d_init,
                  int init_rows, bool finalize, float max_depth, float* d_out, cudaStream_t s, int x0 = 0, int x1 = -1);

void launch_resample_bilinear(const float* d_src, int sh, int sw, float* d_dst, int dh, int dw, const fsr_resample_params& p,
                              cudaStream_t s);

// Window grid of one raster pass, resident on the device.
struct WindowGrid {
  int H = 0, W = 0, T = 0, overlap = 0, method = 0;
  std::vector<int> ys, xs;
  bool vec_ok = true;  // all window x-origins are multiples of 4 pixels
  int max_cover = 1;   // largest number of window rows / columns covering one coordinate
  DeviceBuf d_ys, d_xs, d_ramp, d_yfirst, d_ycount, d_xfirst, d_xcount, d_origins, d_wy_tab, d_wx_tab;
  int Hpad = 0, Wpad = 0;
  bool has_ramp = false;
  void release() {
    d_ys.release(); d_xs.release(); d_ramp.release(); d_yfirst.release(); d_ycount.release();
    d_xfirst.release(); d_xcount.release(); d_origins.release(); d_wy_tab.release(); d_wx_tab.release();
  }
};

// Band geometry remembered between fsr_band_run_dev and fsr_band_finalize_dev.
struct BandState {
  int ty0 = 0, ty1 = 0, row0 = 0, n_rows = 0, halo_out_rows = 0;
  float max_depth = 0.f;
};

// Per-op state of the bf16 tensor-core backend (fsr_tc_backend.cu).
struct TcOp {
  DeviceBuf wpack;         // bf16 weights in the streaming layout of the op's kernel
  DeviceBuf packbuf;       // CP8 staging tensor for convs fed by 1-channel fp32 rasters
  bool pack_small = false;
  bool rows = false;       // runs the persistent row-box conv kernel
  bool folded = false;     // UPSAMPLE: 2x nearest upsampling folded into the convolution that reads it; the op does not run
  int up_src = -1;         // CONV: src0 is read through a folded upsampling from this (smaller) tensor
  int kc = 0, C0 = 0, C1 = 0;
  float out_scale = 1.0f;  // split mode: inverse of the power-of-two factor carried by the packed weights
  std::vector<float> h_wdem, h_bias, h_w2;  // head epilogue constants (passed as kernel parameters)
  float h_b2 = 0.f;
};

class Engine {
 public:
  BandState band;
  Engine(const void* plan, size_t plan_bytes, const float* weights, size_t weights_count, int device, int precision);
  ~Engine();

  int lr_tile() const { return hdr_.lr_tile; }
  int hr_tile() const { return hdr_.hr_tile; }
  int scale() const { return hdr_.scale; }
  int64_t macs_per_tile() const;
  LaunchCounter launches;
  Profiler prof;

  // network forward on normalised inputs resident on the device: [n,lr,lr] + [n,hr,hr] -> [n,hr,hr]
  // d_pred_norm and/or d_pred_m (metres, a11 fused or applied right after) may be nullptr
  void forward(int n_tiles, const float* d_depth_norm, const float* d_dem_norm, float* d_pred_norm, float* d_pred_m,
               float max_depth, float denom, cudaStream_t s);
  void debug_read_tensor(int tid, int n_tiles, float* d_out, cudaStream_t s);
  int n_tensors() const { return (int)tensors_.size(); }
  fsr_tensor_desc tensor_desc(int tid) const { return tensors_[tid]; }

  // a5-a11 for tiles described by `grid` origins [tile_base, tile_base+n): writes metres (and optionally
  // the raw normalised prediction) per tile, stats at d_stats[(tile_base+i)*3]
  void run_tiles_from_grid(const float* d_depth, const float* d_dem, const TileGrid& grid, int tile_base, int n_tiles,
                           const fsr_tile_params& p, float* d_pred_m, float* d_pred_norm, float* d_stats, cudaStream_t s);

  // The same work in two phases, for the copy / compute pipelines: the batched low-resolution layers want many windows per
  // launch (a band of one window row runs them at a third of their throughput), copies want small bands.  group_lr runs
  // a5-a8 and the low-resolution layers for a GROUP of windows (<= one chunk); group_hr then runs the fused high-resolution
  // kernel for windows [sub0, sub0 + m) of that group, so that a band's rows can be blended and copied out while the next
  // band of the same group is still computing.  Per-window results do not depend on how windows are batched.
  bool phases_ok() const { return precision_ != FSR_PREC_FP32_SIMT && fused_ct_ >= 0 && pooled_op_ != fused_ct_ && pooled_op_ != fused_hd_; }
  int chunk_tiles() const { return chunk_tiles_; }
  bool compute_bound() const { return precision_ == FSR_PREC_FP32; }  // kernels, not PCIe copies, bound the host-buffer pipelines
  void group_lr(const float* d_depth, const float* d_dem, const TileGrid& grid, int n_tiles, const fsr_tile_params& p, float* d_stats,
                cudaStream_t s);
  void group_hr(int sub0, int m, float* d_pred_m, const fsr_tile_params& p, cudaStream_t s);

  void setup_windows(int H, int W, int method, int overlap, const int* ys, int ny, const int* xs, int nx,
                     const float* ramp, cudaStream_t s);
  BlendGeom blend_geom() const;

  void set_device() const { FSR_CUDA(cudaSetDevice(device_)); }
  unsigned fetch_flags(cudaStream_t s);
  unsigned* d_flags() const { return d_flags_.as<unsigned>(); }

  WindowGrid win;
  DeviceBuf d_tiles;      // [n_tiles][T*T] per-window predictions in metres
  DeviceBuf d_stats;      // [n_tiles][3]
  DeviceBuf d_in_depth, d_in_dem, d_out;  // staging for the host-buffer entry points
  DeviceBuf d_tmp_a, d_tmp_b;
  DeviceBuf d_norm_ws;    // workspace of the three-kernel normalisation (tile_normalize_ws_bytes)
  // fsr_run_raster pipeline: compute / H2D / D2H streams and the two hand-over buffers for rows shared by
  // consecutive bands
  cudaStream_t s_comp = nullptr, s_in = nullptr, s_out = nullptr;
  DeviceBuf d_halo[2];
  int band_halo_rows[2] = {0, 0};
  void ensure_streams() {
    if (s_comp) return;
    FSR_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
    FSR_CUDA(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    FSR_CUDA(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
  }
  int band_tiles_target() const { return band_tiles_; }
  int group_tiles_target() const { return group_tiles_; }
  // fsr_band_host_begin/_end: the sub-band sharing rows with the previous rank is blended last
  DeviceBuf d_tiles0;
  BandState band0;
  float* band0_out = nullptr;
  int band0_rank_row0 = 0;
  bool band0_pending = false;

 private:
  void ensure_arena(int n_tiles);
  void run_ops(bool hr_phase, int n, int sub_start, cudaStream_t s);
  float* tptr(int tid, int sub_start) const;
  // bf16 tensor-core backend (fsr_tc_backend.cu)
  void tc_prepare(const float* host_weights);
  void tc_ensure_arena(int cap);
  long long tc_plane(int tid) const;
  void tc_run_ops(bool hr_phase, int n, int sub_start, float* d_pred_m, float max_depth, float denom, cudaStream_t s);
  void tc_run_one(int op, int n, int sub_start, float* d_pred_m, float max_depth, float denom, cudaStream_t s, int head_sms);
  void tc_run_hr_phase(int n, float* d_pred_m, float max_depth, float denom, cudaStream_t s);
  void tc_prepare_fused(const float* host_weights);
  void tc_run_fused(int n, float* d_pred_m, float max_depth, float denom, cudaStream_t s, int sub0 = 0);
  int group_n_ = 0;                     // windows of the group whose low-resolution result is resident (group_lr)
  void run_hr_simt(int n, float* d_pred_m, float max_depth, float denom, cudaStream_t s);
  bool force_folded_ = false;           // debug_read_tensor: run a folded upsampling op after all
  bool hr_simt_ = false;                // the high-resolution layers run on the fp32 FMA kernels (shapes the tcgen05 kernels do not cover)
  int parts_ = 1;                       // 2: split fp16 (hi, lo) tensors and weights, three MMAs per product (FSR_PREC_FP32)
  int fused_ct_ = -1, fused_hd_ = -1;   // plan ops run by the fused high-resolution kernel, or -1
  DeviceBuf fused_hw_, fused_wt_;
  std::vector<float> fused_bias_t_, fused_bias_h_;
  float fused_scale_t_inv_ = 1.f, fused_scale_h_inv_ = 1.f;
  std::vector<DeviceBuf> simt_buf_;     // fp32 NHWC tensors of the SIMT high-resolution pair (split mode without the fused kernel)
  // convT / head overlap across sub-chunks (tc_run_hr_phase)
  bool hr_overlap_ = true;
  int head_sms_ = 96;
  cudaStream_t s_hd_ = nullptr, s_ct_ = nullptr;
  cudaEvent_t ev_ring_[9] = {};
  DeviceBuf d_big2_;
  std::vector<char> tc_im_;      // CP8 tensor stored image-major (IM8), maps of <= 64 pixels
  std::vector<char> tc_fmt_;     // 1: CP8 bf16, 0: NHWC fp32
  std::vector<int> tc_cpad_;     // channels as stored
  std::vector<TcOp> tc_ops_;

  fsr_plan_header hdr_{};
  std::vector<fsr_tensor_desc> tensors_;
  std::vector<fsr_op> ops_;
  std::vector<char> big_;        // tensor is an HR feature map (only allocated for hr_sub_ tiles)
  std::vector<char> op_hr_;      // op touches a big tensor
  int device_ = 0, precision_ = 0, n_sms_ = 148;
  int cap_tiles_ = 0;            // arena capacity (tiles per chunk)
  int hr_sub_ = 4;               // tiles per HR sub-chunk (tensor-core modes: 64; env FSR_HR_SUB overrides)
  int chunk_tiles_ = 64;
  int band_tiles_ = 170;         // windows per band of the fsr_run_raster copy/compute pipeline (env FSR_BAND_TILES)
  int group_tiles_ = 340;        // two-phase pipeline: largest group of windows whose low-resolution layers run in one batch (env FSR_GROUP_TILES)
  DeviceBuf d_weights_, d_flags_, d_headmid_;
  std::vector<DeviceBuf> tbuf_;
  std::vector<float*> tbase_;    // per-forward tensor base pointers (inputs/outputs alias caller buffers)
  DeviceBuf d_dem_norm_, d_depth_norm_, d_pred_norm_, d_dem_lr_;
  const float* dem_lr_pre_ = nullptr;  // pooled normalised DEM written by the normalisation kernel for the current batch
  bool no_lazy_dem_ = false;           // FSR_NO_LAZY_DEM at construction: always materialise the normalised DEM tiles
  DemSource dem_src_;                  // current batch: the fused kernel normalises the raw raster windows itself (on == 1)
  bool lazy_dem_ok() const;            // the plan reads dem_hr only through the pooled branch and the fused head
  int skip_op_ = -1;                   // op skipped by the current forward pass
  int pooled_op_ = -1;                 // plan op it replaces (scale x scale average pool of the DEM input), or -1
};

}  // namespace fsr
