// C ABI (include/floodsr_b200.h) and the Engine that executes the lowered plan.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>

#include "fsr_engine.cuh"

namespace fsr {

thread_local LaunchCounter* g_launch_counter = nullptr;
static thread_local std::string g_last_error;

// ------------------------------------------------------------------------------------------------------
// Engine
// ------------------------------------------------------------------------------------------------------

Engine::Engine(const void* plan, size_t plan_bytes, const float* weights, size_t weights_count, int device, int precision)
    : device_(device), precision_(precision) {
  FSR_REQUIRE(plan && plan_bytes >= sizeof(fsr_plan_header), "plan is empty");
  memcpy(&hdr_, plan, sizeof(hdr_));
  FSR_REQUIRE(hdr_.magic == FSR_PLAN_MAGIC && hdr_.version == FSR_PLAN_VERSION, "plan magic/version mismatch");
  FSR_REQUIRE(hdr_.n_tensors >= 3 && hdr_.n_ops >= 1, "plan has no layers");
  size_t need = sizeof(hdr_) + (size_t)hdr_.n_tensors * sizeof(fsr_tensor_desc) + (size_t)hdr_.n_ops * sizeof(fsr_op);
  FSR_REQUIRE(plan_bytes == need, "plan size does not match its header");
  const char* p = (const char*)plan + sizeof(hdr_);
  tensors_.resize(hdr_.n_tensors);
  memcpy(tensors_.data(), p, tensors_.size() * sizeof(fsr_tensor_desc));
  p += tensors_.size() * sizeof(fsr_tensor_desc);
  ops_.resize(hdr_.n_ops);
  memcpy(ops_.data(), p, ops_.size() * sizeof(fsr_op));
  FSR_REQUIRE(hdr_.lr_tile > 0 && hdr_.hr_tile == hdr_.lr_tile * hdr_.scale, "inconsistent tile geometry");
  FSR_REQUIRE(hdr_.hr_tile % 64 == 0, "hr tile must be a multiple of 64");
  FSR_REQUIRE(hdr_.out_tensor >= 2 && hdr_.out_tensor < hdr_.n_tensors, "bad output tensor");
  FSR_REQUIRE(precision != 1, "precision mode 1 (bf16 operands) was retired: it misses the 1e-2 m bound; use FSR_PREC_FP16");
  FSR_REQUIRE(precision == FSR_PREC_FP32 || precision == FSR_PREC_FP16 || precision == FSR_PREC_FP32_SIMT, "unknown precision mode");
  if (precision == FSR_PREC_FP32) band_tiles_ = 255;  // compute-bound mode: larger bands (measured e2e 24.8 -> 23.6 ms at 935 windows)
  if (const char* e = getenv("FSR_BAND_TILES")) band_tiles_ = std::max(1, atoi(e));
  if (precision == FSR_PREC_FP16) group_tiles_ = 170;  // copy-bound mode: a group waits for its last row, keep groups short
  if (const char* e = getenv("FSR_GROUP_TILES")) group_tiles_ = std::max(1, atoi(e));
  no_lazy_dem_ = getenv("FSR_NO_LAZY_DEM") != nullptr;

  const size_t hr_px = (size_t)hdr_.hr_tile * hdr_.hr_tile;
  big_.assign(tensors_.size(), 0);
  for (size_t i = 0; i < tensors_.size(); ++i) {
    const auto& t = tensors_[i];
    FSR_REQUIRE(t.h > 0 && t.w > 0 && t.c > 0, "bad tensor shape in plan");
    big_[i] = ((size_t)t.h * t.w * t.c > hr_px) ? 1 : 0;
  }
  auto tensor_ok = [&](int id, bool allow_none) { return (allow_none && id == -1) || (id >= 0 && id < hdr_.n_tensors); };
  std::vector<int> producer(tensors_.size(), -1);
  op_hr_.assign(ops_.size(), 0);
  for (size_t i = 0; i < ops_.size(); ++i) {
    const fsr_op& op = ops_[i];
    FSR_REQUIRE(op.kind >= FSR_OP_CONV && op.kind <= FSR_OP_HEAD, "unknown op kind in plan");
    FSR_REQUIRE(op.act >= FSR_ACT_NONE && op.act <= FSR_ACT_SIGMOID && (op.kind == FSR_OP_ELTWISE || op.act <= FSR_ACT_LEAKY),
                "activation not supported by this op kind");
    FSR_REQUIRE(op.kind != FSR_OP_POOL || (op.mode >= FSR_POOL_MAX && op.mode <= FSR_POOL_PICK && (op.mode != FSR_POOL_PICK || (op.aux >= 0 && op.aux < op.k))),
                "bad pooling mode");
    FSR_REQUIRE(op.kind != FSR_OP_UPSAMPLE || (op.mode >= FSR_UP_NEAREST && op.mode <= FSR_UP_LINEAR_ASYMMETRIC), "bad upsampling mode");
    FSR_REQUIRE(tensor_ok(op.src0, false) && tensor_ok(op.src1, true) && tensor_ok(op.res, true) && tensor_ok(op.dst, false),
                "op references a tensor outside the plan");
    auto off_ok = [&](int off) { return off == -1 || (off >= 0 && (size_t)off < weights_count); };
    FSR_REQUIRE(off_ok(op.w_off) && off_ok(op.b_off) && off_ok(op.w2_off) && off_ok(op.b2_off), "weight offset outside blob");
    for (int sidx : {op.src0, op.src1, op.res})  // every operand is a graph input or was written by an earlier op
      FSR_REQUIRE(sidx < 0 || sidx <= 1 || producer[sidx] >= 0, "plan op reads a tensor that no earlier op has produced");
    FSR_REQUIRE(op.dst > 1 && producer[op.dst] < 0, "plan op overwrites a graph input or an already produced tensor");
    bool hr = big_[op.dst] || big_[op.src0] || (op.src1 >= 0 && big_[op.src1]) || (op.res >= 0 && big_[op.res]);
    if (op.kind == FSR_OP_HEAD) hr = true;
    op_hr_[i] = hr ? 1 : 0;
    if (!hr) {
      for (int s : {op.src0, op.src1, op.res})
        if (s >= 0 && producer[s] >= 0 && op_hr_[producer[s]])
          throw Error(FSR_E_UNSUPPORTED, "plan has a low-resolution layer that consumes a high-resolution feature map");
    }
    producer[op.dst] = (int)i;
  }

  for (size_t i = 0; i < ops_.size(); ++i) {
    const fsr_op& op = ops_[i];
    if (op.kind == FSR_OP_POOL && op.src0 == 1 && op.mode == 1 && op.k == hdr_.scale && hdr_.scale == 16 && hdr_.hr_tile == 512 &&
        tensors_[op.dst].c == 1) {
      pooled_op_ = (int)i;  // the normalisation kernel produces this tensor on the fly
      break;
    }
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    throw Error(FSR_E_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") + cudaGetErrorString(e));
  FSR_REQUIRE(device >= 0 && device < n_dev, "device index out of range");
  set_device();
  d_weights_.ensure(std::max<size_t>(weights_count, 1) * sizeof(float));
  if (weights_count) FSR_CUDA(cudaMemcpy(d_weights_.p, weights, weights_count * sizeof(float), cudaMemcpyHostToDevice));
  d_flags_.ensure(sizeof(unsigned));
  FSR_CUDA(cudaMemset(d_flags_.p, 0, sizeof(unsigned)));
  tbuf_.resize(tensors_.size());
  tbase_.assign(tensors_.size(), nullptr);
  if (precision_ != FSR_PREC_FP32_SIMT) {
    cudaDeviceProp prop;
    FSR_CUDA(cudaGetDeviceProperties(&prop, device_));
    if (prop.major != 10) throw Error(FSR_E_UNSUPPORTED, "the tensor-core backends need an sm_100 (Blackwell) device: tcgen05/TMEM/TMA");
    n_sms_ = prop.multiProcessorCount;
    chunk_tiles_ = 1024;  // the batched low-resolution layers need many tiles per launch to fill 148 SMs
    hr_sub_ = 64;  // the HR feature map of a sub-chunk streams through HBM; long head launches amortise their prologue
    if (const char* e = getenv("FSR_HR_SUB")) hr_sub_ = std::max(1, atoi(e));
    if (const char* e = getenv("FSR_CHUNK")) chunk_tiles_ = std::max(1, atoi(e));
    if (const char* e = getenv("FSR_HR_OVERLAP")) hr_overlap_ = atoi(e) != 0;
    head_sms_ = std::min(n_sms_, 96);  // measured best split on a 148-SM B200: 96 SMs head, the rest free for convT
    if (const char* e = getenv("FSR_HEAD_SMS")) head_sms_ = std::max(1, std::min(n_sms_, atoi(e)));
    tc_prepare(weights);
  }
}

Engine::~Engine() {
  cudaSetDevice(device_);
  for (auto& b : tbuf_) b.release();
  for (auto& b : simt_buf_) b.release();
  for (DeviceBuf* b : {&d_weights_, &d_flags_, &d_headmid_, &d_dem_norm_, &d_depth_norm_, &d_pred_norm_, &d_dem_lr_, &d_tiles, &d_stats,
                       &d_in_depth, &d_in_dem, &d_out, &d_tmp_a, &d_tmp_b, &d_norm_ws})
    b->release();
  win.release();
  d_big2_.release();
  d_tiles0.release();
  fused_hw_.release();
  fused_wt_.release();
  if (s_hd_) cudaStreamDestroy(s_hd_);
  if (s_ct_) cudaStreamDestroy(s_ct_);
  for (auto e : ev_ring_) if (e) cudaEventDestroy(e);
  d_halo[0].release();
  d_halo[1].release();
  if (s_comp) cudaStreamDestroy(s_comp);
  if (s_in) cudaStreamDestroy(s_in);
  if (s_out) cudaStreamDestroy(s_out);
}

bool Engine::lazy_dem_ok() const {
  // split mode only: there a row of the fused kernel lasts ~4 us and its DEM warp has time for 514 divisions; in the 16-bit
  // kernel (1.4 us per row) the in-kernel normalisation became the critical path (measured: 4.6 -> 8.6 ms)
  if (parts_ != 2 || fused_ct_ < 0 || pooled_op_ < 0) return false;
  for (size_t i = 0; i < ops_.size(); ++i) {
    if ((int)i == pooled_op_ || (int)i == fused_hd_) continue;
    const fsr_op& op = ops_[i];
    if (op.src0 == 1 || op.src1 == 1 || op.res == 1) return false;
  }
  return ops_[fused_hd_].src1 == 1 && ops_[pooled_op_].src0 == 1;
}

int64_t Engine::macs_per_tile() const {
  int64_t total = 0;
  for (const fsr_op& op : ops_) {
    const auto& d = tensors_[op.dst];
    if (op.kind == FSR_OP_CONV || op.kind == FSR_OP_HEAD) {
      int cin = tensors_[op.src0].c + (op.src1 >= 0 ? tensors_[op.src1].c : 0);
      total += (int64_t)d.h * d.w * op.k * op.k * cin * op.cout;
      if (op.kind == FSR_OP_HEAD) total += (int64_t)d.h * d.w * op.cout;
    } else if (op.kind == FSR_OP_CONVT) {
      total += (int64_t)d.h * d.w * tensors_[op.src0].c * op.cout;
    }
  }
  return total;
}

void Engine::ensure_arena(int n_tiles) {
  if (n_tiles <= cap_tiles_) return;
  const int cap = std::min(ceil_div(n_tiles, 64) * 64, std::max(chunk_tiles_, n_tiles));  // grow-only, in steps of 64 tiles
  size_t headmid = 0;
  const size_t hr_px0 = (size_t)hdr_.hr_tile * hdr_.hr_tile, lr_px0 = (size_t)hdr_.lr_tile * hdr_.lr_tile;
  if (precision_ != FSR_PREC_FP32_SIMT) {
    tc_ensure_arena(cap);
    d_dem_norm_.ensure(hr_px0 * sizeof(float) * cap);
    d_pred_norm_.ensure(hr_px0 * sizeof(float) * cap);
    d_depth_norm_.ensure(lr_px0 * sizeof(float) * cap);
    d_dem_lr_.ensure(lr_px0 * sizeof(float) * cap);
    cap_tiles_ = cap;
    return;
  }
  for (size_t i = 0; i < tensors_.size(); ++i) {
    if ((int)i == 0 || (int)i == 1 || (int)i == hdr_.out_tensor) continue;  // alias caller buffers
    const auto& t = tensors_[i];
    size_t per_tile = (size_t)t.h * t.w * t.c * sizeof(float);
    tbuf_[i].ensure(per_tile * (big_[i] ? hr_sub_ : cap));
  }
  for (const fsr_op& op : ops_)
    if (op.kind == FSR_OP_HEAD) {
      const auto& d = tensors_[op.dst];
      headmid = std::max(headmid, (size_t)d.h * d.w * op.cout * sizeof(float) * hr_sub_);
    }
  if (headmid) d_headmid_.ensure(headmid);
  const size_t hr_px = (size_t)hdr_.hr_tile * hdr_.hr_tile, lr_px = (size_t)hdr_.lr_tile * hdr_.lr_tile;
  d_dem_norm_.ensure(hr_px * sizeof(float) * cap);
  d_pred_norm_.ensure(hr_px * sizeof(float) * cap);
  d_depth_norm_.ensure(lr_px * sizeof(float) * cap);
  d_dem_lr_.ensure(lr_px * sizeof(float) * cap);
  cap_tiles_ = cap;
}

float* Engine::tptr(int tid, int sub_start) const {
  if (tid < 0 || !tbase_[tid]) return nullptr;
  const auto& t = tensors_[tid];
  float* base = tbase_[tid];
  if (big_[tid]) return base;
  return base + (size_t)sub_start * t.h * t.w * t.c;
}

void Engine::run_ops(bool hr_phase, int n, int sub_start, cudaStream_t s) {
  const float* W = d_weights_.as<float>();
  auto wp = [&](int off) -> const float* { return off >= 0 ? W + off : nullptr; };
  for (size_t i = 0; i < ops_.size(); ++i) {
    if ((op_hr_[i] != 0) != hr_phase || (int)i == skip_op_) continue;
    const fsr_op& op = ops_[i];
    const auto& ts = tensors_[op.src0];
    const auto& td = tensors_[op.dst];
    const float* s0 = tptr(op.src0, sub_start);
    const float* s1 = tptr(op.src1, sub_start);
    const float* rs = tptr(op.res, sub_start);
    float* dst = tptr(op.dst, sub_start);
    const int c1 = op.src1 >= 0 ? tensors_[op.src1].c : 0;
    const int cat = op.kind == FSR_OP_HEAD ? PROF_HEAD : op.kind == FSR_OP_CONVT ? PROF_CONVT
                  : op.kind == FSR_OP_CONV ? PROF_LR_CONV : PROF_LR_MISC;
    ProfScope scope(prof, cat, s);
    switch (op.kind) {
      case FSR_OP_CONV:
        launch_conv_fp32(s0, ts.c, s1, c1, wp(op.w_off), wp(op.b_off), rs, dst, n, td.h, td.w, op.k, op.cout, op.act, op.alpha, s);
        break;
      case FSR_OP_POOL:
        launch_pool_fp32(s0, dst, n, ts.h, ts.w, ts.c, op.k, op.mode, op.aux, s);
        break;
      case FSR_OP_UPSAMPLE:
        launch_upsample_fp32(s0, dst, n, ts.h, ts.w, ts.c, op.k, op.mode, s);
        break;
      case FSR_OP_CONVT:
        launch_convt_fp32(s0, wp(op.w_off), wp(op.b_off), dst, n, ts.h, ts.w, ts.c, op.cout, op.k, op.act, op.alpha, s);
        break;
      case FSR_OP_ELTWISE:
        launch_eltwise_fp32(s0, s1, dst, (size_t)n * td.h * td.w * td.c, op.act, op.alpha, op.beta, s);
        break;
      case FSR_OP_HEAD: {
        float* mid = d_headmid_.as<float>();
        launch_conv_fp32(s0, ts.c, s1, c1, wp(op.w_off), wp(op.b_off), nullptr, mid, n, td.h, td.w, op.k, op.cout, op.act, op.alpha, s);
        launch_head_1x1_fp32(mid, wp(op.w2_off), wp(op.b2_off), dst, (size_t)n * td.h * td.w, op.cout, s);
        break;
      }
      default:
        throw Error(FSR_E_UNSUPPORTED, "op kind not implemented");
    }
  }
}

void Engine::forward(int n_tiles, const float* d_depth_norm, const float* d_dem_norm, float* d_pred_norm, float* d_pred_m,
                     float max_depth, float denom, cudaStream_t s) {
  if (n_tiles <= 0) return;
  const size_t hr_px = (size_t)hdr_.hr_tile * hdr_.hr_tile, lr_px = (size_t)hdr_.lr_tile * hdr_.lr_tile;
  for (int c0 = 0; c0 < n_tiles; c0 += chunk_tiles_) {
    const int n = std::min(chunk_tiles_, n_tiles - c0);
    ensure_arena(n);
    for (size_t i = 0; i < tensors_.size(); ++i) tbase_[i] = tbuf_[i].as<float>();
    tbase_[0] = const_cast<float*>(d_depth_norm) + (size_t)c0 * lr_px;
    tbase_[1] = d_dem_norm ? const_cast<float*>(d_dem_norm) + (size_t)c0 * hr_px : nullptr;
    skip_op_ = -1;
    if (dem_lr_pre_ && pooled_op_ >= 0) {
      tbase_[ops_[pooled_op_].dst] = const_cast<float*>(dem_lr_pre_) + (size_t)c0 * lr_px;
      skip_op_ = pooled_op_;
    }
    float* pm = d_pred_m ? d_pred_m + (size_t)c0 * hr_px : nullptr;
    if (precision_ != FSR_PREC_FP32_SIMT) {
      tbase_[hdr_.out_tensor] = d_pred_norm ? d_pred_norm + (size_t)c0 * hr_px : nullptr;
      tc_run_ops(false, n, 0, nullptr, max_depth, denom, s);
      tc_run_hr_phase(n, pm, max_depth, denom, s);
      continue;
    }
    // fp32 path always materialises the normalised prediction, then inverts it
    float* pn = d_pred_norm ? d_pred_norm + (size_t)c0 * hr_px : d_pred_norm_.as<float>();
    tbase_[hdr_.out_tensor] = pn;
    run_ops(false, n, 0, s);
    for (int sub = 0; sub < n; sub += hr_sub_) run_ops(true, std::min(hr_sub_, n - sub), sub, s);
    if (pm) {
      ProfScope scope(prof, PROF_INVERT, s);
      launch_invert_depth(pn, pm, (size_t)n * hr_px, max_depth, denom, s);
    }
  }
}

void Engine::run_tiles_from_grid(const float* d_depth, const float* d_dem, const TileGrid& grid, int tile_base, int n_tiles,
                                 const fsr_tile_params& p, float* d_pred_m, float* d_pred_norm, float* d_stats_out,
                                 cudaStream_t s) {
  const int T = hdr_.hr_tile, TL = hdr_.lr_tile;
  const size_t hr_px = (size_t)T * T;
  for (int c0 = 0; c0 < n_tiles; c0 += chunk_tiles_) {
    const int n = std::min(chunk_tiles_, n_tiles - c0);
    ensure_arena(n);
    // The normalised DEM tile is not materialised when its only readers are the pooled low-resolution branch (produced by
    // the normalisation kernel itself) and the fused high-resolution kernel (which then normalises the raster window from
    // the tile's statistics): one 1 MiB write and one 1 MiB read per tile less.  FSR_NO_LAZY_DEM=1 keeps the old path (A/B).
    const bool lazy = p.normalize_inputs && !no_lazy_dem_ && lazy_dem_ok();
    {
      ProfScope scope(prof, PROF_PROLOGUE, s);
      d_norm_ws.ensure(tile_normalize_ws_bytes(cap_tiles_));
      launch_tile_normalize(d_dem, d_depth, grid, tile_base + c0, n, T, TL, hdr_.scale, p, lazy ? nullptr : d_dem_norm_.as<float>(),
                            d_depth_norm_.as<float>(), d_stats_out, d_dem_lr_.as<float>(), d_flags(), s, d_norm_ws.p);
    }
    dem_lr_pre_ = p.normalize_inputs ? d_dem_lr_.as<float>() : nullptr;
    dem_src_ = DemSource{};
    if (lazy) {
      dem_src_.on = 1;
      dem_src_.ras = d_dem;
      dem_src_.origins = grid.origins + tile_base + c0;
      dem_src_.stats = d_stats_out + (size_t)(tile_base + c0) * 3;
      dem_src_.H = grid.H;
      dem_src_.W = grid.W;
      dem_src_.has_nodata = p.has_dem_nodata;
      dem_src_.nodata = p.dem_nodata;
      dem_src_.nodata_tol = p.dem_nodata_tol;
    }
    float* pn = d_pred_norm ? d_pred_norm + (size_t)c0 * hr_px : nullptr;
    forward(n, d_depth_norm_.as<float>(), lazy ? nullptr : d_dem_norm_.as<float>(), pn, d_pred_m + (size_t)c0 * hr_px, p.max_depth,
            p.depth_denom, s);
    dem_lr_pre_ = nullptr;
    dem_src_ = DemSource{};
  }
}

void Engine::group_lr(const float* d_depth, const float* d_dem, const TileGrid& grid, int n, const fsr_tile_params& p, float* d_stats,
                      cudaStream_t s) {
  FSR_REQUIRE(phases_ok(), "the two-phase path needs the fused tensor-core kernels");
  FSR_REQUIRE(n > 0 && n <= chunk_tiles_, "a group is at most one chunk of windows");
  group_n_ = 0;
  ensure_arena(n);
  const bool lazy = p.normalize_inputs && !no_lazy_dem_ && lazy_dem_ok();  // see run_tiles_from_grid
  {
    ProfScope scope(prof, PROF_PROLOGUE, s);
    d_norm_ws.ensure(tile_normalize_ws_bytes(cap_tiles_));
    launch_tile_normalize(d_dem, d_depth, grid, 0, n, hdr_.hr_tile, hdr_.lr_tile, hdr_.scale, p, lazy ? nullptr : d_dem_norm_.as<float>(),
                          d_depth_norm_.as<float>(), d_stats, d_dem_lr_.as<float>(), d_flags(), s, d_norm_ws.p);
  }
  dem_lr_pre_ = p.normalize_inputs ? d_dem_lr_.as<float>() : nullptr;
  dem_src_ = DemSource{};
  if (lazy) {
    dem_src_.on = 1;
    dem_src_.ras = d_dem;
    dem_src_.origins = grid.origins;
    dem_src_.stats = d_stats;
    dem_src_.H = grid.H;
    dem_src_.W = grid.W;
    dem_src_.has_nodata = p.has_dem_nodata;
    dem_src_.nodata = p.dem_nodata;
    dem_src_.nodata_tol = p.dem_nodata_tol;
  }
  for (size_t i = 0; i < tensors_.size(); ++i) tbase_[i] = tbuf_[i].as<float>();
  tbase_[0] = d_depth_norm_.as<float>();
  tbase_[1] = lazy ? nullptr : d_dem_norm_.as<float>();
  tbase_[hdr_.out_tensor] = nullptr;
  skip_op_ = -1;
  if (dem_lr_pre_ && pooled_op_ >= 0) {
    tbase_[ops_[pooled_op_].dst] = const_cast<float*>(dem_lr_pre_);
    skip_op_ = pooled_op_;
  }
  tc_run_ops(false, n, 0, nullptr, p.max_depth, p.depth_denom, s);
  group_n_ = n;
}

void Engine::group_hr(int sub0, int m, float* d_pred_m, const fsr_tile_params& p, cudaStream_t s) {
  FSR_REQUIRE(sub0 >= 0 && m > 0 && sub0 + m <= group_n_, "windows outside the resident group");
  tc_run_fused(m, d_pred_m, p.max_depth, p.depth_denom, s, sub0);
}

unsigned Engine::fetch_flags(cudaStream_t s) {
  unsigned f = 0;
  FSR_CUDA(cudaMemcpyAsync(&f, d_flags_.p, sizeof(f), cudaMemcpyDeviceToHost, s));
  FSR_CUDA(cudaMemsetAsync(d_flags_.p, 0, sizeof(f), s));
  FSR_CUDA(cudaStreamSynchronize(s));
  return f;
}

static void upload_ints(DeviceBuf& b, const std::vector<int>& v, cudaStream_t s) {
  b.ensure(std::max<size_t>(v.size(), 1) * sizeof(int));
  if (!v.empty()) FSR_CUDA(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, s));
}

void Engine::setup_windows(int H, int W, int method, int overlap, const int* ys, int ny, const int* xs, int nx,
                           const float* ramp, cudaStream_t s) {
  const int T = hdr_.hr_tile;
  FSR_REQUIRE(H > 0 && W > 0 && ny > 0 && nx > 0 && ys && xs, "empty raster or window grid");
  FSR_REQUIRE(method == FSR_WINDOW_HARD || method == FSR_WINDOW_FEATHER, "unknown window method");
  FSR_REQUIRE(method == FSR_WINDOW_HARD || ramp != nullptr, "feather windowing needs the ramp");
  FSR_REQUIRE(overlap >= 0 && overlap < T, "overlap must be in [0, tile)");
  const int Hpad = ceil_div(H, T) * T, Wpad = ceil_div(W, T) * T;
  win.H = H; win.W = W; win.T = T; win.overlap = overlap; win.method = method;
  win.ys.assign(ys, ys + ny);
  win.xs.assign(xs, xs + nx);
  win.vec_ok = true;
  auto check_axis = [&](const std::vector<int>& st, int pad, const char* name) {
    for (size_t i = 0; i < st.size(); ++i) {
      FSR_REQUIRE(st[i] >= 0 && st[i] + T <= pad, std::string(name) + " window origin outside the padded raster");
      FSR_REQUIRE(i == 0 || st[i] > st[i - 1], std::string(name) + " window origins must be strictly increasing");
      FSR_REQUIRE(st[i] % hdr_.scale == 0, std::string(name) + " window origin must be a multiple of the model scale");
    }
  };
  check_axis(win.ys, Hpad, "y");
  check_axis(win.xs, Wpad, "x");
  for (int x : win.xs) if (x % 4) win.vec_ok = false;
  // windows covering each padded coordinate form a contiguous index range because origins are sorted
  auto cover = [&](const std::vector<int>& st, int pad, std::vector<int>& first, std::vector<int>& count) {
    first.assign(pad, 0);
    count.assign(pad, 0);
    for (int i = (int)st.size() - 1; i >= 0; --i)
      for (int c = st[i]; c < st[i] + T; ++c) {
        first[c] = i;
        count[c] += 1;
      }
  };
  std::vector<int> yf, yc, xf, xc;
  cover(win.ys, Hpad, yf, yc);
  cover(win.xs, Wpad, xf, xc);
  win.max_cover = std::max(*std::max_element(yc.begin(), yc.end()), *std::max_element(xc.begin(), xc.end()));
  std::vector<int> org((size_t)ny * nx * 2);
  for (int yi = 0; yi < ny; ++yi)
    for (int xi = 0; xi < nx; ++xi) {
      org[((size_t)yi * nx + xi) * 2 + 0] = ys[yi];
      org[((size_t)yi * nx + xi) * 2 + 1] = xs[xi];
    }
  upload_ints(win.d_ys, win.ys, s);
  upload_ints(win.d_xs, win.xs, s);
  upload_ints(win.d_yfirst, yf, s);
  upload_ints(win.d_ycount, yc, s);
  upload_ints(win.d_xfirst, xf, s);
  upload_ints(win.d_xcount, xc, s);
  upload_ints(win.d_origins, org, s);
  win.has_ramp = (method == FSR_WINDOW_FEATHER);
  if (win.has_ramp) {
    win.d_ramp.ensure((size_t)T * sizeof(float));
    FSR_CUDA(cudaMemcpyAsync(win.d_ramp.p, ramp, (size_t)T * sizeof(float), cudaMemcpyHostToDevice, s));
  }
  // weight of the first / second covering window at every coordinate: the feather ramp with the scene-edge flattening of
  // ResUNet_16x_DEM.py:344-352 (the values k_blend.cu's edge_weight() returns), looked up once here instead of per pixel
  win.Hpad = Hpad; win.Wpad = Wpad;
  auto weights = [&](const std::vector<int>& st, int pad, const std::vector<int>& first, const std::vector<int>& count) {
    std::vector<float> tab((size_t)2 * pad, 1.0f);
    const int n = (int)st.size();
    for (int c = 0; c < pad; ++c)
      for (int d = 0; d < 2 && d < count[c]; ++d) {
        const int idx = first[c] + d, local = c - st[idx];
        float w = 1.0f;
        if (win.has_ramp && !(idx == 0 && local < overlap) && !(idx == n - 1 && local >= T - overlap)) w = ramp[local];
        tab[(size_t)d * pad + c] = w;
      }
    return tab;
  };
  const std::vector<float> wy_tab = weights(win.ys, Hpad, yf, yc), wx_tab = weights(win.xs, Wpad, xf, xc);
  win.d_wy_tab.ensure(wy_tab.size() * sizeof(float));
  win.d_wx_tab.ensure(wx_tab.size() * sizeof(float));
  FSR_CUDA(cudaMemcpyAsync(win.d_wy_tab.p, wy_tab.data(), wy_tab.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(win.d_wx_tab.p, wx_tab.data(), wx_tab.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  // the host vectors above are pageable temporaries: make sure the copies are done before they die
  FSR_CUDA(cudaStreamSynchronize(s));
}

BlendGeom Engine::blend_geom() const {
  BlendGeom g;
  g.y_starts = win.d_ys.as<int>();
  g.x_starts = win.d_xs.as<int>();
  g.ny = (int)win.ys.size();
  g.nx = (int)win.xs.size();
  g.ramp = win.has_ramp ? win.d_ramp.as<float>() : nullptr;
  g.y_first = win.d_yfirst.as<int>();
  g.y_count = win.d_ycount.as<int>();
  g.x_first = win.d_xfirst.as<int>();
  g.x_count = win.d_xcount.as<int>();
  g.T = win.T;
  g.overlap = win.overlap;
  g.H = win.H;
  g.W = win.W;
  g.vec_ok = win.vec_ok ? 1 : 0;
  g.max_cover = win.max_cover;
  g.wy_tab = win.d_wy_tab.as<float>();
  g.wx_tab = win.d_wx_tab.as<float>();
  g.Hpad = win.Hpad;
  g.Wpad = win.Wpad;
  return g;
}

// ------------------------------------------------------------------------------------------------------
// band-level driver shared by the host and device entry points
// ------------------------------------------------------------------------------------------------------

static void band_rows(const Engine& e, int ty0, int ty1, int& row0, int& n_rows, int& halo_out) {
  const auto& ys = e.win.ys;
  const int ny = (int)ys.size(), T = e.win.T, H = e.win.H;
  row0 = ty0 == 0 ? 0 : std::min(ys[ty0], H);
  const int row_end = ty1 >= ny ? H : std::min(ys[ty1], H);
  n_rows = std::max(row_end - row0, 0);
  halo_out = ty1 >= ny ? 0 : std::max(std::min(ys[ty1 - 1] + T, H) - row_end, 0);
}

// Window grid of the window rows [ty0, ty1) over raster rows that start at band_row0 (origins relative to that row).
static TileGrid band_grid(Engine& e, int band_row0, int band_rows_hr, int ty0, int ty1, const int2* d_origins, cudaStream_t s) {
  const int ny = (int)e.win.ys.size(), nx = (int)e.win.xs.size(), T = e.win.T;
  FSR_REQUIRE(ty0 >= 0 && ty0 < ty1 && ty1 <= ny, "bad tile-row range");
  FSR_REQUIRE(band_row0 % e.scale() == 0 && band_row0 <= e.win.ys[ty0], "band rows do not cover the band's first window");
  const int need_end = std::min(e.win.ys[ty1 - 1] + T, e.win.H);
  FSR_REQUIRE(band_row0 + band_rows_hr >= need_end, "band rows do not cover the band's last window");
  const int n_tiles = (ty1 - ty0) * nx;
  TileGrid grid;
  if (d_origins) {
    grid.origins = d_origins;  // uploaded by the caller, relative to band_row0
  } else if (band_row0 == 0) {
    // the buffers hold the raster from its first row: the global window origins uploaded by setup_windows apply
    grid.origins = e.win.d_origins.as<int2>() + (size_t)ty0 * nx;
  } else {
    // origins local to the band's rows
    std::vector<int> org((size_t)n_tiles * 2);
    for (int yi = ty0; yi < ty1; ++yi)
      for (int xi = 0; xi < nx; ++xi) {
        org[((size_t)(yi - ty0) * nx + xi) * 2 + 0] = e.win.ys[yi] - band_row0;
        org[((size_t)(yi - ty0) * nx + xi) * 2 + 1] = e.win.xs[xi];
      }
    e.d_tmp_a.ensure(org.size() * sizeof(int));
    FSR_CUDA(cudaMemcpyAsync(e.d_tmp_a.p, org.data(), org.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    FSR_CUDA(cudaStreamSynchronize(s));
    grid.origins = e.d_tmp_a.as<int2>();
  }
  grid.H = std::min(e.win.H - band_row0, band_rows_hr);
  grid.W = e.win.W;
  grid.Hl = ceil_div(grid.H, e.scale());
  grid.Hl = std::min(grid.Hl, e.win.H / e.scale() - band_row0 / e.scale());
  grid.Wl = e.win.W / e.scale();
  return grid;
}

// First phase for a GROUP of bands, window rows [gy0, gy1): a5-a8 and the low-resolution layers of all its windows in one
// batch (Engine::group_lr).  The bands of the group then call band_run with group_ty0 = gy0.
static void group_run_lr(Engine& e, const float* d_depth, const float* d_dem, int band_row0, int band_rows_hr, int gy0, int gy1,
                         const fsr_tile_params& p, float* d_stats, cudaStream_t s, const int2* d_origins = nullptr) {
  const TileGrid grid = band_grid(e, band_row0, band_rows_hr, gy0, gy1, d_origins, s);
  e.group_lr(d_depth, d_dem, grid, (gy1 - gy0) * (int)e.win.xs.size(), p, d_stats, s);
}

// d_depth/d_dem hold raster rows starting at band_row0 (HR rows; band_row0 % scale == 0), band_rows_hr of them.
// group_ty0 >= 0: the band belongs to the group whose first phase has run (group_run_lr), only the high-resolution kernel and
// the hand-over sums remain.
static void band_run(Engine& e, const float* d_depth, const float* d_dem, int band_row0, int band_rows_hr, int ty0, int ty1,
                     const fsr_tile_params& p, float* d_halo_out, float* d_stats, cudaStream_t s, DeviceBuf* tiles_buf = nullptr,
                     const int2* d_origins = nullptr, int group_ty0 = -1) {
  const int nx = (int)e.win.xs.size(), T = e.win.T;
  const int n_tiles = (ty1 - ty0) * nx;
  DeviceBuf& tiles = tiles_buf ? *tiles_buf : e.d_tiles;
  tiles.ensure((size_t)n_tiles * T * T * sizeof(float));
  if (group_ty0 >= 0) {
    FSR_REQUIRE(ty0 >= group_ty0 && ty0 < ty1 && ty1 <= (int)e.win.ys.size(), "bad tile-row range");
    e.group_hr((ty0 - group_ty0) * nx, n_tiles, tiles.as<float>(), p, s);
  } else {
    float* stats = d_stats;
    if (!stats) {
      e.d_stats.ensure((size_t)n_tiles * 3 * sizeof(float));
      stats = e.d_stats.as<float>();
    }
    const TileGrid grid = band_grid(e, band_row0, band_rows_hr, ty0, ty1, d_origins, s);
    e.run_tiles_from_grid(d_depth, d_dem, grid, 0, n_tiles, p, tiles.as<float>(), nullptr, stats, s);
  }
  int row0, n_rows, halo_out;
  band_rows(e, ty0, ty1, row0, n_rows, halo_out);
  e.band = BandState{ty0, ty1, row0, n_rows, halo_out, p.max_depth};
  if (halo_out > 0 && d_halo_out) {
    BlendGeom g = e.blend_geom();
    ProfScope scope(e.prof, PROF_BLEND, s);
    launch_blend(tiles.as<float>(), ty0, ty1, g, row0 + n_rows, halo_out, nullptr, 0, false, p.max_depth, d_halo_out, s);
  }
}

// Sub-bands of the window rows [ty0, ty1) for the copy / compute pipelines: enough windows per band to keep the batched layer
// kernels efficient, small enough to overlap copies; the first band is a single window row so that the pipeline fills fast.
// A band must own at least the rows it receives partial sums for (three or more window rows covering one coordinate:
// overlap >= tile / 2 or a forced trailing window close to its predecessor): such a band takes over the following window rows.
static std::vector<int> sub_band_plan(const Engine& e, int ty0, int ty1, bool two_phase = false) {
  const int ny = (int)e.win.ys.size(), nx = (int)e.win.xs.size(), T = e.win.T, H = e.win.H;
  // band size: the engine's target, but at least ~4 bands per call so that small rasters still overlap copies and kernels;
  // two_phase: the two-phase pipeline, where the batched layers run per GROUP of bands and a band only has to feed the fused kernel
  const int total = (ty1 - ty0) * nx;
  const int target = std::min(e.band_tiles_target(), std::max(64, ceil_div(total, 4)));
  std::vector<int> band_ty{ty0};  // band b covers window rows [band_ty[b], band_ty[b + 1])
  if (two_phase) {
    // one window row per band, or as many rows as make ~64 windows on narrow rasters (a launch of the fused kernel wants
    // >= 148 x a few rows); what follows the last kernel is exposed, so the last band's worth of rows goes in bands of ~16
    // windows (one row on wide rasters, where the last one is cut into column parts as well)
    const int rows_per_band = std::max(1, ceil_div(64, nx)), tail_rows = std::max(1, ceil_div(16, nx));
    int y = ty0;
    while (ty1 - y >= 2 * rows_per_band) band_ty.push_back(y += rows_per_band);
    while (ty1 - y > tail_rows) band_ty.push_back(y += tail_rows);
  } else {
    const int rows_per_band = std::max(1, ceil_div(target, nx));
    if (ty1 - ty0 >= 3 && rows_per_band > 1) band_ty.push_back(ty0 + 1);
    while (band_ty.back() + rows_per_band < ty1) band_ty.push_back(band_ty.back() + rows_per_band);
  }
  band_ty.push_back(ty1);
  for (size_t b = 0; b + 1 < band_ty.size();) {
    const bool receives = band_ty[b] > 0;
    const int halo_end = receives ? std::min(e.win.ys[band_ty[b] - 1] + T, H) : 0;
    if (receives && band_ty[b + 1] < ty1 && band_ty[b + 1] < ny && std::min(e.win.ys[band_ty[b + 1]], H) < halo_end)
      band_ty.erase(band_ty.begin() + b + 1);
    else
      ++b;
  }
  // the last sub-band cannot grow downwards: if it is still too short, it joins the one above it
  while (band_ty.size() > 2) {
    const int start = band_ty[band_ty.size() - 2];
    const int row_end = ty1 >= ny ? H : std::min(e.win.ys[ty1], H);
    if (start > 0 && row_end - std::min(e.win.ys[start], H) < std::min(e.win.ys[start - 1] + T, H) - std::min(e.win.ys[start], H))
      band_ty.erase(band_ty.end() - 2);
    else
      break;
  }
  return band_ty;
}

// Groups of consecutive bands for the two-phase pipeline: group k covers bands [g[k], g[k + 1]).  The first group is the first
// band alone (the pipeline starts as soon as one window row has arrived), the following groups double in size up to the
// engine's target: the batched low-resolution layers lose throughput on small batches, while a group cannot start before its
// last row has been copied in.  Without the two-phase path every band is its own group.
static std::vector<int> group_plan(const Engine& e, const std::vector<int>& band_ty, bool phases) {
  const int n_bands = (int)band_ty.size() - 1, nx = (int)e.win.xs.size();
  std::vector<int> g{0};
  if (!phases) {
    for (int b = 1; b <= n_bands; ++b) g.push_back(b);
    return g;
  }
  const int cap = std::min(e.chunk_tiles(), e.group_tiles_target());
  long long limit = (long long)(band_ty[1] - band_ty[0]) * nx;
  for (int b = 0; b < n_bands;) {
    int b1 = b + 1;
    while (b1 < n_bands && (long long)(band_ty[b1 + 1] - band_ty[b]) * nx <= limit) ++b1;
    g.push_back(b1);
    b = b1;
    limit = std::min<long long>(2 * limit, cap);
  }
  return g;
}

// the two-phase pipeline applies when the fused tensor-core kernels run the high-resolution layers and no band exceeds a chunk
static bool use_phases(const Engine& e, const fsr_tile_params& p, const std::vector<int>& band_ty) {
  // FSR_PHASES=0 / 1 overrides the default (read per call: the tests switch it).  Default: the compute-bound fp32 mode only --
  // in the 16-bit mode a window row is computed faster than PCIe delivers it, what matters there is an early first band, and
  // the band pipeline measured equal (4096 x 32768) or faster (8192 x 8192: 8.6 vs 9.3 ms; 256 tiles: 8.2 vs 9.6 ms)
  const char* force = getenv("FSR_PHASES");
  const bool want = force ? atoi(force) != 0 : e.compute_bound();
  if (!want || !e.phases_ok() || !p.normalize_inputs) return false;
  for (size_t b = 0; b + 1 < band_ty.size(); ++b)
    if ((long long)(band_ty[b + 1] - band_ty[b]) * (long long)e.win.xs.size() > e.chunk_tiles()) return false;
  return true;
}

static void band_finalize(Engine& e, const float* d_halo_in, int halo_rows_in, float* d_out_rows, cudaStream_t s,
                          const DeviceBuf* tiles_buf = nullptr, const BandState* st = nullptr, int row_begin = 0, int row_end = -1,
                          int x0 = 0, int x1 = -1) {
  BlendGeom g = e.blend_geom();
  const BandState& b = st ? *st : e.band;
  const DeviceBuf& tiles = tiles_buf ? *tiles_buf : e.d_tiles;
  if (row_end < 0) row_end = b.n_rows;
  if (row_end <= row_begin) return;
  const bool first = row_begin == 0;  // only the band's first rows start from the previous band's partial sums
  FSR_REQUIRE(!d_halo_in || halo_rows_in <= b.n_rows,
              "band owns fewer rows than the incoming halo covers: merge window rows when planning bands (dist.chain_safe_bands)");
  ProfScope scope(e.prof, PROF_BLEND, s);
  launch_blend(tiles.as<float>(), b.ty0, b.ty1, g, b.row0 + row_begin, row_end - row_begin, first ? d_halo_in : nullptr,
               first ? halo_rows_in : 0, true, b.max_depth, d_out_rows + (size_t)row_begin * g.W, s, x0, x1);
}

// The LAST band of a two-phase pipeline in column parts: nothing can overlap what follows the last kernel, so the band's
// window row is cut into `parts` runs of windows; each run's high-resolution kernel is followed by the blend of the columns it
// completes (pixels left of window k's origin are covered by windows < k only) and their D2H copy, which then runs next to
// the following run's kernel.  Same tiles, same per-pixel sums: bit-identical to the band in one piece.
// d_out_rows / h_out_rows: the band's first owned row on the device / in the caller's (host) raster, both W floats per row.
static void band_run_tail(Engine& e, int ty0, int group_ty0, const fsr_tile_params& p, const float* d_halo_in, int halo_rows_in,
                          float* d_halo_out, float* d_out_base, float* h_out_base, int rank_row0, int parts, cudaStream_t sc,
                          cudaStream_t so) {
  const int nx = (int)e.win.xs.size(), T = e.win.T, W = e.win.W;
  e.d_tiles.ensure((size_t)nx * T * T * sizeof(float));
  int row0, n_rows, halo_out;
  band_rows(e, ty0, ty0 + 1, row0, n_rows, halo_out);
  e.band = BandState{ty0, ty0 + 1, row0, n_rows, halo_out, p.max_depth};
  float* d_rows = d_out_base + (size_t)(row0 - rank_row0) * W;
  float* h_rows = h_out_base + (size_t)(row0 - rank_row0) * W;
  for (int part = 0; part < parts; ++part) {
    const int k0 = nx * part / parts, k1 = nx * (part + 1) / parts;
    if (k1 <= k0) continue;
    e.group_hr((ty0 - group_ty0) * nx + k0, k1 - k0, e.d_tiles.as<float>() + (size_t)k0 * T * T, p, sc);
    const int x0 = part == 0 ? 0 : e.win.xs[k0], x1 = part == parts - 1 ? W : std::min(e.win.xs[k1], W);
    if (n_rows <= 0 || x1 <= x0) continue;
    band_finalize(e, d_halo_in, halo_rows_in, d_rows, sc, nullptr, nullptr, 0, -1, x0, x1);
    cudaEvent_t ev;
    FSR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    FSR_CUDA(cudaEventRecord(ev, sc));
    FSR_CUDA(cudaStreamWaitEvent(so, ev, 0));
    FSR_CUDA(cudaEventDestroy(ev));  // released once the recorded work has completed
    FSR_CUDA(cudaMemcpy2DAsync(h_rows + x0, (size_t)W * sizeof(float), d_rows + x0, (size_t)W * sizeof(float),
                               (size_t)(x1 - x0) * sizeof(float), (size_t)n_rows, cudaMemcpyDeviceToHost, so));
  }
  if (halo_out > 0 && d_halo_out) {
    BlendGeom g = e.blend_geom();
    ProfScope scope(e.prof, PROF_BLEND, sc);
    launch_blend(e.d_tiles.as<float>(), ty0, ty0 + 1, g, row0 + n_rows, halo_out, nullptr, 0, false, p.max_depth, d_halo_out, sc);
  }
}

// column parts of the last band (env FSR_TAIL_PARTS; 1 = off): needs the two-phase path, a band of one window row and window
// origins the vector blend accepts
static int tail_parts(const Engine& e, bool phases, int ty0, int ty1) {
  int parts = 4;
  if (const char* v = getenv("FSR_TAIL_PARTS")) parts = std::max(1, atoi(v));
  const int nx = (int)e.win.xs.size();
  if (!phases || ty1 - ty0 != 1 || !e.win.vec_ok || e.win.W % 4 != 0) return 1;
  return std::max(1, std::min(parts, nx / 8));
}

}  // namespace fsr

// ------------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------------

using namespace fsr;

struct fsr_engine {
  Engine impl;
  fsr_engine(const void* plan, size_t pb, const float* w, size_t wc, int dev, int prec) : impl(plan, pb, w, wc, dev, prec) {}
};

#define FSR_API_BEGIN(eng)                          \
  try {                                             \
    if (!(eng)) throw Error(FSR_E_INVALID, "engine handle is NULL"); \
    (eng)->impl.set_device();                       \
    g_launch_counter = &(eng)->impl.launches;
#define FSR_API_END()                               \
    return FSR_OK;                                  \
  } catch (const Error& e) {                        \
    g_last_error = e.what();                        \
    cudaGetLastError();                             \
    return e.code;                                  \
  } catch (const std::exception& e) {               \
    g_last_error = e.what();                        \
    return FSR_E_INVALID;                           \
  }

static void check_params(const fsr_tile_params* p, int n_px) {
  FSR_REQUIRE(p != nullptr, "params is NULL");
  FSR_REQUIRE(p->max_depth > 0.f && p->depth_denom > 0.f, "max_depth and depth_denom must be > 0");
  if (p->normalize_inputs) {
    FSR_REQUIRE(p->rank_lo >= 0 && p->rank_lo <= p->rank_hi && p->rank_hi < n_px, "percentile ranks outside the tile");
    FSR_REQUIRE(p->gamma >= 0.f && p->gamma < 1.f, "percentile gamma must be in [0, 1)");
  }
}

extern "C" {

int fsr_abi_version(void) { return FSR_ABI_VERSION; }

const char* fsr_last_error(void) { return g_last_error.c_str(); }

int fsr_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int fsr_create(const void* plan, size_t plan_bytes, const float* weights, size_t weights_count, int device, int precision,
               fsr_engine** out) {
  try {
    FSR_REQUIRE(out != nullptr, "out is NULL");
    *out = nullptr;
    *out = new fsr_engine(plan, plan_bytes, weights, weights_count, device, precision);
    return FSR_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    cudaGetLastError();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return FSR_E_INVALID;
  }
}

int fsr_destroy(fsr_engine* eng) {
  if (!eng) return FSR_OK;
  g_launch_counter = nullptr;
  delete eng;
  return FSR_OK;
}

int fsr_contract(const fsr_engine* eng, int32_t* lr_tile, int32_t* hr_tile, int32_t* scale) {
  if (!eng) {
    g_last_error = "engine handle is NULL";
    return FSR_E_INVALID;
  }
  if (lr_tile) *lr_tile = eng->impl.lr_tile();
  if (hr_tile) *hr_tile = eng->impl.hr_tile();
  if (scale) *scale = eng->impl.scale();
  return FSR_OK;
}

int64_t fsr_launch_count(const fsr_engine* eng) { return eng ? eng->impl.launches.n : 0; }

int64_t fsr_macs_per_tile(const fsr_engine* eng) { return eng ? eng->impl.macs_per_tile() : 0; }

int fsr_run_tiles(fsr_engine* eng, const float* depth_lr, const float* dem_hr, int32_t n_tiles, const fsr_tile_params* params,
                  float* out_pred_m, float* out_pred_norm, float* out_stats, uint32_t* out_flags) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  const int T = e.hr_tile(), TL = e.lr_tile();
  FSR_REQUIRE(n_tiles > 0 && depth_lr && dem_hr && out_pred_m, "empty batch or NULL buffer");
  check_params(params, T * T);
  const size_t hr_px = (size_t)T * T, lr_px = (size_t)TL * TL;
  cudaStream_t s = 0;
  e.d_in_depth.ensure(lr_px * n_tiles * sizeof(float));
  e.d_in_dem.ensure(hr_px * n_tiles * sizeof(float));
  e.d_tiles.ensure(hr_px * n_tiles * sizeof(float));
  e.d_stats.ensure((size_t)n_tiles * 3 * sizeof(float));
  if (out_pred_norm) e.d_out.ensure(hr_px * n_tiles * sizeof(float));
  FSR_CUDA(cudaMemcpyAsync(e.d_in_depth.p, depth_lr, lr_px * n_tiles * sizeof(float), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(e.d_in_dem.p, dem_hr, hr_px * n_tiles * sizeof(float), cudaMemcpyHostToDevice, s));
  // a batch of independent tiles is the raster [n*T, T] with one window per tile
  std::vector<int> org((size_t)n_tiles * 2);
  for (int t = 0; t < n_tiles; ++t) {
    org[2 * t] = t * T;
    org[2 * t + 1] = 0;
  }
  e.d_tmp_a.ensure(org.size() * sizeof(int));
  FSR_CUDA(cudaMemcpyAsync(e.d_tmp_a.p, org.data(), org.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  TileGrid grid{e.d_tmp_a.as<int2>(), n_tiles * T, T, n_tiles * TL, TL};
  e.run_tiles_from_grid(e.d_in_depth.as<float>(), e.d_in_dem.as<float>(), grid, 0, n_tiles, *params, e.d_tiles.as<float>(),
                        out_pred_norm ? e.d_out.as<float>() : nullptr, e.d_stats.as<float>(), s);
  FSR_CUDA(cudaMemcpyAsync(out_pred_m, e.d_tiles.p, hr_px * n_tiles * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (out_pred_norm) FSR_CUDA(cudaMemcpyAsync(out_pred_norm, e.d_out.p, hr_px * n_tiles * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (out_stats) FSR_CUDA(cudaMemcpyAsync(out_stats, e.d_stats.p, (size_t)n_tiles * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
  unsigned f = e.fetch_flags(s);
  if (out_flags) *out_flags = f;
  if (f) throw Error(FSR_E_ASSERT, "input validation failed on the device (see flags)");
  FSR_API_END()
}

int fsr_run_raster(fsr_engine* eng, const float* depth_lr, const float* dem_hr, int32_t H, int32_t W, int32_t window_method,
                   int32_t overlap_hr, const int32_t* y_starts, int32_t ny, const int32_t* x_starts, int32_t nx,
                   const float* ramp, const fsr_tile_params* params, float* out_sr, float* out_stats, uint32_t* out_flags) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  const int T = e.hr_tile(), sc = e.scale();
  FSR_REQUIRE(depth_lr && dem_hr && out_sr && H > 0 && W > 0, "empty raster or NULL buffer");
  check_params(params, T * T);
  const int Hl = H / sc, Wl = W / sc;
  FSR_REQUIRE(Hl > 0 && Wl > 0, "raster smaller than one low-resolution cell");
  // Software pipeline over bands of window rows: the H2D copy of the rows band b+1 reads, the kernels of band b and
  // the D2H copy of the rows band b-1 owns run on three streams.  Bands hand their shared rows over as partial
  // sums exactly like ranks do (fsr_band_*), so the result is bit-identical to a single pass.  Copies only overlap
  // when the host buffers are page-locked (fsr_host_alloc); pageable buffers work but serialise.
  e.ensure_streams();
  cudaStream_t sc_ = e.s_comp, si = e.s_in, so = e.s_out;
  e.setup_windows(H, W, window_method, overlap_hr, y_starts, ny, x_starts, nx, ramp, sc_);
  e.d_in_depth.ensure((size_t)Hl * Wl * sizeof(float));
  e.d_in_dem.ensure((size_t)H * W * sizeof(float));
  e.d_out.ensure((size_t)H * W * sizeof(float));
  e.d_stats.ensure((size_t)ny * nx * 3 * sizeof(float));
  // (a single-row LAST band was measured slower than letting the last band keep its two rows: small batches run the LR
  // layers at a third of their throughput)
  std::vector<int> band_ty = sub_band_plan(e, 0, ny, true);
  const bool phases = use_phases(e, *params, band_ty);
  if (!phases) band_ty = sub_band_plan(e, 0, ny);
  const std::vector<int> group_b = group_plan(e, band_ty, phases);
  const int n_bands = (int)band_ty.size() - 1;
  e.d_halo[0].ensure((size_t)T * W * sizeof(float));
  e.d_halo[1].ensure((size_t)T * W * sizeof(float));
  std::vector<cudaEvent_t> ev_in(n_bands), ev_done(n_bands);
  for (int b = 0; b < n_bands; ++b) {
    FSR_CUDA(cudaEventCreateWithFlags(&ev_in[b], getenv("FSR_RASTER_TIMING") ? cudaEventDefault : cudaEventDisableTiming));
    FSR_CUDA(cudaEventCreateWithFlags(&ev_done[b], getenv("FSR_RASTER_TIMING") ? cudaEventDefault : cudaEventDisableTiming));
  }
  struct EvGuard {
    std::vector<cudaEvent_t>&a, &b;
    ~EvGuard() {
      for (auto x : a) cudaEventDestroy(x);
      for (auto x : b) cudaEventDestroy(x);
    }
  } guard{ev_in, ev_done};
  const bool timing = getenv("FSR_RASTER_TIMING") != nullptr;
  const auto h0 = std::chrono::steady_clock::now();
  double h_enq_in = 0, h_band[16] = {0};
  cudaEvent_t t_start = nullptr, t_in = nullptr, t_comp = nullptr, t_out = nullptr;
  if (timing) {
    for (cudaEvent_t* ev : {&t_start, &t_in, &t_comp, &t_out}) FSR_CUDA(cudaEventCreate(ev));
    FSR_CUDA(cudaEventRecord(t_start, si));
  }
  // all input copies are queued up front, in band order
  FSR_CUDA(cudaMemcpyAsync(e.d_in_depth.p, depth_lr, (size_t)Hl * Wl * sizeof(float), cudaMemcpyHostToDevice, si));
  int copied = 0;
  for (int b = 0; b < n_bands; ++b) {
    const int ty1 = band_ty[b + 1];
    const int need = b == n_bands - 1 ? H : std::min(e.win.ys[ty1 - 1] + T, (int)H);
    if (need > copied) {
      FSR_CUDA(cudaMemcpyAsync(e.d_in_dem.as<float>() + (size_t)copied * W, dem_hr + (size_t)copied * W,
                               (size_t)(need - copied) * W * sizeof(float), cudaMemcpyHostToDevice, si));
      copied = need;
    }
    FSR_CUDA(cudaEventRecord(ev_in[b], si));
  }
  h_enq_in = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
  for (int b = 0, g = 0; b < n_bands; ++b) {
    const int ty0 = band_ty[b], ty1 = band_ty[b + 1];
    if (b == group_b[g]) {  // a new group starts: it needs the rows of its last band
      const int gb1 = group_b[g + 1];
      FSR_CUDA(cudaStreamWaitEvent(sc_, ev_in[gb1 - 1], 0));
      if (phases)
        group_run_lr(e, e.d_in_depth.as<float>(), e.d_in_dem.as<float>(), 0, H, ty0, band_ty[gb1], *params,
                     e.d_stats.as<float>() + (size_t)ty0 * nx * 3, sc_);
      ++g;
    }
    const int n_tail = (b == n_bands - 1 && b > 0) ? tail_parts(e, phases, ty0, ty1) : 1;
    if (n_tail > 1) {
      band_run_tail(e, ty0, band_ty[group_b[g - 1]], *params, e.d_halo[(b - 1) & 1].as<float>(), e.band_halo_rows[(b - 1) & 1], nullptr,
                    e.d_out.as<float>(), out_sr, 0, n_tail, sc_, so);
      FSR_CUDA(cudaEventRecord(ev_done[b], sc_));
      if (b < 16) h_band[b] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
      continue;
    }
    band_run(e, e.d_in_depth.as<float>(), e.d_in_dem.as<float>(), 0, H, ty0, ty1, *params, e.d_halo[b & 1].as<float>(),
             e.d_stats.as<float>() + (size_t)ty0 * nx * 3, sc_, nullptr, nullptr, phases ? band_ty[group_b[g - 1]] : -1);
    const int halo_in = b == 0 ? 0 : e.band_halo_rows[(b - 1) & 1];
    e.band_halo_rows[b & 1] = e.band.halo_out_rows;
    band_finalize(e, b == 0 ? nullptr : e.d_halo[(b - 1) & 1].as<float>(), halo_in, e.d_out.as<float>() + (size_t)e.band.row0 * W, sc_);
    FSR_CUDA(cudaEventRecord(ev_done[b], sc_));
    FSR_CUDA(cudaStreamWaitEvent(so, ev_done[b], 0));
    if (e.band.n_rows > 0)
      FSR_CUDA(cudaMemcpyAsync(out_sr + (size_t)e.band.row0 * W, e.d_out.as<float>() + (size_t)e.band.row0 * W,
                               (size_t)e.band.n_rows * W * sizeof(float), cudaMemcpyDeviceToHost, so));
    if (b < 16) h_band[b] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
  }
  if (out_stats) FSR_CUDA(cudaMemcpyAsync(out_stats, e.d_stats.p, (size_t)ny * nx * 3 * sizeof(float), cudaMemcpyDeviceToHost, sc_));
  if (timing) {
    FSR_CUDA(cudaEventRecord(t_in, si));
    FSR_CUDA(cudaEventRecord(t_comp, sc_));
    FSR_CUDA(cudaEventRecord(t_out, so));
  }
  unsigned f = e.fetch_flags(sc_);
  FSR_CUDA(cudaStreamSynchronize(so));
  if (timing) {
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, t_start, t_in);
    cudaEventElapsedTime(&b, t_start, t_comp);
    cudaEventElapsedTime(&c, t_start, t_out);
    fprintf(stderr, "[fsr_run_raster] %d bands in %d groups: H2D done %.2f ms, kernels done %.2f ms, D2H done %.2f ms | host: inputs queued %.2f, bands queued",
            n_bands, (int)group_b.size() - 1, a, b, c, h_enq_in);
    for (int k = 0; k < n_bands && k < 16; ++k) fprintf(stderr, " %.2f", h_band[k]);
    fprintf(stderr, " ms | per band (H2D done, kernels done):");
    for (int k = 0; k < n_bands; ++k) {
      float x = 0, y = 0;
      cudaEventElapsedTime(&x, t_start, ev_in[k]);
      cudaEventElapsedTime(&y, t_start, ev_done[k]);
      fprintf(stderr, " (%.2f, %.2f)", x, y);
    }
    fprintf(stderr, "\n");
    for (cudaEvent_t ev : {t_start, t_in, t_comp, t_out}) cudaEventDestroy(ev);
  }
  if (out_flags) *out_flags = f;
  if (f) throw Error(FSR_E_ASSERT, "input validation failed on the device (see flags)");
  FSR_API_END()
}

int fsr_set_windows(fsr_engine* eng, int32_t H, int32_t W, int32_t window_method, int32_t overlap_hr, const int32_t* y_starts,
                    int32_t ny, const int32_t* x_starts, int32_t nx, const float* ramp, void* stream) {
  FSR_API_BEGIN(eng)
  eng->impl.setup_windows(H, W, window_method, overlap_hr, y_starts, ny, x_starts, nx, ramp, (cudaStream_t)stream);
  FSR_API_END()
}

int fsr_band_geometry(fsr_engine* eng, int32_t ty0, int32_t ty1, int32_t* row0, int32_t* n_rows, int32_t* halo_out_rows,
                      int32_t* in_row0, int32_t* in_rows) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  const int ny = (int)e.win.ys.size();
  FSR_REQUIRE(ny > 0, "fsr_set_windows has not been called");
  FSR_REQUIRE(ty0 >= 0 && ty0 < ty1 && ty1 <= ny, "bad tile-row range");
  int r0, nr, ho;
  band_rows(e, ty0, ty1, r0, nr, ho);
  if (row0) *row0 = r0;
  if (n_rows) *n_rows = nr;
  if (halo_out_rows) *halo_out_rows = ho;
  const int first = std::min(e.win.ys[ty0], e.win.H);
  const int last = std::min(e.win.ys[ty1 - 1] + e.win.T, e.win.H);
  if (in_row0) *in_row0 = first;
  if (in_rows) *in_rows = std::max(last - first, 0);
  FSR_API_END()
}

int fsr_band_run_dev(fsr_engine* eng, const float* d_depth_lr, const float* d_dem_hr, int32_t band_row0, int32_t band_rows_hr,
                     int32_t ty0, int32_t ty1, const fsr_tile_params* params, float* d_halo_out, float* d_stats, void* stream) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  check_params(params, e.hr_tile() * e.hr_tile());
  FSR_REQUIRE(!e.win.ys.empty(), "fsr_set_windows has not been called");
  FSR_REQUIRE(d_depth_lr && d_dem_hr, "NULL device buffer");
  band_run(e, d_depth_lr, d_dem_hr, band_row0, band_rows_hr, ty0, ty1, *params, d_halo_out, d_stats, (cudaStream_t)stream);
  FSR_API_END()
}

int fsr_band_finalize_rows_dev(fsr_engine* eng, const float* d_halo_in, int32_t halo_rows_in, float* d_out_rows,
                               int32_t row_begin, int32_t row_end, void* stream) {
  FSR_API_BEGIN(eng)
  FSR_REQUIRE(d_out_rows != nullptr, "NULL device buffer");
  Engine& e = eng->impl;
  FSR_REQUIRE(e.band.ty1 > e.band.ty0, "fsr_band_run_dev has not been called");
  FSR_REQUIRE(row_begin >= 0 && row_begin <= row_end && row_end <= e.band.n_rows, "row range outside the band");
  FSR_REQUIRE(row_begin == 0 ? halo_rows_in <= row_end : halo_rows_in <= row_begin,
              "the rows that start from the previous band's sums must lie in one call");
  band_finalize(e, d_halo_in, halo_rows_in, d_out_rows, (cudaStream_t)stream, nullptr, nullptr, row_begin, row_end);
  FSR_API_END()
}

int fsr_band_finalize_dev(fsr_engine* eng, const float* d_halo_in, int32_t halo_rows_in, float* d_out_rows, void* stream) {
  FSR_API_BEGIN(eng)
  FSR_REQUIRE(d_out_rows != nullptr, "NULL device buffer");
  FSR_REQUIRE(eng->impl.band.ty1 > eng->impl.band.ty0, "fsr_band_run_dev has not been called");
  band_finalize(eng->impl, d_halo_in, halo_rows_in, d_out_rows, (cudaStream_t)stream);
  FSR_API_END()
}

// Pipelined host-buffer variant of fsr_band_run_dev + fsr_band_finalize_dev for one rank's band of window rows
// [ty0, ty1): sub-bands overlap their H2D copy, kernels and D2H copy exactly like fsr_run_raster.  The sub-band that
// shares rows with the previous rank is finalised last (fsr_band_host_end), after the caller has received that rank's
// partial sums, so ranks never wait for one another while computing.
int fsr_band_host_begin(fsr_engine* eng, const float* depth_lr, const float* dem_hr, int32_t band_row0, int32_t band_rows_hr,
                        int32_t ty0, int32_t ty1, const fsr_tile_params* params, float* out_rows, float* d_halo_out) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  const int T = e.hr_tile(), sc = e.scale();
  check_params(params, T * T);
  const int ny = (int)e.win.ys.size(), nx = (int)e.win.xs.size(), W = e.win.W, H = e.win.H;
  FSR_REQUIRE(ny > 0, "fsr_set_windows has not been called");
  FSR_REQUIRE(depth_lr && dem_hr && out_rows, "NULL host buffer");
  FSR_REQUIRE(ty0 >= 0 && ty0 < ty1 && ty1 <= ny, "bad tile-row range");
  FSR_REQUIRE(band_row0 % sc == 0 && band_row0 <= e.win.ys[ty0] && band_row0 + band_rows_hr >= std::min(e.win.ys[ty1 - 1] + T, H),
              "host rows do not cover the band's windows");
  e.ensure_streams();
  cudaStream_t sc_ = e.s_comp, si = e.s_in, so = e.s_out;
  const int lr0 = band_row0 / sc;
  const int lr_rows = std::min(ceil_div(band_row0 + band_rows_hr, sc), H / sc) - lr0;
  const int Wl = W / sc;
  int rank_row0, rank_rows, rank_halo;
  band_rows(e, ty0, ty1, rank_row0, rank_rows, rank_halo);
  e.d_in_depth.ensure((size_t)lr_rows * Wl * sizeof(float));
  e.d_in_dem.ensure((size_t)band_rows_hr * W * sizeof(float));
  e.d_out.ensure((size_t)std::max(rank_rows, 1) * W * sizeof(float));
  e.d_halo[0].ensure((size_t)T * W * sizeof(float));
  e.d_halo[1].ensure((size_t)T * W * sizeof(float));
  // window origins relative to the first host row, for all windows of the rank
  {
    std::vector<int> org((size_t)(ty1 - ty0) * nx * 2);
    for (int yi = ty0; yi < ty1; ++yi)
      for (int xi = 0; xi < nx; ++xi) {
        org[((size_t)(yi - ty0) * nx + xi) * 2 + 0] = e.win.ys[yi] - band_row0;
        org[((size_t)(yi - ty0) * nx + xi) * 2 + 1] = e.win.xs[xi];
      }
    e.d_tmp_a.ensure(org.size() * sizeof(int));
    FSR_CUDA(cudaMemcpyAsync(e.d_tmp_a.p, org.data(), org.size() * sizeof(int), cudaMemcpyHostToDevice, sc_));
    FSR_CUDA(cudaStreamSynchronize(sc_));
  }
  std::vector<int> band_ty = sub_band_plan(e, ty0, ty1, true);
  const bool phases = use_phases(e, *params, band_ty);
  if (!phases) band_ty = sub_band_plan(e, ty0, ty1);
  const std::vector<int> group_b = group_plan(e, band_ty, phases);
  const int n_bands = (int)band_ty.size() - 1;
  if (phases) e.d_stats.ensure((size_t)(ty1 - ty0) * nx * 3 * sizeof(float));
  std::vector<cudaEvent_t> ev_in(n_bands), ev_done(n_bands);
  for (int b = 0; b < n_bands; ++b) {
    FSR_CUDA(cudaEventCreateWithFlags(&ev_in[b], cudaEventDisableTiming));
    FSR_CUDA(cudaEventCreateWithFlags(&ev_done[b], cudaEventDisableTiming));
  }
  struct EvGuard {
    std::vector<cudaEvent_t>&a, &b;
    ~EvGuard() {
      for (auto x : a) cudaEventDestroy(x);
      for (auto x : b) cudaEventDestroy(x);
    }
  } guard{ev_in, ev_done};
  FSR_CUDA(cudaMemcpyAsync(e.d_in_depth.p, depth_lr, (size_t)lr_rows * Wl * sizeof(float), cudaMemcpyHostToDevice, si));
  int copied = 0;  // host rows copied so far
  for (int b = 0; b < n_bands; ++b) {
    const int need = b == n_bands - 1 ? band_rows_hr : std::min(e.win.ys[band_ty[b + 1] - 1] + T, H) - band_row0;
    if (need > copied) {
      FSR_CUDA(cudaMemcpyAsync(e.d_in_dem.as<float>() + (size_t)copied * W, dem_hr + (size_t)copied * W,
                               (size_t)(need - copied) * W * sizeof(float), cudaMemcpyHostToDevice, si));
      copied = need;
    }
    FSR_CUDA(cudaEventRecord(ev_in[b], si));
  }
  const bool defer_first = ty0 > 0;  // rows shared with the previous rank: blended in fsr_band_host_end
  for (int b = 0, g = 0; b < n_bands; ++b) {
    const int s0 = band_ty[b], s1 = band_ty[b + 1];
    const bool last = b == n_bands - 1;
    if (b == group_b[g]) {  // a new group starts: it needs the rows of its last band
      const int gb1 = group_b[g + 1];
      FSR_CUDA(cudaStreamWaitEvent(sc_, ev_in[gb1 - 1], 0));
      if (phases)
        group_run_lr(e, e.d_in_depth.as<float>(), e.d_in_dem.as<float>(), band_row0, band_rows_hr, s0, band_ty[gb1], *params,
                     e.d_stats.as<float>() + (size_t)(s0 - ty0) * nx * 3, sc_, e.d_tmp_a.as<int2>() + (size_t)(s0 - ty0) * nx);
      ++g;
    }
    float* halo_dst = (last && d_halo_out) ? d_halo_out : e.d_halo[b & 1].as<float>();
    const int n_tail = (last && b > 0) ? tail_parts(e, phases, s0, s1) : 1;
    if (n_tail > 1) {
      band_run_tail(e, s0, band_ty[group_b[g - 1]], *params, e.d_halo[(b - 1) & 1].as<float>(), e.band_halo_rows[(b - 1) & 1], halo_dst,
                    e.d_out.as<float>(), out_rows, rank_row0, n_tail, sc_, so);
      e.band_halo_rows[b & 1] = e.band.halo_out_rows;
      continue;
    }
    band_run(e, e.d_in_depth.as<float>(), e.d_in_dem.as<float>(), band_row0, band_rows_hr, s0, s1, *params, halo_dst, nullptr, sc_,
             (b == 0 && defer_first) ? &e.d_tiles0 : nullptr, e.d_tmp_a.as<int2>() + (size_t)(s0 - ty0) * nx,
             phases ? band_ty[group_b[g - 1]] : -1);
    const int halo_in = b == 0 ? 0 : e.band_halo_rows[(b - 1) & 1];
    e.band_halo_rows[b & 1] = e.band.halo_out_rows;
    if (b == 0 && defer_first) {
      e.band0 = e.band;
      e.band0_out = out_rows;
      e.band0_rank_row0 = rank_row0;
    } else {
      float* d_rows = e.d_out.as<float>() + (size_t)(e.band.row0 - rank_row0) * W;
      band_finalize(e, b == 0 ? nullptr : e.d_halo[(b - 1) & 1].as<float>(), halo_in, d_rows, sc_);
      FSR_CUDA(cudaEventRecord(ev_done[b], sc_));
      FSR_CUDA(cudaStreamWaitEvent(so, ev_done[b], 0));
      if (e.band.n_rows > 0)
        FSR_CUDA(cudaMemcpyAsync(out_rows + (size_t)(e.band.row0 - rank_row0) * W, d_rows, (size_t)e.band.n_rows * W * sizeof(float),
                                 cudaMemcpyDeviceToHost, so));
    }
  }
  e.band0_pending = defer_first;
  FSR_CUDA(cudaStreamSynchronize(sc_));  // the rank's halo (d_halo_out) is complete; D2H copies may still be in flight
  FSR_API_END()
}

int fsr_band_host_end(fsr_engine* eng, const float* d_halo_in, int32_t halo_rows_in, uint32_t* out_flags) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  FSR_REQUIRE(e.s_comp != nullptr, "fsr_band_host_begin has not been called");
  const int W = e.win.W;
  if (e.band0_pending) {
    float* d_rows = e.d_out.as<float>() + (size_t)(e.band0.row0 - e.band0_rank_row0) * W;
    band_finalize(e, d_halo_in, halo_rows_in, d_rows, e.s_comp, &e.d_tiles0, &e.band0);
    if (e.band0.n_rows > 0)
      FSR_CUDA(cudaMemcpyAsync(e.band0_out + (size_t)(e.band0.row0 - e.band0_rank_row0) * W, d_rows,
                               (size_t)e.band0.n_rows * W * sizeof(float), cudaMemcpyDeviceToHost, e.s_comp));
    e.band0_pending = false;
  }
  unsigned f = e.fetch_flags(e.s_comp);
  FSR_CUDA(cudaStreamSynchronize(e.s_out));
  if (out_flags) *out_flags = f;
  if (f) throw Error(FSR_E_ASSERT, "input validation failed on the device (see flags)");
  FSR_API_END()
}

static void check_resample(const fsr_resample_params* p, int sh, int sw, int dh, int dw) {
  FSR_REQUIRE(p != nullptr, "params is NULL");
  FSR_REQUIRE(sh > 0 && sw > 0 && dh > 0 && dw > 0, "empty raster");
  FSR_REQUIRE(p->x_a_src != 0.0 && p->y_a_src != 0.0 && p->x_a_dst != 0.0 && p->y_a_dst != 0.0, "degenerate transform");
  const double xs = fabs(p->x_a_src / p->x_a_dst), ys = fabs(p->y_a_src / p->y_a_dst);
  FSR_REQUIRE(xs >= 1.0 / 64 && ys >= 1.0 / 64, "down-sampling by more than 64x is not supported");
}

int fsr_resample_bilinear_dev(fsr_engine* eng, const float* d_src, int32_t sh, int32_t sw, float* d_dst, int32_t dh, int32_t dw,
                              const fsr_resample_params* params, void* stream) {
  FSR_API_BEGIN(eng)
  FSR_REQUIRE(d_src && d_dst, "NULL device buffer");
  check_resample(params, sh, sw, dh, dw);
  launch_resample_bilinear(d_src, sh, sw, d_dst, dh, dw, *params, (cudaStream_t)stream);
  FSR_API_END()
}

int fsr_resample_bilinear(fsr_engine* eng, const float* src, int32_t sh, int32_t sw, float* dst, int32_t dh, int32_t dw,
                          const fsr_resample_params* params) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  FSR_REQUIRE(src && dst, "NULL host buffer");
  check_resample(params, sh, sw, dh, dw);
  cudaStream_t s = 0;
  e.d_in_dem.ensure((size_t)sh * sw * sizeof(float));
  e.d_out.ensure((size_t)dh * dw * sizeof(float));
  FSR_CUDA(cudaMemcpyAsync(e.d_in_dem.p, src, (size_t)sh * sw * sizeof(float), cudaMemcpyHostToDevice, s));
  launch_resample_bilinear(e.d_in_dem.as<float>(), sh, sw, e.d_out.as<float>(), dh, dw, *params, s);
  FSR_CUDA(cudaMemcpyAsync(dst, e.d_out.p, (size_t)dh * dw * sizeof(float), cudaMemcpyDeviceToHost, s));
  FSR_CUDA(cudaStreamSynchronize(s));
  FSR_API_END()
}

int fsr_fetch_flags(fsr_engine* eng, void* stream, uint32_t* out_flags) {
  FSR_API_BEGIN(eng)
  unsigned f = eng->impl.fetch_flags((cudaStream_t)stream);
  if (out_flags) *out_flags = f;
  FSR_API_END()
}

int fsr_stage_normalize(fsr_engine* eng, const float* depth_lr, const float* dem_hr, int32_t n_tiles,
                        const fsr_tile_params* params, float* out_depth_norm, float* out_dem_norm, float* out_stats,
                        uint32_t* out_flags) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  const int T = e.hr_tile(), TL = e.lr_tile();
  FSR_REQUIRE(n_tiles > 0 && depth_lr && dem_hr, "empty batch or NULL buffer");
  check_params(params, T * T);
  const size_t hr_px = (size_t)T * T, lr_px = (size_t)TL * TL;
  cudaStream_t s = 0;
  e.d_in_depth.ensure(lr_px * n_tiles * sizeof(float));
  e.d_in_dem.ensure(hr_px * n_tiles * sizeof(float));
  e.d_tmp_b.ensure(lr_px * n_tiles * sizeof(float));
  e.d_out.ensure(hr_px * n_tiles * sizeof(float));
  e.d_stats.ensure((size_t)n_tiles * 3 * sizeof(float));
  FSR_CUDA(cudaMemcpyAsync(e.d_in_depth.p, depth_lr, lr_px * n_tiles * sizeof(float), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(e.d_in_dem.p, dem_hr, hr_px * n_tiles * sizeof(float), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemsetAsync(e.d_stats.p, 0, (size_t)n_tiles * 3 * sizeof(float), s));
  std::vector<int> org((size_t)n_tiles * 2);
  for (int t = 0; t < n_tiles; ++t) {
    org[2 * t] = t * T;
    org[2 * t + 1] = 0;
  }
  e.d_tmp_a.ensure(org.size() * sizeof(int));
  FSR_CUDA(cudaMemcpyAsync(e.d_tmp_a.p, org.data(), org.size() * sizeof(int), cudaMemcpyHostToDevice, s));
  TileGrid grid{e.d_tmp_a.as<int2>(), n_tiles * T, T, n_tiles * TL, TL};
  e.d_norm_ws.ensure(tile_normalize_ws_bytes(n_tiles));
  launch_tile_normalize(e.d_in_dem.as<float>(), e.d_in_depth.as<float>(), grid, 0, n_tiles, T, TL, e.scale(), *params,
                        e.d_out.as<float>(), e.d_tmp_b.as<float>(), e.d_stats.as<float>(), nullptr, e.d_flags(), s, e.d_norm_ws.p);
  if (out_depth_norm) FSR_CUDA(cudaMemcpyAsync(out_depth_norm, e.d_tmp_b.p, lr_px * n_tiles * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (out_dem_norm) FSR_CUDA(cudaMemcpyAsync(out_dem_norm, e.d_out.p, hr_px * n_tiles * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (out_stats) FSR_CUDA(cudaMemcpyAsync(out_stats, e.d_stats.p, (size_t)n_tiles * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
  unsigned f = e.fetch_flags(s);
  if (out_flags) *out_flags = f;
  if (f) throw Error(FSR_E_ASSERT, "input validation failed on the device (see flags)");
  FSR_API_END()
}

int fsr_stage_forward(fsr_engine* eng, const float* depth_norm, const float* dem_norm, int32_t n_tiles, float* out_pred_norm) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  const int T = e.hr_tile(), TL = e.lr_tile();
  FSR_REQUIRE(n_tiles > 0 && depth_norm && dem_norm && out_pred_norm, "empty batch or NULL buffer");
  const size_t hr_px = (size_t)T * T, lr_px = (size_t)TL * TL;
  cudaStream_t s = 0;
  e.d_in_depth.ensure(lr_px * n_tiles * sizeof(float));
  e.d_in_dem.ensure(hr_px * n_tiles * sizeof(float));
  e.d_out.ensure(hr_px * n_tiles * sizeof(float));
  FSR_CUDA(cudaMemcpyAsync(e.d_in_depth.p, depth_norm, lr_px * n_tiles * sizeof(float), cudaMemcpyHostToDevice, s));
  FSR_CUDA(cudaMemcpyAsync(e.d_in_dem.p, dem_norm, hr_px * n_tiles * sizeof(float), cudaMemcpyHostToDevice, s));
  e.forward(n_tiles, e.d_in_depth.as<float>(), e.d_in_dem.as<float>(), e.d_out.as<float>(), nullptr, 5.0f, 1.791759f, s);
  FSR_CUDA(cudaMemcpyAsync(out_pred_norm, e.d_out.p, hr_px * n_tiles * sizeof(float), cudaMemcpyDeviceToHost, s));
  FSR_CUDA(cudaStreamSynchronize(s));
  FSR_API_END()
}

int fsr_stage_invert(fsr_engine* eng, const float* pred_norm, size_t n, float max_depth, float depth_denom, float* out_pred_m) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  FSR_REQUIRE(n > 0 && pred_norm && out_pred_m, "empty input or NULL buffer");
  cudaStream_t s = 0;
  e.d_in_dem.ensure(n * sizeof(float));
  e.d_out.ensure(n * sizeof(float));
  FSR_CUDA(cudaMemcpyAsync(e.d_in_dem.p, pred_norm, n * sizeof(float), cudaMemcpyHostToDevice, s));
  launch_invert_depth(e.d_in_dem.as<float>(), e.d_out.as<float>(), n, max_depth, depth_denom, s);
  FSR_CUDA(cudaMemcpyAsync(out_pred_m, e.d_out.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
  FSR_CUDA(cudaStreamSynchronize(s));
  FSR_API_END()
}

int fsr_stage_blend(fsr_engine* eng, const float* tiles, int32_t H, int32_t W, int32_t window_method, int32_t overlap_hr,
                    const int32_t* y_starts, int32_t ny, const int32_t* x_starts, int32_t nx, const float* ramp,
                    float max_depth, float* out_sr) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  FSR_REQUIRE(tiles && out_sr, "NULL buffer");
  cudaStream_t s = 0;
  e.setup_windows(H, W, window_method, overlap_hr, y_starts, ny, x_starts, nx, ramp, s);
  const size_t tile_px = (size_t)e.hr_tile() * e.hr_tile();
  e.d_tiles.ensure(tile_px * ny * nx * sizeof(float));
  e.d_out.ensure((size_t)H * W * sizeof(float));
  FSR_CUDA(cudaMemcpyAsync(e.d_tiles.p, tiles, tile_px * ny * nx * sizeof(float), cudaMemcpyHostToDevice, s));
  BlendGeom g = e.blend_geom();
  launch_blend(e.d_tiles.as<float>(), 0, ny, g, 0, H, nullptr, 0, true, max_depth, e.d_out.as<float>(), s);
  FSR_CUDA(cudaMemcpyAsync(out_sr, e.d_out.p, (size_t)H * W * sizeof(float), cudaMemcpyDeviceToHost, s));
  FSR_CUDA(cudaStreamSynchronize(s));
  FSR_API_END()
}

int fsr_profile_enable(fsr_engine* eng, int32_t on) {
  FSR_API_BEGIN(eng)
  eng->impl.prof.reset();
  eng->impl.prof.on = on != 0;
  FSR_API_END()
}

int fsr_profile_fetch(fsr_engine* eng, double* out_ms, int64_t* out_count, int32_t n_cat) {
  FSR_API_BEGIN(eng)
  FSR_REQUIRE(out_ms && out_count && n_cat >= PROF_NCAT, "need room for all profile categories");
  eng->impl.prof.collect();
  for (int i = 0; i < PROF_NCAT; ++i) {
    out_ms[i] = eng->impl.prof.ms[i];
    out_count[i] = eng->impl.prof.count[i];
  }
  FSR_API_END()
}

int fsr_debug_tensor_shape(fsr_engine* eng, int32_t tensor, int32_t* h, int32_t* w, int32_t* c) {
  FSR_API_BEGIN(eng)
  FSR_REQUIRE(tensor >= 0 && tensor < eng->impl.n_tensors(), "tensor index out of range");
  const fsr_tensor_desc d = eng->impl.tensor_desc(tensor);
  if (h) *h = d.h;
  if (w) *w = d.w;
  if (c) *c = d.c;
  FSR_API_END()
}

int fsr_debug_read_tensor(fsr_engine* eng, int32_t tensor, int32_t n_tiles, float* out) {
  FSR_API_BEGIN(eng)
  Engine& e = eng->impl;
  FSR_REQUIRE(tensor >= 0 && tensor < e.n_tensors() && n_tiles > 0 && out, "bad tensor index or buffer");
  const fsr_tensor_desc d = e.tensor_desc(tensor);
  const size_t n = (size_t)n_tiles * d.h * d.w * d.c;
  e.d_tmp_b.ensure(n * sizeof(float));
  e.debug_read_tensor(tensor, n_tiles, e.d_tmp_b.as<float>(), 0);
  FSR_CUDA(cudaMemcpy(out, e.d_tmp_b.p, n * sizeof(float), cudaMemcpyDeviceToHost));
  FSR_API_END()
}

void* fsr_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void fsr_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

}  // extern "C"
