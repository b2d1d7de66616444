/*
 * floodsr_b200 -- C ABI of the B200 engine backend for floodsr's ToHR hot path.
 *
 * The reference (cefect/floodsr) is pure Python and has no FFI of its own: its engine boundary is the
 * Python ABC `EngineBase` (floodsr/engine/base.py:10-29) implemented by `EngineORT`
 * (floodsr/engine/ort.py:28-208), whose only native call is `onnxruntime.InferenceSession.run`
 * (ort.py:193).  This header is what the replacement backend binds instead (via ctypes, see
 * floodsr_b200/_lib.py and INTEGRATION.md).  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *  - plain C types only; all arrays are dense row-major float32 unless stated.
 *  - every function returns 0 on success or a negative FSR_E_* code; fsr_last_error() gives the message
 *    of the last failure on the calling thread.  FSR_E_ASSERT marks conditions for which the reference
 *    raises AssertionError (non-finite input, flat non-zero DEM tile, shape mismatch ...): the Python
 *    binding re-raises them as AssertionError with the reference's wording.
 *  - "host" entry points take caller-owned host buffers and perform the H2D/D2H copies themselves;
 *    "_dev" entry points take device pointers (memory owned by the caller, e.g. torch tensors) and a
 *    CUDA stream handle (cudaStream_t cast to void*; NULL = default stream) and do not synchronise.
 *  - there is no CPU fallback: without a CUDA device fsr_create fails.
 */
#ifndef FLOODSR_B200_H
#define FLOODSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSR_ABI_VERSION 1

/* error codes */
#define FSR_OK 0
#define FSR_E_INVALID -1 /* bad argument / malformed plan                         */
#define FSR_E_CUDA -2    /* CUDA runtime or driver error                            */
#define FSR_E_ASSERT -3  /* a condition the reference reports with AssertionError   */
#define FSR_E_UNSUPPORTED -4

/* precision modes of the network forward pass */
#define FSR_PREC_FP32 0 /* fp32-tolerance mode (<=1e-4 m vs the fp32 oracle) on tcgen05 tensor cores: every activation and
                           weight is a split fp16 pair (hi, lo), three MMAs per product, fp32 accumulate in TMEM      */
/* 1 was bf16 operands: measured 2.4e-2 m on the H1 graph, outside the <=1e-2 m bound of the 16-bit mode; retired, fsr_create rejects it */
#define FSR_PREC_FP16 2 /* the 16-bit tensor-core mode (<=1e-2 m, same wet/dry mask at 0.01 m): tcgen05 kind::f16, fp16
                           operands and activations, fp32 accumulate in TMEM                                            */
#define FSR_PREC_FP32_SIMT 3 /* diagnostic: plain fp32 FMA on the CUDA cores (independent check of the tensor-core modes) */

/* window methods of the tile loop (floodsr/models/ResUNet_16x_DEM.py:297, :315) */
#define FSR_WINDOW_HARD 0
#define FSR_WINDOW_FEATHER 1

/* bits of the per-call assertion report (fsr_run_* out_flags, may be NULL) */
#define FSR_FLAG_DEPTH_NONFINITE 1u /* ort.py:155  */
#define FSR_FLAG_DEM_NONFINITE 2u   /* ort.py:156  */
#define FSR_FLAG_DEM_FLAT_NONZERO 4u /* preprocessing.py:82 "DEM range must be > 0" */
#define FSR_FLAG_DEPTH_NOT_UNIT 8u  /* ort.py:171  (normalize_inputs=False) */
#define FSR_FLAG_DEM_NOT_UNIT 16u   /* ort.py:174  */
#define FSR_FLAG_PRED_NONFINITE 32u /* the network produced Inf / NaN (a 16-bit activation overflowed, or the weights hold
                                       non-finite values): the clip of preprocessing.py:161 would otherwise hide it as 0 m */

typedef struct fsr_engine fsr_engine;

/* Per-call preprocessing parameters: the keyword arguments of EngineORT.run_tile (ort.py:128-138). */
typedef struct fsr_tile_params {
  float max_depth;          /* D; depth scaled by log1p(clip(x,0,D))/log1p(D) (preprocessing.py:141-151)  */
  float depth_denom;        /* float32(log1p(D)), computed by the host exactly as preprocessing.py:129-138 */
  int32_t normalize_inputs; /* 0: inputs are already in [0,1] (ort.py:163-180)                            */
  int32_t has_depth_nodata; /* replace_nodata_with_zero (preprocessing.py:167-172)                        */
  float depth_nodata;
  float depth_nodata_tol;   /* float32(atol + rtol*|nodata|) of np.isclose, <0: exact-equality only       */
  int32_t has_dem_nodata;
  float dem_nodata;
  float dem_nodata_tol;
  /* np.nanpercentile(x, pct) in numpy's float32 'linear' method (preprocessing.py:117), resolved on the
     host: p = lerp(sorted[rank_lo], sorted[rank_hi], gamma)                                             */
  int32_t rank_lo;
  int32_t rank_hi;
  float gamma;
  float dem_pct_clip;       /* only echoed into stats when normalize_inputs == 0 */
  /* dem_ref_stats of run_tile (ort.py:133, preprocessing.py:122-123): when set, every tile is normalised
     with these stats instead of its own (the worker always passes None, ResUNet_16x_DEM.py:274) */
  int32_t has_ref_stats;
  float ref_p_clip;
  float ref_dem_min;
  float ref_dem_max;
} fsr_tile_params;

/* Bilinear change of grid (SURVEY.md section 8 f#2): the two rasterio.warp.reproject(..., Resampling.bilinear) calls of the
 * reference, floodsr/preprocessing.py:371-387 (raw DEM grid -> model grid) and
 * floodsr/models/ResUNet_16x_DEM.py:554-573 (prediction -> raw DEM grid), for north-up grids in one CRS.
 * Source coordinate of destination column i: ((x_a_dst * (i + 0.5) + x_c_dst) - x_c_src) / x_a_src with the
 * (a, c) / (e, f) terms of the two affine transforms; rows likewise. */
typedef struct fsr_resample_params {
  double x_a_dst, x_c_dst, x_a_src, x_c_src;
  double y_a_dst, y_c_dst, y_a_src, y_c_src;
  int32_t has_src_nodata;   /* source pixels equal to src_nodata are skipped and the weights renormalised */
  float src_nodata;
  float dst_fill;           /* value of destination pixels nothing maps to (dst_nodata, 0 when there is none) */
} fsr_resample_params;

/* ---- lifecycle -------------------------------------------------------------------------------------
 * Replaces EngineORT.__init__/load (ort.py:31-59): `plan` is the lowered layer list and `weights` the
 * float32 initializer blob produced by floodsr_b200/graph.py from model_infer.onnx. */
int fsr_abi_version(void);
const char* fsr_last_error(void);
int fsr_device_count(void);
int fsr_create(const void* plan, size_t plan_bytes, const float* weights, size_t weights_count, int device,
               int precision, fsr_engine** out);
/* Replaces EngineORT.close (ort.py:61-64). */
int fsr_destroy(fsr_engine* eng);
/* LR/HR tile edge and scale of the loaded graph (ModelIOContract, ort.py:15-25). */
int fsr_contract(const fsr_engine* eng, int32_t* lr_tile, int32_t* hr_tile, int32_t* scale);
/* Number of kernel launches issued by this engine since creation (bench.py "gpu_launches"). */
int64_t fsr_launch_count(const fsr_engine* eng);
/* Multiply-accumulates per tile of the loaded graph's conv/convT layers (roofline numerator / 2). */
int64_t fsr_macs_per_tile(const fsr_engine* eng);

/* ---- EngineORT.run_tile for a batch of independent tiles (ort.py:128-208), host buffers ------------
 * depth_lr [B, lr, lr], dem_hr [B, hr, hr] -> out_pred_m [B, hr, hr] (metres), out_pred_norm [B, hr, hr]
 * (may be NULL), out_stats [B, 3] = (p_clip, dem_min, dem_max) (may be NULL). */
int fsr_run_tiles(fsr_engine* eng, const float* depth_lr, const float* dem_hr, int32_t n_tiles,
                  const fsr_tile_params* params, float* out_pred_m, float* out_pred_norm, float* out_stats,
                  uint32_t* out_flags);

/* ---- ModelWorker._run_tiled_model_on_prepared (ResUNet_16x_DEM.py:140-393), host buffers -----------
 * depth_lr [H/scale, W/scale] and dem_hr [H, W] are the prepared model-space rasters; padding to whole
 * tiles (:215-235) is virtual.  `y_starts/x_starts` are the window origins and `ramp` the feather weights
 * (floodsr/tiling.py), computed by the host with the reference's own arithmetic; ramp is NULL for
 * FSR_WINDOW_HARD.  out_sr [H, W] is the stitched, cropped, clipped depth in metres (:391).
 * out_stats [ny*nx, 3] (may be NULL) receives per-tile DEM stats in window order (:284-291). */
int fsr_run_raster(fsr_engine* eng, const float* depth_lr, const float* dem_hr, int32_t H, int32_t W,
                   int32_t window_method, int32_t overlap_hr, const int32_t* y_starts, int32_t ny,
                   const int32_t* x_starts, int32_t nx, const float* ramp, const fsr_tile_params* params,
                   float* out_sr, float* out_stats, uint32_t* out_flags);

/* ---- device-resident, band-wise variants --------------------------------------------------------------
 * Inputs already in HBM (memory owned by the caller, e.g. torch tensors): used for the kernel-only bench
 * figure and for row-band sharding across GPUs, where each rank runs a contiguous block
 * of window rows [ty0, ty1) of the same global window grid and NCCL moves only the halo rows.
 *
 *   fsr_set_windows        upload the global window grid (same arguments as fsr_run_raster).
 *   fsr_band_geometry      rows owned by the band: output rows [row0, row0+n_rows); halo_out_rows = rows just
 *                          below them that this band's windows also touch (owned by the next band);
 *                          [in_row0, in_row0+in_rows) = raster rows the band's windows read.
 *   fsr_band_run_dev       normalise + forward + invert every window of the band (predictions stay inside
 *                          the engine) and write this band's partial sums sum(pred*w) for the halo rows to
 *                          d_halo_out [halo_out_rows, W] (may be NULL when halo_out_rows == 0).
 *                          d_depth_lr / d_dem_hr hold raster rows starting at HR row band_row0
 *                          (band_row0 % scale == 0, band_row0 <= in_row0), band_rows_hr HR rows of them.
 *   fsr_band_finalize_dev  blend the band's own rows into d_out_rows [n_rows, W]: the window sums start from
 *                          the previous band's partial sums d_halo_in [halo_rows_in, W] (NULL for the first
 *                          band) and continue in the reference's window order, so the sharded result is
 *                          bit-identical to the single-GPU one; weights are analytic and never exchanged.
 *   fsr_band_finalize_rows_dev  the same for the band-relative rows [row_begin, row_end) only (d_out_rows still points
 *                          at the band's first row): rows >= halo_rows_in do not depend on the previous band, so a rank
 *                          can blend them while the halo exchange is in flight and the first halo_rows_in rows after it.
 * With ty0 = 0, ty1 = ny the sequence equals fsr_run_raster without the host copies. */
int fsr_set_windows(fsr_engine* eng, int32_t H, int32_t W, int32_t window_method, int32_t overlap_hr,
                    const int32_t* y_starts, int32_t ny, const int32_t* x_starts, int32_t nx, const float* ramp,
                    void* stream);
int fsr_band_geometry(fsr_engine* eng, int32_t ty0, int32_t ty1, int32_t* row0, int32_t* n_rows,
                      int32_t* halo_out_rows, int32_t* in_row0, int32_t* in_rows);
int fsr_band_run_dev(fsr_engine* eng, const float* d_depth_lr, const float* d_dem_hr, int32_t band_row0,
                     int32_t band_rows_hr, int32_t ty0, int32_t ty1, const fsr_tile_params* params,
                     float* d_halo_out, float* d_stats, void* stream);
int fsr_band_finalize_dev(fsr_engine* eng, const float* d_halo_in, int32_t halo_rows_in, float* d_out_rows,
                          void* stream);
int fsr_band_finalize_rows_dev(fsr_engine* eng, const float* d_halo_in, int32_t halo_rows_in, float* d_out_rows,
                               int32_t row_begin, int32_t row_end, void* stream);
/* Host-buffer, pipelined variant for one rank's band (new; used by bench.py's multi-GPU end-to-end leg and by
 * floodsr_b200/dist.py): depth_lr / dem_hr / out_rows are (page-locked) host buffers holding the raster rows from HR row
 * band_row0 on and the rows the band owns; sub-bands overlap H2D, kernels and D2H like fsr_run_raster.
 *   fsr_band_host_begin  queues everything except the blend of the rows shared with the previous rank, writes the rows
 *                        shared with the next rank as partial sums to d_halo_out (device, may be NULL for the last rank)
 *                        and returns when those are complete.
 *   fsr_band_host_end    after the caller has received the previous rank's partial sums into d_halo_in (device, NULL for
 *                        the first rank): blends the deferred rows, waits for all copies, reports the assertion flags. */
int fsr_band_host_begin(fsr_engine* eng, const float* depth_lr, const float* dem_hr, int32_t band_row0, int32_t band_rows_hr,
                        int32_t ty0, int32_t ty1, const fsr_tile_params* params, float* out_rows, float* d_halo_out);
int fsr_band_host_end(fsr_engine* eng, const float* d_halo_in, int32_t halo_rows_in, uint32_t* out_flags);
/* Poll and clear the assertion flags raised by _dev calls (synchronises `stream`). */
int fsr_fetch_flags(fsr_engine* eng, void* stream, uint32_t* out_flags);

/* ---- stage-level entry points (parity tests of the memory-bound kernels, host buffers) --------------
 * a5-a8: nodata->0, finite checks, log1p depth scaling, tile-local DEM stats + normalisation
 * (preprocessing.py:97-172).  out_dem_norm [B,hr,hr], out_depth_norm [B,lr,lr], out_stats [B,3]. */
int fsr_stage_normalize(fsr_engine* eng, const float* depth_lr, const float* dem_hr, int32_t n_tiles,
                        const fsr_tile_params* params, float* out_depth_norm, float* out_dem_norm,
                        float* out_stats, uint32_t* out_flags);
/* a10 only: network forward on already-normalised inputs -> out_pred_norm [B,hr,hr]. */
int fsr_stage_forward(fsr_engine* eng, const float* depth_norm, const float* dem_norm, int32_t n_tiles,
                      float* out_pred_norm);
/* a11: invert_depth_log1p_np (preprocessing.py:154-164) over n floats. */
int fsr_stage_invert(fsr_engine* eng, const float* pred_norm, size_t n, float max_depth, float depth_denom,
                     float* out_pred_m);
/* a15 mosaic only (ResUNet_16x_DEM.py:297-363, :391): tiles [ny*nx, hr, hr] in window order -> out_sr. */
int fsr_stage_blend(fsr_engine* eng, const float* tiles, int32_t H, int32_t W, int32_t window_method,
                    int32_t overlap_hr, const int32_t* y_starts, int32_t ny, const int32_t* x_starts,
                    int32_t nx, const float* ramp, float max_depth, float* out_sr);

/* ---- per-stage device timing (bench.py roofline) -------------------------------------------------------
 * When enabled, CUDA events bracket every launch group on its stream; fsr_profile_fetch synchronises and
 * returns, per category, the summed device time in ms and the number of groups since the last enable.
 * Categories: 0 tile normalisation, 1 LR convs, 2 LR pool/upsample/eltwise, 3 transposed conv, 4 head conv,
 * 5 log1p inversion, 6 blend. */
#define FSR_PROF_NCAT 7
int fsr_profile_enable(fsr_engine* eng, int32_t on);
int fsr_profile_fetch(fsr_engine* eng, double* out_ms, int64_t* out_count, int32_t n_cat);

/* ---- debugging: read an intermediate activation of the last forward pass as NHWC float32 ------------------
 * (first n_tiles tiles of the last chunk; 16-bit and split CP8 tensors are converted).  Used by the layer-by-layer error
 * budget of the precision modes (tests/x3_error_budget.py). */
int fsr_debug_tensor_shape(fsr_engine* eng, int32_t tensor, int32_t* h, int32_t* w, int32_t* c);
int fsr_debug_read_tensor(fsr_engine* eng, int32_t tensor, int32_t n_tiles, float* out);

/* Pinned host memory for callers that want full-speed H2D/D2H through the host entry points. */
void* fsr_host_alloc(size_t bytes);
void fsr_host_free(void* p);

/* ---- grid change (section 8 f#2) ---------------------------------------------------------------------
 * fsr_resample_bilinear      host buffers: src [sh, sw] -> dst [dh, dw] (copies in, kernel, copy out).
 * fsr_resample_bilinear_dev  device buffers on `stream`, for callers that keep the rasters resident. */
int fsr_resample_bilinear(fsr_engine* eng, const float* src, int32_t sh, int32_t sw, float* dst, int32_t dh, int32_t dw,
                          const fsr_resample_params* params);
int fsr_resample_bilinear_dev(fsr_engine* eng, const float* d_src, int32_t sh, int32_t sw, float* d_dst, int32_t dh,
                              int32_t dw, const fsr_resample_params* params, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOODSR_B200_H */
