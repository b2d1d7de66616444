#!/usr/bin/env python
"""Benchmark of the ToHR hot path on B200: hires megapixels per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp16|bf16|fp32] [--impl reference]

One "step" = one full pass pad -> tile -> normalise -> ResUNet_16x_DEM forward -> invert -> overlap-blend
stitch -> clip over a synthetic model-space raster.  Workload (weak scaling): a 4096-row x 32768-column band
of hires raster per GPU with the reference's default feather windows (overlap 8 LR px); at N = 8 this is
BASELINE config 5 (32k x 32k sharded in row bands, halo rows exchanged between neighbouring ranks).

`value`   device-resident inputs/outputs, CUDA-event timed, max over ranks.
`e2e`     same pass through the public API from pinned host buffers, H2D/D2H inside the timed region.
`roofline` the dominant kernel (fused head conv) against the measured bf16/fp16 dense tensor peak.
`cpu_baseline` / `--impl reference`: the CPU oracle (torch-CPU restatement of the ONNX Runtime path; onnxruntime
and the model asset are not available offline) on a bounded sample of the same workload, all host threads.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))
# stdout carries exactly one JSON line: keep NCCL's own banner / debug output on stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "hires_mpx_per_s"
UNIT = "Mpx/s"
ROWS_PER_GPU = 4096
WIDTH = 32768
OVERLAP_LR = 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("FLOODSR_B200_PRECISION", "fp16"), choices=["fp32", "bf16", "fp16"])
    ap.add_argument("--rows-per-gpu", type=int, default=ROWS_PER_GPU)
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def find_model(tmpdir: Path) -> tuple[Path, str]:
    """The reference's lookup order (tests/conftest.py:200-215): ./_inputs/<ver>/*.onnx, then the user cache; else random init."""
    from floodsr_b200.h1 import write_h1_model
    from floodsr_b200.model_store import find_model as locate_model

    ver = "ResUNet_16x_DEM"
    try:
        hit = locate_model(ver, inputs_dir=REPO / "_inputs")
    except ValueError:
        hit = None  # a cached file that is not the release asset
    if hit is not None and hit.stat().st_size > 1_000_000:
        return hit, "model_infer.onnx (release asset)"
    return write_h1_model(tmpdir / ver / "model_infer.onnx", seed=0), "random-init H1 graph (12,045,568 parameters; release asset unavailable offline)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.device_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        assert self.proc is not None and self.proc.stdout is not None
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def head_traffic_per_launch(flops_per_launch: float, fused: bool = True):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed ncu capture, scaled to this launch size."""
    fp = REPO / "profiles" / ("fused_hr_kernel_ncu.json" if fused else "head_kernel_ncu.json")
    if not fp.exists():
        return None
    rec = json.loads(fp.read_text())
    try:
        per_flop = (rec["dram_bytes_read"] + rec["dram_bytes_write"]) / rec["flops_per_launch"]
        return per_flop * flops_per_launch
    except (KeyError, ZeroDivisionError):
        return None


def cpu_reference_rate(model_fp: Path, sample_hw=(2048, 3072), steps: int = 1, warmup: int = 0) -> dict:
    """Oracle (CPU restatement of the reference path) on a bounded sample raster: hires Mpx/s."""
    import torch

    from floodsr_b200.synth import synth_raster
    from oracle.engine_ref import OracleEngine
    from oracle.stitch_np import run_tiled

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    eng = OracleEngine(model_fp, threads=cores)
    h, w = sample_hw
    depth, dem = synth_raster(h, w, seed=1)
    for _ in range(warmup):
        run_tiled(eng, depth, dem, overlap_lr=OVERLAP_LR)
    times = []
    n_tiles = 0
    for _ in range(max(steps, 1)):
        t0 = time.perf_counter()
        out, n_tiles, _ = run_tiled(eng, depth, dem, overlap_lr=OVERLAP_LR)
        times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    return {
        "value": h * w / 1e6 / dt,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{h}x{w} hires raster, {n_tiles} feather windows, oracle tile loop + stitch (torch-CPU fp32 restatement of the ORT path, "
                  f"{dt:.2f} s/pass, {dt / max(n_tiles, 1) * 1e3:.0f} ms/tile); onnxruntime is not installed here",
        "ms_per_step": dt * 1e3,
    }


def run_reference(args, stdout_fd: int) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    with tempfile.TemporaryDirectory() as td:
        model_fp, model_desc = find_model(Path(td))
        cb = cpu_reference_rate(model_fp, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": cb["value"],
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.rows_per_gpu}x{args.width} hires rows per GPU, feather windows overlap_lr={OVERLAP_LR}; "
                               f"reference arm runs a bounded {cb['sample']}", "model": model_desc},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line, stdout_fd)


def emit(line: dict, stdout_fd: int) -> None:
    """Print the ONE JSON line on the real stdout (everything else a library writes to fd 1 was sent to stderr)."""
    sys.stdout.flush()
    os.dup2(stdout_fd, 1)
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    # libraries (NCCL's version banner, warnings) write to fd 1: park the real stdout and point fd 1 at stderr until the
    # result line is printed
    sys.stdout.flush()
    stdout_fd = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, stdout_fd)
        return

    import torch
    import torch.distributed as dist

    from floodsr_b200 import _lib
    from floodsr_b200.dist import CudaBandExecutor, plan_bands, run_band_step, run_band_step_host
    from floodsr_b200.engine import EngineB200
    from floodsr_b200.synth import synth_dem, synth_depth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: floodsr_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    tmp = tempfile.TemporaryDirectory()
    model_fp, model_desc = find_model(Path(tmp.name))
    eng = EngineB200(model_fp, precision=args.precision, device=local_rank)
    scale, hr_tile = eng.contract.scale, eng.contract.dem_hr_hwc[0]

    H, W = args.rows_per_gpu * world, args.width
    overlap_hr = OVERLAP_LR * scale
    plans, ys, xs = plan_bands(H, W, hr_tile, "feather", overlap_hr, world)
    plan = plans[rank]
    n_tiles_total = len(ys) * len(xs)
    n_tiles_mine = (plan.ty1 - plan.ty0) * len(xs)

    # synthetic band inputs in pinned host memory (only the rows this rank's windows read)
    r0, rows = plan.in_row0, plan.in_rows
    lr0, lr_rows = r0 // scale, (r0 + rows + scale - 1) // scale - r0 // scale
    lr_rows = min(lr_rows, H // scale - lr0)
    h_dem = _lib.pinned_empty((rows, W))
    h_depth = _lib.pinned_empty((lr_rows, W // scale))
    blk = 1024
    for y in range(0, rows, blk):
        n = min(blk, rows - y)
        h_dem[y : y + n] = synth_dem(n, W, seed=7, y0=r0 + y)
    h_depth[:] = synth_depth(lr_rows, W // scale, seed=7, y0=lr0)
    h_out = _lib.pinned_empty((plan.n_rows, W))
    t_dem_host = torch.from_numpy(h_dem)
    t_depth_host = torch.from_numpy(h_depth)
    t_out_host = torch.from_numpy(h_out)

    ex = CudaBandExecutor(eng, H, W, "feather", overlap_hr)
    d_dem = t_dem_host.to(dev, non_blocking=True)
    d_depth = t_depth_host.to(dev, non_blocking=True)
    d_out = torch.empty((plan.n_rows, W), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    dist_mod = dist if world > 1 else None

    def step_device():
        run_band_step(ex, plan, plans, d_depth, d_dem, r0, dist_mod, None, d_out)

    def step_e2e():
        if world == 1:
            # the call a user makes: host arrays in, host array out (H2D + D2H inside)
            eng.run_raster(h_depth, h_dem, window_method="feather", overlap_lr=OVERLAP_LR, out=h_out)
        else:
            # each rank: its rows from pinned host memory, through the engine's copy/compute pipeline, halo via NCCL
            run_band_step_host(ex, plan, plans, h_depth, h_dem, r0, h_out, dist_mod, None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident leg ----------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    ex.check_flags()
    launches0 = eng.launch_count()
    eng.profile(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else {}
    prof = eng.profile_fetch()
    eng.profile(False)
    launches = torch.tensor([eng.launch_count() - launches0], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    ms_per_step = ms_total / args.steps
    mpx = H * W / 1e6
    value = mpx / (ms_per_step / 1e3)

    # ---- head kernel alone (no convT running beside it): explains the in-step roofline figure ------------
    head_isolated = None
    if args.precision in ("fp16", "bf16") and os.environ.get("FSR_NO_FUSED_HR"):
        os.environ["FSR_HR_OVERLAP"] = "0"
        eng_iso = EngineB200(model_fp, precision=args.precision, device=local_rank)
        del os.environ["FSR_HR_OVERLAP"]
        ex_iso = CudaBandExecutor(eng_iso, H, W, "feather", overlap_hr)
        for _ in range(2):
            run_band_step(ex_iso, plan, plans, d_depth, d_dem, r0, None, None, d_out)
        torch.cuda.synchronize()
        eng_iso.profile(True)
        run_band_step(ex_iso, plan, plans, d_depth, d_dem, r0, None, None, d_out)
        torch.cuda.synchronize()
        head_isolated = eng_iso.profile_fetch()["head"]
        eng_iso.close()
        del ex_iso, eng_iso

    # ---- end-to-end leg ---------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        e2e_ms_dev = timed(step_e2e, args.steps)
        wall = (time.perf_counter() - t0) * 1e3
        e2e_ms = max(e2e_ms_dev, wall) / args.steps if world == 1 else e2e_ms_dev / args.steps
        h2d = (h_dem.nbytes + h_depth.nbytes)
        d2h = h_out.nbytes
        tot = torch.tensor([h2d, d2h], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        e2e = {"value": mpx / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
               "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": e2e_ms}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused head conv; tensor-bound) -------------------------------
    peaks = {}
    pk = REPO / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    head_ms, head_launches = prof["head"]
    lm = eng.lowered
    head_op = lm.ops[-1]
    hh, hw, _ = lm.tensors[head_op.dst]
    cin_head = lm.tensors[head_op.src0][2] + 1
    head_macs_tile = hh * hw * (head_op.k * head_op.k * cin_head * head_op.cout + head_op.cout)
    fused = prof["convt"][1] == 0  # the fused kernel also does the transposed convolution
    if fused:
        ct_op = lm.ops[-2]
        head_macs_tile += hh * hw * lm.tensors[ct_op.src0][2] * ct_op.cout
    tiles_timed = n_tiles_mine * args.steps
    head_flops = 2.0 * head_macs_tile * tiles_timed
    achieved_tf = head_flops / (head_ms / 1e3) / 1e12 if head_ms > 0 else 0.0
    # fp32-tolerance mode: three fp16 MMAs per product (split operands) -> its ceiling is a third of the dense 16-bit peak
    mma_per_product = 3 if args.precision == "fp32" else 1
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0)) / mma_per_product
    roofline = {
        "bound": "tensor",
        "kernel": ("fused_hr_x3_kernel (split fp16 operands, 3 MMAs per product: convT16x16+act -> conv3x3+DEM+act -> conv1x1 -> expm1, tcgen05)"
                   if args.precision == "fp32" else
                   "fused_hr_kernel (convT16x16+act -> conv3x3+DEM+act -> conv1x1 -> expm1, tcgen05)") if fused else
                  "head_tc_kernel (fused conv3x3+DEM+act+conv1x1+expm1, tcgen05)",
        "achieved": achieved_tf,
        "peak": peak_tf,
        "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf,
        "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained")
                       + (" / 3 MMAs per product" if mma_per_product == 3 else ""),
        "flops_per_launch": head_flops / max(head_launches, 1),
        "ms_per_launch": head_ms / max(head_launches, 1),
        "launches": head_launches,
        "traffic": head_traffic_per_launch(head_flops / max(head_launches, 1), fused),
        "note": "timed inside the step with CUDA events around each launch; FLOPs = transposed convolution + head of the tiles in the launch",
    }
    if not fused and head_isolated is not None and head_isolated[0] > 0:
        iso_tf = 2.0 * head_macs_tile * n_tiles_mine / (head_isolated[0] / 1e3) / 1e12
        roofline["isolated"] = {"achieved": iso_tf, "frac": iso_tf / peak_tf, "ms_per_launch": head_isolated[0] / max(head_isolated[1], 1),
                                "launches": head_isolated[1]}
    stage_ms = {k: round(v[0] / args.steps, 3) for k, v in prof.items()}
    hbm_gbs = float(peaks.get("hbm_gbs", 6650.0))
    # memory-bound stages: algorithmic bytes per step on this rank
    blend_bytes = plan.n_rows * W * 4 * (1 + 1.78)
    norm_bytes = n_tiles_mine * (hr_tile * hr_tile * 4 * 2 + 32 * 32 * 4 * 2)
    mem_stages = {
        "normalize": {"GBps": norm_bytes * args.steps / (prof["normalize"][0] / 1e3) / 1e9 if prof["normalize"][0] else None},
        "blend": {"GBps": blend_bytes * args.steps / (prof["blend"][0] / 1e3) / 1e9 if prof["blend"][0] else None},
        "hbm_peak_GBps": hbm_gbs,
    }

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_rate(model_fp, steps=2, warmup=1)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": args.precision,
        "data": "synthetic",
        "config": {
            "workload": f"synthetic {H}x{W} hires raster ({H // scale}x{W // scale} lores), {args.rows_per_gpu} rows per GPU in row bands, "
                        f"{n_tiles_total} feather windows (overlap_lr={OVERLAP_LR}), halo rows exchanged between neighbour ranks",
            "model": model_desc,
            "tiles_per_step": n_tiles_total,
            "gflop_per_tile": 2 * eng.macs_per_tile() / 1e9,
            "l2_policy": "inputs larger than L2 (512 MiB DEM band per GPU vs 126 MB L2), no explicit flush",
            "precision": f"{args.precision} operands, fp32 accumulate (tcgen05 kind::f16)" if args.precision != "fp32" else
                         "fp32 tolerance (<= 1e-4 m): split fp16 (hi, lo) operands, 3 tcgen05 kind::f16 MMAs per product, fp32 accumulate in TMEM",
            "parallelism": f"row-bands x{world}",
        },
        "roofline": roofline,
        "stage_ms_per_step_rank0": stage_ms,
        "memory_bound_stages": mem_stages,
        "cpu_baseline": cpu_baseline,
        "e2e": e2e,
        "gpu_launches": int(launches.item()),
        "clocks": clocks,
    }
    emit(line, stdout_fd)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
