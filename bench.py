#!/usr/bin/env python
"""Benchmark of the ToHR hot path on B200: hires megapixels per second (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|fp16] [--config weak|c3|c4] [--impl reference]

One "step" = one full pass pad -> tile -> normalise -> ResUNet_16x_DEM forward -> invert -> overlap-blend
stitch -> clip over a synthetic model-space raster.  Default workload (weak scaling): a 4096-row x 32768-column band
of hires raster per GPU with the reference's default feather windows (overlap 8 LR px); at N = 8 this is
BASELINE config 5 (32k x 32k sharded in row bands, halo rows exchanged between neighbouring ranks).  `--config c3` /
`c4` measure BASELINE configs 3 (256 independent tiles) and 4 (8k x 8k raster, strong scaling over 1/2/4 GPUs).

The headline (`value`, `dtype: fp32`) is the fp32-tolerance mode (<= 1e-4 m vs the reference's fp32 CPU path), which runs
on tcgen05 with split fp16 operands; the 16-bit mode (<= 1e-2 m, "stated separately" in the north star) is measured in
the same run and reported under `modes.fp16`.

`value`   device-resident inputs/outputs, CUDA-event timed, max over ranks.
`e2e`     same pass through the public API from pinned host buffers, H2D/D2H inside the timed region; `copy_ceiling_ms`
          is the measured time this step's bytes need when nothing else runs (all ranks copying at once).
`roofline` the dominant kernel (fused convT + head) against the measured dense 16-bit tensor peak (/ 3 in fp32 mode).
`cpu_baseline` / `--impl reference`: the reference's CPU path on a bounded sample of the same workload, all host threads:
a live onnxruntime session when one is importable (kind "reference"), else the torch-CPU restatement (kind "port").
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
    ap.add_argument("--precision", default=os.environ.get("FLOODSR_B200_PRECISION", "fp32"), choices=["fp32", "fp16"])
    ap.add_argument("--config", default="weak", choices=["weak", "c3", "c4"],
                    help="weak: 4096 x 32768 rows per GPU (N = 8: BASELINE config 5); c3: 256 independent tiles; c4: 8192 x 8192 raster")
    ap.add_argument("--no-modes", action="store_true", help="skip the secondary fp16 figures under `modes`")
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


def bind_to_gpu_numa_node(device_index: int):
    """Run this process (its pinned allocations, the copy-engine doorbells, the NCCL proxy) on the CPUs of the GPU's NUMA
    node.  Best effort: silently skipped when the topology files are not readable (containers without /sys access).
    Returns the affinity mask to restore for CPU-side work (the cpu_baseline leg uses every allowed core)."""
    before = os.sched_getaffinity(0)
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node_fp = Path("/sys/bus/pci/devices") / bus.lower()[-12:] / "numa_node"
        node = int(node_fp.read_text().strip())
        if node < 0:
            return before
        cpus = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
        ids: set[int] = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        if ids & allowed:
            os.sched_setaffinity(0, ids & allowed)
    except Exception:
        pass
    return before


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


def cpu_reference_rate(model_fp: Path, sample_hw=(2048, 3072), steps: int = 1, warmup: int = 0) -> dict:
    """Oracle (CPU restatement of the reference path) on a bounded sample raster: hires Mpx/s."""
    import torch

    from floodsr_b200.synth import synth_raster
    from oracle.engine_ref import OracleEngine
    from oracle.stitch_np import run_tiled

    cores = len(os.sched_getaffinity(0)) or os.cpu_count() or 1
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
    live = eng.backend.startswith("onnxruntime")
    return {
        "value": h * w / 1e6 / dt,
        "unit": UNIT,
        "cores": cores,
        "kind": "reference" if live else "port",
        "sample": f"{h}x{w} hires raster, {n_tiles} feather windows, tile loop + stitch of floodsr/models/ResUNet_16x_DEM.py:140-393 "
                  f"over {eng.backend}, {dt:.2f} s/pass, {dt / max(n_tiles, 1) * 1e3:.0f} ms/tile",
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
        "scaling": "weak" if args.config == "weak" else "strong",
        "vs_baseline": None,
        "dtype": "fp32",
        "data": "synthetic",
        "config": {"workload": f"config '{args.config}' of the B200 arm (hires Mpx/s is size-normalised); reference arm runs a bounded {cb['sample']}",
                   "name": args.config, "model": model_desc},
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


def copy_ceiling(torch, dist, world, dev, t_in_host, t_out_host, d_in, d_out, reps=5):
    """Floor of one end-to-end step when it is copy bound: this step's input bytes host->device and output bytes
    device->host, from / to the same page-locked buffers, on two streams at once, all ranks at the same time (they share
    the host's memory system and PCIe root complexes).  Returns (ms, aggregate GB/s over all ranks), max over ranks."""
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    times = []
    for _ in range(reps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s_in.wait_event(e0)
        s_out.wait_event(e0)
        with torch.cuda.stream(s_in):
            d_in.copy_(t_in_host, non_blocking=True)
            e1.record(s_in)
        with torch.cuda.stream(s_out):
            t_out_host.copy_(d_out, non_blocking=True)
            e2.record(s_out)
        torch.cuda.synchronize()
        times.append(max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
    ms = torch.tensor([sorted(times)[len(times) // 2]], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    nbytes = (t_in_host.numel() * t_in_host.element_size() + t_out_host.numel() * t_out_host.element_size()) * world
    return ms, nbytes / (ms / 1e3) / 1e9


def workload(args, world, scale, hr_tile):
    """(H, W, window method, scaling, description) of the configuration being measured."""
    if args.config == "c3":
        # BASELINE config 3: 256 independent tiles = hard (non-overlapping) windows of a 256-tile-high, one-tile-wide raster
        return 256 * hr_tile, hr_tile, "hard", "strong", "BASELINE config 3: batch of 256 independent DEM-conditioned tiles (hard windows of a 131072x512 raster)"
    if args.config == "c4":
        return 8192, 8192, "feather", "strong", "BASELINE config 4: synthetic 8192x8192 hires raster (512x512 lores), feather windows overlap_lr=8, row bands over the GPUs"
    H = args.rows_per_gpu * world
    return H, args.width, "feather", "weak", (
        f"synthetic {H}x{args.width} hires raster ({H // scale}x{args.width // scale} lores), {args.rows_per_gpu} rows per GPU in row bands"
        + (" = BASELINE config 5 (32k x 32k over 8 GPUs)" if H == 32768 and args.width == 32768 else ""))


def measure_mode(args, precision, model_fp, ctx):
    """Device-resident leg, per-stage profile and end-to-end leg of one precision mode on the prepared workload."""
    import torch

    from floodsr_b200.dist import CudaBandExecutor, run_band_step, run_band_step_host
    from floodsr_b200.engine import EngineB200

    dist, world, rank, dev = ctx["dist"], ctx["world"], ctx["rank"], ctx["dev"]
    plan, plans, r0 = ctx["plan"], ctx["plans"], ctx["r0"]
    H, W, method = ctx["H"], ctx["W"], ctx["method"]
    eng = EngineB200(model_fp, precision=precision, device=ctx["local_rank"])
    ex = CudaBandExecutor(eng, H, W, method, ctx["overlap_hr"])
    dist_mod = dist if world > 1 else None
    d_depth, d_dem, d_out = ctx["d_depth"], ctx["d_dem"], ctx["d_out"]
    h_depth, h_dem, h_out = ctx["h_depth"], ctx["h_dem"], ctx["h_out"]

    def step_device():
        run_band_step(ex, plan, plans, d_depth, d_dem, r0, dist_mod, None, d_out)

    def step_e2e():
        if world == 1:
            # the call a user makes: host arrays in, host array out (H2D + D2H inside)
            eng.run_raster(h_depth, h_dem, window_method=method, overlap_lr=OVERLAP_LR, out=h_out)
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

    res = {"precision": precision}
    if not plan.empty:
        for _ in range(max(args.warmup, 3)):
            step_device()
        ex.check_flags()
    launches0 = eng.launch_count()
    eng.profile(True)
    sampler = ClockSampler(ctx["local_rank"])
    if rank == 0:
        sampler.start()
    ms_total = timed(step_device if not plan.empty else (lambda: None), args.steps)
    res["clocks"] = sampler.stop() if rank == 0 else {}
    res["prof"] = eng.profile_fetch()
    eng.profile(False)
    launches = torch.tensor([eng.launch_count() - launches0], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    res["launches"] = int(launches.item())
    res["ms_per_step"] = ms_total / args.steps
    res["e2e"] = None
    if not args.no_e2e:
        for _ in range(2):
            if not plan.empty:
                step_e2e()
        barrier()
        t0 = time.perf_counter()
        e2e_ms_dev = timed(step_e2e if not plan.empty else (lambda: None), args.steps)
        wall = (time.perf_counter() - t0) * 1e3
        res["e2e_ms"] = max(e2e_ms_dev, wall) / args.steps if world == 1 else e2e_ms_dev / args.steps
    res["macs_per_tile"] = eng.macs_per_tile()
    res["lowered"] = eng.lowered
    eng.close()
    return res


def roofline_of(args, res, peaks, n_tiles_mine):
    """Dominant kernel (the fused high-resolution kernel) against the measured dense 16-bit tensor peak."""
    prof, lm, precision = res["prof"], res["lowered"], res["precision"]
    head_ms, head_launches = prof["head"]
    head_op = lm.ops[-1]
    hh, hw, _ = lm.tensors[head_op.dst]
    cin_head = lm.tensors[head_op.src0][2] + 1
    head_macs_tile = hh * hw * (head_op.k * head_op.k * cin_head * head_op.cout + head_op.cout)
    fused = prof["convt"][1] == 0  # the fused kernel also does the transposed convolution
    if fused:
        ct_op = lm.ops[-2]
        head_macs_tile += hh * hw * lm.tensors[ct_op.src0][2] * ct_op.cout
    head_flops = 2.0 * head_macs_tile * n_tiles_mine * args.steps
    achieved_tf = head_flops / (head_ms / 1e3) / 1e12 if head_ms > 0 else 0.0
    # fp32-tolerance mode: three fp16 MMAs per product (split operands) -> its ceiling is a third of the dense 16-bit peak
    mma_per_product = 3 if precision == "fp32" else 1
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0)) / mma_per_product
    ncu = None
    fp = REPO / "profiles" / ("fused_hr_x3_kernel_ncu.json" if precision == "fp32" else "fused_hr_kernel_ncu.json")
    if fp.exists():
        ncu = json.loads(fp.read_text())
    # tensor-pipe activity of every tensor-core kernel of the step (committed ncu captures of this command, see profiles/README.md)
    by_kernel = {}
    if ncu and ncu.get("tensor_pipe_pct") is not None:
        by_kernel["fused_hr_x3_kernel" if precision == "fp32" else "fused_hr_kernel"] = {"tensor_pipe_pct": ncu["tensor_pipe_pct"]}
    fp_conv = REPO / "profiles" / "r2" / f"conv_kernels_{precision}_ncu.json"
    if fp_conv.exists():
        for k, v in json.loads(fp_conv.read_text()).get("kernels", {}).items():
            by_kernel[k] = {"tensor_pipe_pct": v["tensor_pipe_pct_time_weighted"], "range": [v["tensor_pipe_pct_min"], v["tensor_pipe_pct_max"]],
                            "launches_per_step": v["launches"]}
    traffic = None
    if ncu:
        try:
            traffic = (ncu["dram_bytes_read"] + ncu["dram_bytes_write"]) / ncu["flops_per_launch"] * (head_flops / max(head_launches, 1))
        except (KeyError, ZeroDivisionError):
            traffic = None
    return {
        "bound": "tensor",
        "kernel": ("fused_hr_x3_kernel (split fp16 operands, 3 MMAs per product: convT16x16+act -> conv3x3+DEM+act -> conv1x1 -> expm1, tcgen05)"
                   if precision == "fp32" else
                   "fused_hr_kernel (convT16x16+act -> conv3x3+DEM+act -> conv1x1 -> expm1, tcgen05)") if fused else
                  "head kernel of the unfused pair",
        "achieved": achieved_tf,
        "peak": peak_tf,
        "unit": "TFLOP/s",
        "frac": achieved_tf / peak_tf,
        "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained")
                       + (" / 3 MMAs per product; achieved counts each product once (algorithmic FLOPs)" if mma_per_product == 3 else ""),
        "flops_per_launch": head_flops / max(head_launches, 1),
        "ms_per_launch": head_ms / max(head_launches, 1),
        "launches": head_launches,
        "traffic": traffic,
        "tensor_pipe_pct": (ncu or {}).get("tensor_pipe_pct"),
        "tensor_pipe_source": (ncu or {}).get("tensor_pipe_metric"),
        "tensor_pipe_pct_by_kernel": by_kernel or None,
        "note": "timed inside the step with CUDA events around each launch; FLOPs = transposed convolution + head of the tiles in the launch",
    }


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
    from floodsr_b200.dist import plan_bands
    from floodsr_b200.graph import lower_onnx
    from floodsr_b200.synth import synth_dem, synth_depth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: floodsr_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity_before = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    tmp = tempfile.TemporaryDirectory()
    model_fp, model_desc = find_model(Path(tmp.name))
    contract = lower_onnx(model_fp).contract
    scale, hr_tile = contract.scale, contract.dem_hr_hwc[0]

    H, W, method, scaling, wl_desc = workload(args, world, scale, hr_tile)
    overlap_hr = OVERLAP_LR * scale
    plans, ys, xs = plan_bands(H, W, hr_tile, method, overlap_hr, world)
    plan = plans[rank]
    n_tiles_total = len(ys) * len(xs)
    n_tiles_mine = (plan.ty1 - plan.ty0) * len(xs)

    # synthetic band inputs in pinned host memory (only the rows this rank's windows read)
    r0, rows = plan.in_row0, max(plan.in_rows, scale)
    r0 = min(r0, H - rows)
    lr0, lr_rows = r0 // scale, (r0 + rows + scale - 1) // scale - r0 // scale
    lr_rows = min(lr_rows, H // scale - lr0)
    h_dem = _lib.pinned_empty((rows, W))
    h_depth = _lib.pinned_empty((lr_rows, W // scale))
    blk = 1024
    for y in range(0, rows, blk):
        n = min(blk, rows - y)
        h_dem[y : y + n] = synth_dem(n, W, seed=7, y0=r0 + y)
    h_depth[:] = synth_depth(lr_rows, W // scale, seed=7, y0=lr0)
    h_out = _lib.pinned_empty((max(plan.n_rows, 1), W))
    t_dem_host = torch.from_numpy(h_dem)
    t_depth_host = torch.from_numpy(h_depth)
    t_out_host = torch.from_numpy(h_out)
    d_dem = t_dem_host.to(dev, non_blocking=True)
    d_depth = t_depth_host.to(dev, non_blocking=True)
    d_out = torch.empty((max(plan.n_rows, 1), W), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    if plan.empty:
        h_out, d_out = h_out[:0], d_out[:0]

    ctx = dict(dist=dist, world=world, rank=rank, dev=dev, local_rank=local_rank, plan=plan, plans=plans, r0=plan.in_row0, H=H, W=W,
               method=method, overlap_hr=overlap_hr, d_depth=d_depth, d_dem=d_dem, d_out=d_out, h_depth=h_depth, h_dem=h_dem, h_out=h_out)

    # ---- the floor of a copy-bound end-to-end step on this box, all ranks copying at once ---------------------------------
    ceiling = None
    if not args.no_e2e:
        ceiling = copy_ceiling(torch, dist, world, dev, t_dem_host, t_out_host if not plan.empty else t_dem_host[:1], d_dem,
                               d_out if not plan.empty else d_dem[:1])

    main_res = measure_mode(args, args.precision, model_fp, ctx)
    other = {}
    if args.precision == "fp32" and not args.no_modes:
        other["fp16"] = measure_mode(args, "fp16", model_fp, ctx)

    h2d = h_dem.nbytes + h_depth.nbytes
    d2h = h_out.nbytes
    tot = torch.tensor([h2d, d2h], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = REPO / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    mpx = H * W / 1e6
    hbm_gbs = float(peaks.get("hbm_gbs", 6650.0))

    def e2e_of(res):
        if res.get("e2e_ms") is None:
            return None
        out = {"value": mpx / (res["e2e_ms"] / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(tot[0].item()),
               "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": res["e2e_ms"]}
        if ceiling is not None:
            out["copy_ceiling_ms"] = ceiling[0]
            out["copy_ceiling_GBps_all_ranks"] = ceiling[1]
            out["frac_of_copy_ceiling"] = ceiling[0] / res["e2e_ms"]
            out["copy_ceiling_note"] = ("this step's bytes moved with nothing else running: H2D and D2H on two streams at once from the same "
                                        "page-locked buffers, all ranks together, max over ranks; an end-to-end step cannot be faster than "
                                        "max(this, the device-resident step)")
        return out

    def stage_block(res):
        prof = res["prof"]
        stage_ms = {k: round(v[0] / args.steps, 3) for k, v in prof.items()}
        # memory-bound stages: algorithmic bytes per step on this rank
        blend_bytes = plan.n_rows * W * 4 * (1 + (1.78 if method == "feather" else 1.0))
        norm_bytes = n_tiles_mine * (hr_tile * hr_tile * 4 * 2 + 32 * 32 * 4 * 2)
        mem = {
            "normalize": {"GBps": norm_bytes * args.steps / (prof["normalize"][0] / 1e3) / 1e9 if prof["normalize"][0] else None},
            "blend": {"GBps": blend_bytes * args.steps / (prof["blend"][0] / 1e3) / 1e9 if prof["blend"][0] else None},
            "hbm_peak_GBps": hbm_gbs,
        }
        for k in ("normalize", "blend"):
            if mem[k]["GBps"]:
                mem[k]["frac_of_hbm_peak"] = mem[k]["GBps"] / hbm_gbs
        return stage_ms, mem

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, affinity_before)  # the CPU leg runs on every core this process may use
        cb = cpu_reference_rate(model_fp, steps=2, warmup=1)
        cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}

    def precision_text(prec):
        return ("fp32 tolerance (<= 1e-4 m vs the fp32 CPU path): split fp16 (hi, lo) operands, 3 tcgen05 kind::f16 MMAs per product, "
                "fp32 accumulate in TMEM, partial accumulators summed in fp32 registers") if prec == "fp32" else \
               f"{prec} operands, fp32 accumulate (tcgen05 kind::f16): <= 1e-2 m mode, stated separately"

    stage_ms, mem_stages = stage_block(main_res)
    line = {
        "metric": METRIC,
        "value": mpx / (main_res["ms_per_step"] / 1e3),
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": main_res["ms_per_step"],
        "higher_is_better": True,
        "scaling": scaling,
        "vs_baseline": None,
        "dtype": args.precision,
        "data": "synthetic",
        "config": {
            "workload": f"{wl_desc}; {n_tiles_total} {method} windows" + (f" (overlap_lr={OVERLAP_LR}), halo rows exchanged between neighbour ranks" if method == "feather" else ""),
            "name": args.config,
            "model": model_desc,
            "tiles_per_step": n_tiles_total,
            "gflop_per_tile": 2 * main_res["macs_per_tile"] / 1e9,
            "l2_policy": f"inputs larger than L2 ({h_dem.nbytes / 2**20:.0f} MiB DEM rows per GPU vs 126 MB L2), no explicit flush",
            "precision": precision_text(args.precision),
            "parallelism": f"row-bands x{world}",
        },
        "roofline": roofline_of(args, main_res, peaks, n_tiles_mine),
        "stage_ms_per_step_rank0": stage_ms,
        "memory_bound_stages": mem_stages,
        "cpu_baseline": cpu_baseline,
        "e2e": e2e_of(main_res),
        "gpu_launches": main_res["launches"],
        "clocks": main_res["clocks"],
    }
    if other:
        line["modes"] = {}
        for prec, res in other.items():
            st, mem = stage_block(res)
            line["modes"][prec] = {
                "value": mpx / (res["ms_per_step"] / 1e3), "unit": UNIT, "ms_per_step": res["ms_per_step"], "dtype": prec,
                "precision": precision_text(prec), "roofline": roofline_of(args, res, peaks, n_tiles_mine),
                "stage_ms_per_step_rank0": st, "memory_bound_stages": mem, "e2e": e2e_of(res), "gpu_launches": res["launches"],
                "clocks": res["clocks"],
            }
    emit(line, stdout_fd)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
