"""Row-band sharding across ranks on CPU (gloo, world_size 2 and 3): the halo exchange of floodsr_b200.dist.

Each rank runs its band with the NumPy band executor (tests/band_np.py), sends its halo partial sums to the next
rank and blends the rows it owns; the gathered raster must equal the oracle's single-process tile loop bit for
bit (same window order => same float32 accumulation order).
"""

from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parents[1]


class AnalyticEngine:
    """Cheap stand-in with the engine contract: bit-reproducible function of the tile inputs."""

    def run_tile(self, depth, dem, max_depth=5.0, **_):
        d = np.asarray(depth, np.float32)
        up = np.repeat(np.repeat(d, 16, axis=0), 16, axis=1)
        e = np.asarray(dem, np.float32)
        pred = np.clip(np.float32(0.5) * up + np.float32(0.001) * e - np.float32(0.0625), 0.0, max_depth).astype(np.float32)
        return {"prediction_m": pred, "dem_stats_used": {"p_clip": 1.0, "dem_min": 0.0, "dem_max": 1.0}}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, h, w, out_dir, overlap_lr=8):
    sys.path.insert(0, str(REPO))
    from floodsr_b200.dist import plan_bands, run_band_step
    from floodsr_b200.synth import synth_raster
    from tests.band_np import NumpyBandExecutor

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        depth, dem = synth_raster(h, w, seed=h + w)
        plans, ys, xs = plan_bands(h, w, 512, "feather", overlap_lr * 16, world)
        plan = plans[rank]
        ex = NumpyBandExecutor(AnalyticEngine(), h, w, overlap_hr=overlap_lr * 16)
        rows = None
        if not plan.empty:
            r0 = plan.in_row0
            rows = run_band_step(ex, plan, plans, depth[r0 // 16 : (r0 + plan.in_rows + 15) // 16], dem[r0 : r0 + plan.in_rows], r0, dist, None)
        np.save(Path(out_dir) / f"rows_{rank}.npy", np.zeros((0, w), np.float32) if rows is None else rows.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("h,w,world,overlap_lr", [(1552, 1040, 2, 8), (2048, 1024, 3, 8), (512, 1024, 2, 8),
                                                  (1536, 1024, 3, 12), (2048, 528, 3, 20)])
def test_band_exchange_over_gloo_is_bit_identical_to_single_process(tmp_path, h, w, world, overlap_lr):
    """overlap_lr 12 at H = 1536 gives window rows [0, 320, 640, 960, 1024] (forced trailing window) and overlap_lr 20 covers
    coordinates with three window rows: bands must then own at least the rows of their incoming halo (dist.chain_safe_bands)."""
    from floodsr_b200.dist import plan_bands
    from floodsr_b200.synth import synth_raster
    from oracle.stitch_np import run_tiled

    port = _free_port()
    mp.spawn(_worker, args=(world, port, h, w, str(tmp_path), overlap_lr), nprocs=world, join=True)
    plans, _, _ = plan_bands(h, w, 512, "feather", overlap_lr * 16, world)
    got = np.concatenate([np.load(tmp_path / f"rows_{r}.npy") for r in range(world)], axis=0)
    depth, dem = synth_raster(h, w, seed=h + w)
    want, n_tiles, _ = run_tiled(AnalyticEngine(), depth, dem, window_method="feather", overlap_lr=overlap_lr)
    assert sum(p.n_rows for p in plans) == h
    assert got.shape == want.shape and np.array_equal(got, want)
