"""Row-band sharding of one raster across the GPUs of a box.

The reference is single-process (`floodsr/models/ResUNet_16x_DEM.py:307-356` is a serial Python loop), so
this is new: the window grid's tile rows are split into contiguous bands, one rank per band.  Every window
is independent (tile-local DEM stats), so the only data exchanged is, per band boundary, the partial sums
`sum(pred * w)` of the output rows that the upper band's last windows share with the lower band — one
point-to-point transfer of `halo_rows x W` float32 from rank r to rank r+1.  Feather weights are analytic
and recomputed locally.  The receiving band continues the sum in the reference's window order, so the
sharded raster is bit-identical to the single-GPU one.

The exchange itself goes through `torch.distributed` (NCCL on GPUs, gloo in the CPU tests); the band
arithmetic is delegated to an executor (the CUDA engine in production).
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from floodsr_b200.tiling import build_feather_ramp, split_tile_rows, window_grid


@dataclass(frozen=True)
class BandPlan:
    rank: int
    ty0: int
    ty1: int
    row0: int           # first output row owned by the band
    n_rows: int         # output rows owned
    halo_out_rows: int  # rows below the band that its windows also touch (sent to the next band)
    halo_in_rows: int   # rows at the top of the band that the previous band's windows touch (received)
    in_row0: int        # raster rows read by the band's windows: [in_row0, in_row0 + in_rows)
    in_rows: int

    @property
    def empty(self) -> bool:
        return self.ty1 <= self.ty0


def plan_bands(h: int, w: int, hr_tile: int, window_method: str, overlap_hr: int, world: int) -> tuple[list[BandPlan], list[int], list[int]]:
    """Partition the window grid of an `h x w` raster into `world` row bands (trailing bands may be empty)."""
    ys, xs = window_grid(h, w, hr_tile, window_method, overlap_hr)
    ny = len(ys)
    plans: list[BandPlan] = []
    prev_halo = 0
    for rank, (ty0, ty1) in enumerate(chain_safe_bands(split_tile_rows(ny, world), ys, hr_tile, h)):
        if ty1 <= ty0:
            plans.append(BandPlan(rank, ty0, ty1, h, 0, 0, 0, h, 0))
            continue
        row0 = 0 if ty0 == 0 else min(ys[ty0], h)
        row_end = h if ty1 >= ny else min(ys[ty1], h)
        halo_out = 0 if ty1 >= ny else max(min(ys[ty1 - 1] + hr_tile, h) - row_end, 0)
        in_row0 = min(ys[ty0], h)
        in_end = min(ys[ty1 - 1] + hr_tile, h)
        plans.append(BandPlan(rank, ty0, ty1, row0, max(row_end - row0, 0), halo_out, prev_halo, in_row0, max(in_end - in_row0, 0)))
        prev_halo = halo_out
    return plans, ys, xs


def chain_safe_bands(bands: list[tuple[int, int]], ys: list[int], hr_tile: int, h: int) -> list[tuple[int, int]]:
    """Move band boundaries so that every band OWNS at least the rows it receives partial sums for.

    A band's incoming halo ends where the previous band's last window ends.  When window rows overlap so much that a
    coordinate is covered by three or more of them (overlap >= tile / 2, or a forced trailing window close to its
    predecessor: `build_tile_starts`, floodsr/tiling.py:7-16), that point can lie beyond the next band's first window
    row; the rows past the band's own would have nowhere to go.  Such a band takes over the following window rows until
    its owned rows reach the end of the incoming halo; bands emptied that way move to the end of the list.
    """
    ny = len(ys)
    out: list[tuple[int, int]] = []
    ty0 = 0
    for _, want_end in bands:
        if ty0 >= ny:
            break
        ty1 = max(want_end, ty0 + 1) if want_end > ty0 else ty0
        if ty1 <= ty0:
            continue
        if ty0 > 0:
            halo_end = min(ys[ty0 - 1] + hr_tile, h)
            while ty1 < ny and min(ys[ty1], h) < halo_end:
                ty1 += 1
        out.append((ty0, ty1))
        ty0 = ty1
    if out and out[-1][1] < ny:
        out[-1] = (out[-1][0], ny)
    while len(out) < len(bands):
        out.append((ny, ny))
    return out


def exchange_halo(plan: BandPlan, plans: list[BandPlan], halo_out, make_recv, dist_mod, group=None):
    """Send this band's halo partial sums to the next band and receive the previous band's.

    `halo_out`: tensor [halo_out_rows, W] or None; `make_recv(rows)` allocates the receive tensor.
    Returns the received tensor or None.  Uses one batched isend/irecv pair (NCCL groups them).
    """
    ops = []
    recv = None
    nxt = plan.rank + 1
    if not plan.empty and plan.halo_out_rows > 0 and nxt < len(plans) and not plans[nxt].empty:
        ops.append(dist_mod.P2POp(dist_mod.isend, halo_out, nxt, group=group))
    if not plan.empty and plan.halo_in_rows > 0 and plan.rank > 0:
        recv = make_recv(plan.halo_in_rows)
        ops.append(dist_mod.P2POp(dist_mod.irecv, recv, plan.rank - 1, group=group))
    if ops:
        for req in dist_mod.batch_isend_irecv(ops):
            req.wait()
    return recv


def start_halo_exchange(plan: BandPlan, plans: list[BandPlan], halo_out, make_recv, dist_mod, group=None):
    """`exchange_halo` without the wait: returns (receive tensor or None, outstanding requests)."""
    ops = []
    recv = None
    nxt = plan.rank + 1
    if not plan.empty and plan.halo_out_rows > 0 and nxt < len(plans) and not plans[nxt].empty:
        ops.append(dist_mod.P2POp(dist_mod.isend, halo_out, nxt, group=group))
    if not plan.empty and plan.halo_in_rows > 0 and plan.rank > 0:
        recv = make_recv(plan.halo_in_rows)
        ops.append(dist_mod.P2POp(dist_mod.irecv, recv, plan.rank - 1, group=group))
    return recv, (dist_mod.batch_isend_irecv(ops) if ops else [])


class CudaBandExecutor:
    """Band arithmetic on the CUDA engine: inputs/outputs are torch CUDA tensors owned by the caller."""

    def __init__(self, engine, h: int, w: int, window_method: str, overlap_hr: int, max_depth: float = 5.0, dem_pct_clip: float = 95.0):
        import torch

        from floodsr_b200 import _lib

        self.torch = torch
        self.lib = _lib.load_library()
        self._lib_mod = _lib
        self.engine = engine
        self.h, self.w = int(h), int(w)
        c = engine.contract
        self.hr_tile, self.scale = c.dem_hr_hwc[0], c.scale
        self.window_method, self.overlap_hr = window_method, int(overlap_hr)
        self.ys, self.xs = window_grid(h, w, self.hr_tile, window_method, overlap_hr)
        self.ramp = build_feather_ramp(self.hr_tile, overlap_hr) if window_method == "feather" else None
        self.params = engine._tile_params(max_depth, dem_pct_clip, None, None, None, True)
        self.device = torch.device("cuda", engine.device)
        self._windows_set = False

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def set_windows(self):
        if self._windows_set:
            return
        ys = np.asarray(self.ys, np.int32)
        xs = np.asarray(self.xs, np.int32)
        self._lib_mod.check(
            self.lib.fsr_set_windows(
                self.engine._handle, self.h, self.w,
                self._lib_mod.WINDOW_FEATHER if self.window_method == "feather" else self._lib_mod.WINDOW_HARD,
                self.overlap_hr, self._lib_mod.iptr(ys), len(ys), self._lib_mod.iptr(xs), len(xs),
                self._lib_mod.fptr(self.ramp) if self.ramp is not None else None, self._stream(),
            )
        )
        self._windows_set = True

    def band_run(self, plan: BandPlan, depth_band, dem_band, band_row0: int):
        """depth_band / dem_band: CUDA float32 tensors holding raster rows from HR row `band_row0` on."""
        torch = self.torch
        self.set_windows()
        assert dem_band.is_cuda and dem_band.dtype == torch.float32 and dem_band.is_contiguous()
        assert depth_band.is_cuda and depth_band.dtype == torch.float32 and depth_band.is_contiguous()
        halo = None
        if plan.halo_out_rows > 0:
            halo = torch.empty((plan.halo_out_rows, self.w), dtype=torch.float32, device=self.device)
        self._lib_mod.check(
            self.lib.fsr_band_run_dev(
                self.engine._handle, C.c_void_p(depth_band.data_ptr()), C.c_void_p(dem_band.data_ptr()), int(band_row0),
                int(dem_band.shape[0]), plan.ty0, plan.ty1, C.byref(self.params),
                C.c_void_p(halo.data_ptr()) if halo is not None else None, None, self._stream(),
            )
        )
        return halo

    def band_finalize(self, plan: BandPlan, halo_in, out_rows=None):
        torch = self.torch
        if out_rows is None:
            out_rows = torch.empty((plan.n_rows, self.w), dtype=torch.float32, device=self.device)
        assert out_rows.shape == (plan.n_rows, self.w) and out_rows.is_contiguous()
        self._lib_mod.check(
            self.lib.fsr_band_finalize_dev(
                self.engine._handle, C.c_void_p(halo_in.data_ptr()) if halo_in is not None else None,
                int(halo_in.shape[0]) if halo_in is not None else 0, C.c_void_p(out_rows.data_ptr()), self._stream(),
            )
        )
        return out_rows

    def band_finalize_rows(self, plan: BandPlan, halo_in, out_rows, row_begin: int, row_end: int):
        """`band_finalize` for the band-relative rows [row_begin, row_end) only (out_rows always spans the whole band)."""
        torch = self.torch
        if out_rows is None:
            out_rows = torch.empty((plan.n_rows, self.w), dtype=torch.float32, device=self.device)
        assert out_rows.shape == (plan.n_rows, self.w) and out_rows.is_contiguous()
        if row_end > row_begin:
            self._lib_mod.check(
                self.lib.fsr_band_finalize_rows_dev(
                    self.engine._handle, C.c_void_p(halo_in.data_ptr()) if halo_in is not None else None,
                    int(halo_in.shape[0]) if halo_in is not None else 0, C.c_void_p(out_rows.data_ptr()), int(row_begin),
                    int(row_end), self._stream(),
                )
            )
        return out_rows

    def make_recv(self, rows: int):
        return self.torch.empty((rows, self.w), dtype=self.torch.float32, device=self.device)

    # -- host-buffer path: the band's H2D copies, kernels and D2H copies overlap inside the engine --------------------
    def band_host_begin(self, plan: BandPlan, depth_host: np.ndarray, dem_host: np.ndarray, band_row0: int, out_rows: np.ndarray):
        """Queue the whole band from (page-locked) host arrays; returns this band's halo partial sums (CUDA tensor or None)."""
        torch = self.torch
        self.set_windows()
        assert dem_host.dtype == np.float32 and dem_host.flags["C_CONTIGUOUS"] and dem_host.shape[1] == self.w
        assert depth_host.dtype == np.float32 and depth_host.flags["C_CONTIGUOUS"]
        assert out_rows.dtype == np.float32 and out_rows.flags["C_CONTIGUOUS"] and out_rows.shape == (plan.n_rows, self.w)
        halo = None
        if plan.halo_out_rows > 0:
            halo = torch.empty((plan.halo_out_rows, self.w), dtype=torch.float32, device=self.device)
        self._lib_mod.check(
            self.lib.fsr_band_host_begin(
                self.engine._handle, self._lib_mod.fptr(depth_host), self._lib_mod.fptr(dem_host), int(band_row0),
                int(dem_host.shape[0]), plan.ty0, plan.ty1, C.byref(self.params), self._lib_mod.fptr(out_rows),
                C.c_void_p(halo.data_ptr()) if halo is not None else None,
            )
        )
        return halo

    def band_host_end(self, halo_in):
        """Blend the rows shared with the previous band (its partial sums in `halo_in`), wait for the copies."""
        flags = C.c_uint32(0)
        code = self.lib.fsr_band_host_end(
            self.engine._handle, C.c_void_p(halo_in.data_ptr()) if halo_in is not None else None,
            int(halo_in.shape[0]) if halo_in is not None else 0, C.byref(flags),
        )
        if code == self._lib_mod.FSR_E_ASSERT:
            self.engine._raise_flags(int(flags.value), True, None)
        self._lib_mod.check(code)

    def check_flags(self):
        flags = C.c_uint32(0)
        self._lib_mod.check(self.lib.fsr_fetch_flags(self.engine._handle, self._stream(), C.byref(flags)))
        self.engine._raise_flags(int(flags.value), True, None)


def run_band_step_host(executor, plan: BandPlan, plans: list[BandPlan], depth_host, dem_host, band_row0: int, out_rows,
                       dist_mod=None, group=None):
    """`run_band_step` from / to host arrays (page-locked for full overlap): copies are pipelined with the kernels."""
    if plan.empty:
        return None
    halo_out = executor.band_host_begin(plan, depth_host, dem_host, band_row0, out_rows)
    halo_in = None
    if dist_mod is not None and len(plans) > 1:
        halo_in = exchange_halo(plan, plans, halo_out, executor.make_recv, dist_mod, group)
        executor.torch.cuda.current_stream(executor.device).synchronize()
    executor.band_host_end(halo_in)
    return out_rows


def run_band_step(executor, plan: BandPlan, plans: list[BandPlan], depth_band, dem_band, band_row0: int, dist_mod=None, group=None, out_rows=None):
    """One rank's share of a sharded raster pass: run the band, exchange halos, blend the owned rows."""
    if plan.empty:
        return None
    halo_out = executor.band_run(plan, depth_band, dem_band, band_row0)
    halo_in = None
    if dist_mod is not None and len(plans) > 1:
        if hasattr(executor, "band_finalize_rows"):
            # the rows below the shared ones do not depend on the previous rank: blend them while the halo is in flight
            halo_in, reqs = start_halo_exchange(plan, plans, halo_out, executor.make_recv, dist_mod, group)
            split = plan.halo_in_rows if halo_in is not None else 0
            out_rows = executor.band_finalize_rows(plan, None, out_rows, split, plan.n_rows)
            for req in reqs:
                req.wait()
            if split > 0:
                executor.band_finalize_rows(plan, halo_in, out_rows, 0, split)
            return out_rows
        halo_in = exchange_halo(plan, plans, halo_out, executor.make_recv, dist_mod, group)
    return executor.band_finalize(plan, halo_in, out_rows)
