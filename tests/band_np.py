"""Test helper: NumPy executor with the semantics of fsr_band_run_dev / fsr_band_finalize_dev.

Lets the rank-to-rank halo exchange of `floodsr_b200.dist` run on CPU (gloo) without the CUDA engine.  The
arithmetic follows the reference mosaic (`floodsr/models/ResUNet_16x_DEM.py:315-363`, :391) restricted to
a band of window rows; per-tile predictions come from any object with the engine contract's `run_tile`.
"""

from __future__ import annotations

import numpy as np
import torch

from floodsr_b200.tiling import build_feather_ramp, window_grid


class NumpyBandExecutor:
    def __init__(self, engine, h, w, window_method="feather", overlap_hr=128, hr_tile=512, scale=16, max_depth=5.0):
        self.engine, self.h, self.w = engine, h, w
        self.T, self.scale, self.overlap, self.max_depth = hr_tile, scale, overlap_hr, max_depth
        self.method = window_method
        self.ys, self.xs = window_grid(h, w, hr_tile, window_method, overlap_hr)
        self.ramp = build_feather_ramp(hr_tile, overlap_hr) if window_method == "feather" else np.ones(hr_tile, np.float32)
        self._tiles = {}

    def _weights(self, yi, xi):
        wy, wx = self.ramp.copy(), self.ramp.copy()
        if self.method == "feather" and self.overlap > 0:
            if yi == 0:
                wy[: self.overlap] = 1.0
            if yi == len(self.ys) - 1:
                wy[-self.overlap:] = 1.0
            if xi == 0:
                wx[: self.overlap] = 1.0
            if xi == len(self.xs) - 1:
                wx[-self.overlap:] = 1.0
        return np.outer(wy, wx).astype(np.float32, copy=False)

    def _accumulate(self, plan, row0, n_rows, init):
        """Sum pred*w of this band's windows (window order) into rows [row0, row0+n_rows) on top of `init`."""
        T = self.T
        hp, wp = -(-self.h // T) * T, -(-self.w // T) * T
        acc = np.zeros((n_rows, wp), np.float32)
        if init is not None:
            acc[: init.shape[0], : self.w] = init
        for yi in range(plan.ty0, plan.ty1):
            for xi, x0 in enumerate(self.xs):
                y0 = self.ys[yi]
                a, b = max(y0, row0), min(y0 + T, row0 + n_rows, hp)
                if a >= b:
                    continue
                contrib = self._tiles[(yi, xi)] * self._weights(yi, xi)
                acc[a - row0 : b - row0, x0 : x0 + T] += contrib[a - y0 : b - y0]
        return acc

    def band_run(self, plan, depth_band, dem_band, band_row0):
        T, sc = self.T, self.scale
        dem = np.asarray(dem_band, np.float32)
        depth = np.asarray(depth_band, np.float32)
        hp, wp = -(-self.h // T) * T, -(-self.w // T) * T
        for yi in range(plan.ty0, plan.ty1):
            for xi, x0 in enumerate(self.xs):
                y0 = self.ys[yi] - band_row0
                dem_t = np.zeros((T, T), np.float32)
                src = dem[y0 : y0 + T, x0 : x0 + T]
                dem_t[: src.shape[0], : src.shape[1]] = src
                dep_t = np.zeros((T // sc, T // sc), np.float32)
                srcd = depth[y0 // sc : y0 // sc + T // sc, x0 // sc : x0 // sc + T // sc]
                dep_t[: srcd.shape[0], : srcd.shape[1]] = srcd
                self._tiles[(yi, xi)] = self.engine.run_tile(dep_t, dem_t, max_depth=self.max_depth)["prediction_m"]
        if plan.halo_out_rows == 0:
            return None
        acc = self._accumulate(plan, plan.row0 + plan.n_rows, plan.halo_out_rows, None)
        return torch.from_numpy(np.ascontiguousarray(acc[:, : self.w]))

    def band_finalize(self, plan, halo_in, out_rows=None):
        T = self.T
        init = None if halo_in is None else halo_in.numpy()
        acc = self._accumulate(plan, plan.row0, plan.n_rows, init)
        # weights are analytic: sum over ALL windows of the grid, own band or not
        wsum = np.zeros_like(acc)
        for yi, y0 in enumerate(self.ys):
            a, b = max(y0, plan.row0), min(y0 + T, plan.row0 + plan.n_rows)
            if a >= b:
                continue
            for xi, x0 in enumerate(self.xs):
                wsum[a - plan.row0 : b - plan.row0, x0 : x0 + T] += self._weights(yi, xi)[a - y0 : b - y0]
        if self.method == "hard":
            sr = acc
        else:
            sr = np.divide(acc, np.maximum(wsum, 1e-6), out=np.zeros_like(acc), where=wsum > 0)
        out = np.clip(sr[:, : self.w], 0.0, self.max_depth).astype(np.float32)
        return torch.from_numpy(out)

    def band_finalize_rows(self, plan, halo_in, out_rows, row_begin, row_end):
        """Semantics of fsr_band_finalize_rows_dev: only the band-relative rows [row_begin, row_end) are written; the rows
        that start from the previous band's sums (the first halo_in rows) must come in the call with row_begin == 0."""
        if out_rows is None:
            out_rows = torch.empty((plan.n_rows, self.w), dtype=torch.float32)
        if row_end <= row_begin:
            return out_rows
        assert halo_in is None or (row_begin == 0 and halo_in.shape[0] <= row_end)
        full = self.band_finalize(plan, halo_in if row_begin == 0 else None)
        if row_begin > 0 and plan.halo_in_rows > 0:
            assert row_begin >= plan.halo_in_rows, "rows below the shared ones only"
        out_rows[row_begin:row_end] = full[row_begin:row_end]
        return out_rows

    def make_recv(self, rows):
        return torch.empty((rows, self.w), dtype=torch.float32)
