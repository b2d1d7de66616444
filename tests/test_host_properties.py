"""Property tests (hypothesis) of the host-side geometry: window grids, band plans, raster windows."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from floodsr_b200.dist import plan_bands
from floodsr_b200.raster_io import bounds_of, clip_to_bounds, window_from_bounds
from floodsr_b200.tiling import build_tile_starts, window_grid


@settings(max_examples=200, deadline=None)
@given(n_tiles=st.integers(1, 80), tile=st.sampled_from([256, 512]), overlap=st.sampled_from([0, 64, 128, 192]))
def test_tile_starts_cover_the_axis(n_tiles, tile, overlap):
    """tiling.py:7-16 on padded extents (the worker pads to whole tiles first, ResUNet_16x_DEM.py:215-235): starts are
    increasing, begin at 0, the last window ends exactly at the extent and no gap is left."""
    total = n_tiles * tile
    stride = tile - overlap
    starts = build_tile_starts(total, tile, stride)
    assert starts[0] == 0 and all(b > a for a, b in zip(starts, starts[1:]))
    assert starts[-1] == total - tile
    assert all(b - a <= stride for a, b in zip(starts, starts[1:]))  # consecutive windows touch or overlap


@settings(max_examples=150, deadline=None)
@given(h=st.integers(16, 9000), w=st.integers(16, 3000), world=st.integers(1, 9), method=st.sampled_from(["feather", "hard"]),
       overlap=st.sampled_from([128, 128, 64, 192, 320, 400]))
def test_band_plans_partition_rows_and_chain_halos(h, w, world, method, overlap):
    h, w = h // 16 * 16, w // 16 * 16
    plans, ys, xs = plan_bands(h, w, 512, method, overlap, world)
    assert len(plans) == world
    live = [p for p in plans if not p.empty]
    assert live and live[0].ty0 == 0 and live[-1].ty1 == len(ys)
    # window rows: contiguous, disjoint; output rows: a partition of [0, h)
    assert all(a.ty1 == b.ty0 for a, b in zip(live, live[1:]))
    assert live[0].row0 == 0 and all(a.row0 + a.n_rows == b.row0 for a, b in zip(live, live[1:]))
    assert live[-1].row0 + live[-1].n_rows == h
    # the halo one band sends is the halo the next one expects, and it never exceeds one tile
    assert live[0].halo_in_rows == 0 and live[-1].halo_out_rows == 0
    assert all(a.halo_out_rows == b.halo_in_rows for a, b in zip(live, live[1:]))
    assert all(0 <= p.halo_out_rows <= 512 for p in live)
    # a band owns at least the rows it receives partial sums for, whatever the overlap (three or more window rows may
    # cover one coordinate): nothing of an incoming halo is left without an owner
    assert all(p.n_rows >= p.halo_in_rows for p in live)
    assert all(not p.empty for p in plans[: len(live)]) and all(p.empty for p in plans[len(live):])
    # every band's input rows contain the rows its windows read (clipped to the raster)
    for p in live:
        assert p.in_row0 == min(ys[p.ty0], h) and p.in_row0 + p.in_rows == min(ys[p.ty1 - 1] + 512, h)
    if method == "hard":
        assert all(p.halo_out_rows == 0 for p in live)
    assert ys == window_grid(h, w, 512, method, overlap)[0]


@settings(max_examples=200, deadline=None)
@given(c0=st.integers(0, 900), r0=st.integers(0, 900), wpx=st.integers(1, 100), hpx=st.integers(1, 100),
       res=st.sampled_from([0.5, 1.0, 2.0, 30.0]))
def test_window_from_bounds_inverts_pixel_aligned_bounds(c0, r0, wpx, hpx, res):
    t = (res, 0.0, 1000.0, 0.0, -res, 9000.0)
    bounds = (1000.0 + c0 * res, 9000.0 - (r0 + hpx) * res, 1000.0 + (c0 + wpx) * res, 9000.0 - r0 * res)
    assert window_from_bounds(bounds, t) == (r0, c0, hpx, wpx)
    ras = {"array": np.zeros((1000, 1000), np.float32), "transform": t, "height": 1000, "width": 1000}
    sub, ts = clip_to_bounds(ras, bounds)
    assert sub.shape == (min(hpx, 1000 - r0), min(wpx, 1000 - c0))
    if r0 + hpx <= 1000 and c0 + wpx <= 1000:
        assert np.allclose(bounds_of(ts, *sub.shape), bounds, rtol=0.0, atol=1e-9)
