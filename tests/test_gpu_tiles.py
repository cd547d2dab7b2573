"""
The tile walk of evenly spaced scans (csrc/isy_raster.cuh: copies filed into tiles of 64 pitch-sorted rays x one azimuth
column) against the column walk it replaces (ISY_FLAG_NO_TILES) and against the C oracle, over the shapes that pick
different paths: one group of headings / two groups (tile lists beside the copies), a large eye in bands, few
headings, rows mode and the per-view passes, positions that see nothing, and headings at the seam.
"""
import numpy as np
import pytest
import torch

from oracle import isy_oracle as o
from oracle import c_oracle

pytestmark = pytest.mark.gpu
ALL = ("rgb", "hit", "depth", "Y", "dop", "aop")


@pytest.fixture(scope="module")
def ib():
    import invertsy_b200
    assert torch.cuda.is_available()
    return invertsy_b200


def _positions(n, seed, lo=1., hi=9.):
    rng = np.random.RandomState(seed)
    return np.c_[rng.uniform(lo, hi, n), rng.uniform(lo, hi, n), np.full(n, 0.01)]


def _gy(ry, h):
    gy = ry + h
    return np.where(gy > np.pi, gy - 2 * np.pi, np.where(gy <= -np.pi, gy + 2 * np.pi, gy))


@pytest.mark.parametrize("R,H,P,first", [
    (1000, 36, 160, -180.),     # the benchmark shape: rows mode, lists over the boxes
    (1000, 72, 80, -180.),      # two groups of 36: lists beside the copies
    (400, 6, 200, -180.),       # few headings, a 16-ray tail chunk
    (2500, 12, 150, 7.5),       # a large eye: three bands
    (999, 5, 160, 0.3),         # odd sizes, the per-ray row copies
    (1280, 28, 150, 179.999),   # the largest small eye, a heading next to the seam
])
def test_tile_walk_equals_the_column_walk_and_the_oracle(ib, R, H, P, first):
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(R)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.Sky(np.deg2rad(55.), 2.5))
    pos = _positions(P, R + H)
    heads = (np.deg2rad(first + np.arange(H) * 360. / H) + np.pi) % (2 * np.pi) - np.pi
    yw = np.tile(heads, (P, 1))
    a = r.render(pos, yw, outputs=ALL)
    b = r.render(pos, yw, outputs=ALL, no_tiles=True)
    c = r.render(pos, yw, outputs=ALL, per_view_sky=True)        # the per-view final passes
    torch.cuda.synchronize()
    r.check()
    for k in ("hit", "depth", "rgb"):
        x = a[k].cpu().numpy()
        assert np.array_equal(x, b[k].cpu().numpy(), equal_nan=True), k
        assert np.array_equal(x, c[k].cpu().numpy(), equal_nan=True), k
    hit = a["hit"].cpu().numpy()
    assert (hit >= 0).mean() > 0.02
    for p_, h_ in ((0, 0), (P // 2, H - 1), (P - 1, H // 3)):
        ref, depth, _ = c_oracle.world_hits(world.polygons, 2.0, pos[p_], pitch, _gy(yaw, heads[h_]))
        assert np.array_equal(hit[p_ * H + h_], ref), (p_, h_)


def test_tile_walk_on_positions_that_see_nothing_and_the_sparse_world(ib):
    world = ib.SimpleWorld()
    pitch, yaw = o.fibonacci_eye(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.Sky(np.deg2rad(30.), 0.5))
    P, H = 600, 8
    pos = _positions(P, 4, -30., 40.)        # most of them far outside
    yw = np.tile(np.deg2rad(np.arange(H) * 45. - 180.), (P, 1))
    a = r.render(pos, yw, outputs=("hit", "depth"))
    b = r.render(pos, yw, outputs=("hit", "depth"), no_tiles=True)
    torch.cuda.synchronize()
    r.check()
    for k in a:
        assert np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), equal_nan=True), k
    assert (a["hit"].cpu().numpy() >= 0).any()
