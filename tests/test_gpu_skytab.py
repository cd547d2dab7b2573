"""
GPU: rows mode of the small-eye kernel.  Template rows (what a ray that hits nothing gets; the sky of a uniform sky;
under a Perez sky the sky of (heading, ray), evaluated once per launch when all positions look along the same row of
headings under one sun -- grids, heat maps) are streamed into the views while the copies are rasterised, and only the
rays that hit something are patched afterwards.  It must give what the per-view passes give
(ISY_FLAG_PER_VIEW_SKY), what the oracle gives, and must step aside as soon as one position has a heading of its own.
"""
import numpy as np
import pytest
import torch

import isy_oracle as o
from conftest import circ_diff

pytestmark = pytest.mark.gpu

ALL = ("rgb", "hit", "depth", "Y", "dop", "aop")


@pytest.fixture(scope="module")
def ib():
    import invertsy_b200
    assert torch.cuda.is_available()
    return invertsy_b200


def _positions(n, seed):
    rng = np.random.RandomState(seed)
    return np.c_[rng.uniform(1, 9, n), rng.uniform(1, 9, n), np.full(n, 0.01)]


def _same(a, b, sky_tol):
    for k in a:
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
        if k in ("rgb", "hit", "depth"):
            assert np.array_equal(x, y, equal_nan=True), k
        elif k == "aop":
            assert circ_diff(x, y).max() <= sky_tol, k
        else:
            assert np.abs(x - y).max() <= sky_tol * max(1.0, np.abs(y).max()), k


@pytest.mark.parametrize("H,outputs,R", [(36, ALL, 1000), (72, ALL, 1000), (5, ("rgb", "Y", "aop"), 1000), (1, ALL, 1000),
                                         (7, ("hit", "Y", "dop", "aop"), 999), (36, ALL, 1001), (3, ALL, 36)])
def test_shared_heading_rows_match_the_per_view_sky(ib, H, outputs, R):
    # (R a multiple of 4: 16 bytes per lane; otherwise the ray-per-lane pass)
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(R)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.Sky(np.deg2rad(60.), np.pi))
    P = max(160, -(-4200 // H))      # rows mode starts at 4096 views
    pos = _positions(P, 31)
    heads = (np.arange(H) * 2 * np.pi / H + 0.3 + np.pi) % (2 * np.pi) - np.pi   # 72 headings: two groups of 36
    yw = np.tile(heads, (P, 1))
    tab = r.render(pos, yw, outputs=outputs)
    ref = r.render(pos, yw, outputs=outputs, per_view_sky=True)
    torch.cuda.synchronize()
    r.check()
    _same(tab, ref, 1e-6)
    # and against the oracle on a few views
    omm = o.eye_rotations(pitch, yaw)
    for b in (0, (P // 2) * H + H // 2, P * H - 1):
        y, dop, aop = o.sky_render(np.deg2rad(60.), np.pi, o.view_rotations(np.rad2deg(heads[b % H]), omm))
        assert np.abs(tab["Y"][b].cpu().numpy() - y).max() < 1e-5
        if "dop" in tab:
            assert np.abs(tab["dop"][b].cpu().numpy() - dop).max() < 1e-5
        assert circ_diff(tab["aop"][b].cpu().numpy(), aop).max() < 1e-5


def test_one_position_with_its_own_heading_turns_the_table_off(ib):
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(600)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.Sky(np.deg2rad(40.), 1.0))
    P, H = 400, 12
    pos = _positions(P, 32)
    yw = np.tile(np.deg2rad(np.arange(H) * 30. - 180.), (P, 1))
    yw[P - 1, H - 1] = np.nextafter(yw[P - 1, H - 1], 4.0)      # one ulp is enough
    yw[77, 3] += 0.5
    a = r.render(pos, yw, outputs=ALL)
    b = r.render(pos, yw, outputs=ALL, per_view_sky=True)
    torch.cuda.synchronize()
    r.check()
    for k in ALL:   # the same code path: bit-identical
        assert np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), equal_nan=True), k
    omm = o.eye_rotations(pitch, yaw)
    y, dop, aop = o.sky_render(np.deg2rad(40.), 1.0, o.view_rotations(np.rad2deg(yw[77, 3]), omm))
    assert np.abs(a["Y"][77 * H + 3].cpu().numpy() - y).max() < 1e-5
    assert circ_diff(a["aop"][77 * H + 3].cpu().numpy(), aop).max() < 1e-5


def test_host_buffer_path_uses_the_same_table(ib):
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.Sky(np.deg2rad(60.), np.pi))
    P, H = 128, 36
    pos = _positions(P, 33)
    yw = np.tile(np.deg2rad(np.arange(H) * 10. - 180.), (P, 1))
    h = r.render_host(pos, yw, outputs=ALL)
    d = r.render(pos, yw, outputs=ALL, per_view_sky=True)
    torch.cuda.synchronize()
    _same({k: h[k] for k in ALL}, {k: d[k] for k in ALL}, 1e-6)


@pytest.mark.parametrize("sky", ["uniform", "none"])
@pytest.mark.parametrize("H,R", [(36, 1000), (3, 250), (72, 999)])
def test_uniform_and_no_sky_rows_match_the_per_view_passes(ib, sky, H, R):
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(R)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.UniformSky(10.) if sky == "uniform" else None)
    P = max(150, -(-4200 // H))      # rows mode starts at 4096 views
    pos = _positions(P, 34)
    rng = np.random.RandomState(3)
    yw = rng.uniform(-np.pi, np.pi, (P, H))          # every position its own headings: no sky table involved
    outs = ALL if sky == "uniform" else ("rgb", "hit", "depth")
    a = r.render(pos, yw, outputs=outs)
    b = r.render(pos, yw, outputs=outs, per_view_sky=True)
    torch.cuda.synchronize()
    r.check()
    for k in outs:
        assert np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), equal_nan=True), k
    if sky == "uniform":
        assert torch.all(a["Y"] == 10.) and torch.all(a["dop"] == 0.) and torch.all(torch.isnan(a["aop"]))
    assert (a["hit"] >= 0).float().mean() > 0.01


def test_empty_and_sparse_worlds_in_rows_mode(ib):
    # positions that see nothing at all (no copies, stale tile) and the sparse world
    world = ib.SimpleWorld()
    pitch, yaw = o.fibonacci_eye(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.Sky(np.deg2rad(30.), 0.5))
    P, H = 600, 8
    rng = np.random.RandomState(4)
    pos = np.c_[rng.uniform(-30, 40, P), rng.uniform(-30, 40, P), np.full(P, 0.01)]   # most of them far outside
    yw = np.tile(np.deg2rad(np.arange(H) * 45. - 180.), (P, 1))
    a = r.render(pos, yw, outputs=ALL)
    b = r.render(pos, yw, outputs=ALL, per_view_sky=True)
    torch.cuda.synchronize()
    r.check()
    _same(a, b, 1e-6)


def test_output_buffers_off_the_16_byte_grid(ib):
    # rows mode copies 16 bytes per lane only when every buffer allows it: buffers that start 4 bytes off get the
    # 4-byte path and the same views
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), ib.Sky(np.deg2rad(60.), np.pi))
    P, H = 120, 36
    pos = _positions(P, 35)
    yw = np.tile(np.deg2rad(np.arange(H) * 10. - 180.), (P, 1))
    a = r.render(pos, yw, outputs=ALL)
    B, R = P * H, 1000
    shapes = {"rgb": (B, R, 3), "hit": (B, R), "depth": (B, R), "Y": (B, R), "dop": (B, R), "aop": (B, R)}
    off = {}
    for k, shp in shapes.items():
        n = int(np.prod(shp))
        base = torch.empty(n + 1, dtype=torch.int32 if k == "hit" else torch.float32, device="cuda")
        off[k] = base[1:].view(shp)
        assert off[k].data_ptr() % 16 == 4
    b = r.render(pos, yw, outputs=ALL, out=off)
    torch.cuda.synchronize()
    r.check()
    for k in ALL:
        assert np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), equal_nan=True), k
