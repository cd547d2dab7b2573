"""
GPU: eyes larger than the shared-memory ray table rendered as scans (more than four headings per position), i.e.
band by band -- plain large eyes and cone-sampled ones (S samples per ommatidium) -- against the C oracle.
"""
import numpy as np
import pytest
import torch

import c_oracle
import isy_oracle as o
from conftest import circ_diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import invertsy_b200
    assert torch.cuda.is_available()
    return invertsy_b200


def _np(r, **kw):
    out = r.render(**kw)
    torch.cuda.synchronize()
    r.check()
    return {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}


def _global_yaw(ry, h):
    gy = ry + h
    return np.where(gy > np.pi, gy - 2 * np.pi, np.where(gy <= -np.pi, gy + 2 * np.pi, gy))


def _positions(n, seed):
    rng = np.random.RandomState(seed)
    return np.c_[rng.uniform(0.5, 9.5, n), rng.uniform(0.5, 9.5, n), np.full(n, 0.01)]


@pytest.mark.parametrize("n_rays,headings", [
    (3000, np.deg2rad([-170., -95.5, -20., 13., 77.7, 101., 179.])),            # irregular: look-up windows
    (2500, np.deg2rad((np.arange(12) * 30. + 180.) % 360. - 180.)),             # even scan: arithmetic windows
    (1281, np.deg2rad((np.arange(36) * 10. + 180.) % 360. - 180.)),             # just over one band + a 257-ray tail
])
def test_large_eye_scan_is_rendered_in_bands_and_stays_exact(ib, n_rays, headings):
    """
    160 positions (more than there are SMs, so a position keeps all its headings: the banded plan) against
    (a) the same views rendered three positions at a time (one heading per group: the ray-grid walk), everywhere, and
    (b) the C oracle on a sample of views.
    """
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(n_rays)
    sun = (np.deg2rad(60.), np.pi)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.Sky(*sun))
    pos = _positions(160, n_rays)
    H = headings.size
    yw = np.tile(headings, (pos.shape[0], 1))
    names = ("hit", "depth", "rgb", "Y", "dop", "aop")
    out = _np(r, pos=pos, yaw=yw, outputs=names)
    for s in range(0, pos.shape[0], 40):                       # (a) a quarter of the positions, 3 at a time
        few = _np(r, pos=pos[s:s + 3], yaw=yw[s:s + 3], outputs=names)
        for k in names:
            a, b = out[k][s * H:(s + 3) * H], few[k]
            assert np.array_equal(a, b, equal_nan=True), k
    omm = o.eye_rotations(pitch, yaw)
    for p, h in ((0, 0), (7, H - 1), (80, H // 2), (159, 1)):  # (b)
        ref, depth, _ = c_oracle.world_hits(world.polygons, 2.0, pos[p], pitch, _global_yaw(yaw, headings[h]))
        b = p * H + h
        assert np.array_equal(out["hit"][b], ref)
        m = ref >= 0
        assert np.allclose(out["depth"][b][m], depth[m], rtol=1e-6) and np.all(np.isnan(out["depth"][b][~m]))
        assert np.array_equal(out["rgb"][b][m], world.colours[ref[m]].astype(np.float32))
    y, dop, aop = o.sky_render(sun[0], sun[1], o.view_rotations(np.rad2deg(headings[1]), omm))
    assert np.abs(out["Y"][1] - y).max() < 1e-5 and np.abs(out["dop"][1] - dop).max() < 1e-5
    assert circ_diff(out["aop"][1], aop).max() < 1e-5


def test_cone_sampled_eye_scan_in_bands(ib):
    """300 ommatidia x 16 samples = 4800 rays, 6 headings, 150 positions: per-ray hits exact, per-ommatidium outputs
    are the cone averages of the reference's per-sample values"""
    rng = np.random.RandomState(7)
    world = ib.Seville2009()
    N, S = 300, 16
    pitch, yaw = o.fibonacci_eye(N)
    rho = np.deg2rad(5.)
    rp = np.clip(pitch[:, None] + rho / 2 * rng.randn(N, S), -1.5, 1.5).ravel()
    ry = ((yaw[:, None] + rho / 2 * rng.randn(N, S) + np.pi) % (2 * np.pi) - np.pi).ravel()
    sun = (np.deg2rad(40.), 1.0)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(rp, ry, samples=S), sky=ib.Sky(*sun))
    headings = np.deg2rad(np.array([-150., -90., -30., 30., 90., 150.]))
    pos = _positions(150, 11)
    pos[0] = (5.1, 4.9, 0.01)
    names = ("hit", "rgb", "Y", "dop", "aop")
    out = _np(r, pos=pos, yaw=np.tile(headings, (150, 1)), outputs=names)
    assert out["hit"].shape == (900, N * S) and out["Y"].shape == (900, N)
    # the same views two positions at a time (one heading per group: ray-grid walk, generic cone average)
    for s in (0, 60, 148):
        few = _np(r, pos=pos[s:s + 2], yaw=np.tile(headings, (2, 1)), outputs=names)
        assert np.array_equal(out["hit"][s * 6:(s + 2) * 6], few["hit"])
        for k in ("rgb", "Y", "dop"):
            a, b = out[k][s * 6:(s + 2) * 6], few[k]
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.nanmax(np.abs(a - b)) < 2e-6, k
        assert circ_diff(out["aop"][s * 6:(s + 2) * 6], few["aop"]).max() < 2e-6
    samp = o.eye_rotations(rp, ry)
    for p, h in ((0, 2), (33, 0), (149, 5)):
        ref = c_oracle.world_hits(world.polygons, 2.0, pos[p], rp, _global_yaw(ry, headings[h]))[0]
        assert np.array_equal(out["hit"][p * 6 + h], ref)
    # one view in full: cone averages of the reference's per-sample values
    y, dop, aop = o.sky_render(sun[0], sun[1], o.view_rotations(np.rad2deg(headings[2]), samp))
    assert np.abs(out["Y"][2] - y.reshape(N, S).mean(1)).max() < 1e-5
    assert np.abs(out["dop"][2] - dop.reshape(N, S).mean(1)).max() < 1e-5
    circ = np.arctan2(np.sin(aop).reshape(N, S).mean(1), np.cos(aop).reshape(N, S).mean(1))   # circular mean of the samples
    assert circ_diff(out["aop"][2], circ).max() < 1e-5
    c = o.world_render(world.polygons, world.colours, 2.0, pos[0:1], o.view_rotations(np.rad2deg(headings[2]), samp))
    avg = c.astype(np.float64).reshape(N, S, 3).mean(1)
    got = out["rgb"][2]
    assert np.array_equal(np.isnan(got), np.isnan(avg)) and np.nanmax(np.abs(got - avg)) < 1e-5
