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


@pytest.mark.parametrize("n_rays,headings", [
    (3000, np.deg2rad([-170., -95.5, -20., 13., 77.7, 101., 179.])),            # irregular: look-up windows
    (2500, np.deg2rad((np.arange(12) * 30. + 180.) % 360. - 180.)),             # even scan: arithmetic windows
    (1281, np.deg2rad((np.arange(36) * 10. + 180.) % 360. - 180.)),             # just over one band + a 257-ray tail
])
def test_large_eye_scan_is_rendered_in_bands_and_stays_exact(ib, n_rays, headings):
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(n_rays)
    sun = (np.deg2rad(60.), np.pi)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.Sky(*sun))
    route = ib.Seville2009.load_routes(degrees=True)["path"][0]
    pos = route[[0, 400, 800], :3].copy()
    yw = np.tile(headings, (pos.shape[0], 1))
    out = _np(r, pos=pos, yaw=yw, outputs=("hit", "depth", "rgb", "Y", "dop", "aop"))
    H = headings.size
    omm = o.eye_rotations(pitch, yaw)
    for p in range(pos.shape[0]):
        for h in range(H):
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
    """300 ommatidia x 16 samples = 4800 rays, 6 headings: per-ray hits exact, per-ommatidium outputs = cone averages"""
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
    pos = np.array([[5.1, 4.9, 0.01], [2.2, 7.3, 0.01]])
    out = _np(r, pos=pos, yaw=np.tile(headings, (2, 1)), outputs=("hit", "rgb", "Y", "dop", "aop"))
    assert out["hit"].shape == (12, N * S) and out["Y"].shape == (12, N)
    samp = o.eye_rotations(rp, ry)
    for p in range(2):
        for h in range(6):
            b = p * 6 + h
            ref = c_oracle.world_hits(world.polygons, 2.0, pos[p], rp, _global_yaw(ry, headings[h]))[0]
            assert np.array_equal(out["hit"][b], ref)
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
