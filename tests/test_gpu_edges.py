"""GPU: edge cases of the world rule -- NaN triangles, other horizons, eye heights, positions outside the arena,
degenerate triangles, double-precision host path."""
import numpy as np
import pytest
import torch

import c_oracle
import isy_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import invertsy_b200
    assert torch.cuda.is_available()
    return invertsy_b200


def _hits(r, pos, yaw_rad):
    out = r.render(pos, yaw_rad, outputs=("hit", "depth"))
    torch.cuda.synchronize()
    r.check()
    return out["hit"].cpu().numpy(), out["depth"].cpu().numpy()


def _oracle(polygons, horizon, pos, yaw_rad, pitch, yaw):
    hits, depths = [], []
    for p in range(pos.shape[0]):
        for h in range(yaw_rad.shape[1]):
            gy = yaw + yaw_rad[p, h]
            gy = np.where(gy > np.pi, gy - 2 * np.pi, np.where(gy <= -np.pi, gy + 2 * np.pi, gy))
            hh, dd, _ = c_oracle.world_hits(polygons, horizon, pos[p], pitch, gy)
            hits.append(hh), depths.append(dd)
    return np.asarray(hits), np.asarray(depths)


@pytest.mark.parametrize("horizon", [0.5, 1.0, 5.0])
def test_other_horizons_and_eye_heights(ib, horizon):
    pol, col = ib.Seville2009.load_world()
    world = ib.WorldBase(pol, col, horizon=horizon)
    pitch, yaw = o.fibonacci_eye(1500)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw))
    rng = np.random.RandomState(int(horizon * 10))
    pos = np.c_[rng.uniform(-1, 11, 5), rng.uniform(-1, 11, 5), rng.uniform(0.0, 0.4, 5)]     # also outside the arena
    yaw_rad = rng.uniform(-np.pi, np.pi, (5, 3))
    hit, depth = _hits(r, pos, yaw_rad)
    ref, dref = _oracle(pol, horizon, pos, yaw_rad, pitch, yaw)
    assert (hit != ref).sum() == 0
    ok = ref >= 0
    assert np.abs(depth[ok] - dref[ok]).max() < 1e-6 and np.all(np.isnan(depth[~ok]))


def test_nan_and_degenerate_triangles(ib):
    """a NaN vertex removes the triangle (world.py:98, :102); a triangle that is degenerate in angle space
    paints every direction (world.py:526): both exactly as the oracle"""
    pol, col = ib.Seville2009.load_world()
    pol = pol.copy()
    pol[::7, 1, 2] = np.nan                                   # every 7th triangle gets a NaN vertex
    pos = np.array([[5.0, 5.0, 0.01]])
    # a triangle whose three vertices are collinear with the eye in (theta, phi): same azimuth and elevation ratios
    extra = np.array([[[5.5, 5.0, 0.01], [6.0, 5.0, 0.01], [6.5, 5.0, 0.01]]], dtype=np.float32)
    pol2 = np.concatenate([pol, extra])
    col2 = np.concatenate([col, np.array([[0.1, 0.2, 0.3]], dtype=np.float32)])
    pitch, yaw = o.fibonacci_eye(800)
    for polygons, colours in ((pol, col), (pol2, col2)):
        world = ib.WorldBase(polygons, colours, horizon=2.0)
        r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw))
        hit, _ = _hits(r, pos, np.array([[0.3, -2.5]]))
        ref, _ = _oracle(polygons, 2.0, pos, np.array([[0.3, -2.5]]), pitch, yaw)
        assert (hit != ref).sum() == 0
        assert not np.isin(hit[hit >= 0], np.arange(0, pol.shape[0], 7)).any()


def test_host_path_double_precision_and_background(ib):
    """isy_render_host with ISY_FLAG_OUT_F64 and the world painted over the sky background"""
    pitch, yaw = o.fibonacci_eye(1000)
    world = ib.Seville2009()
    ths, phs = 0.7, -1.2
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.Sky(ths, phs))
    route = ib.Seville2009.load_routes(degrees=True)["path"][0][[10, 500]]
    host = r.render_host(route[:, :3], np.deg2rad(route[:, 3:4]), outputs=("rgb", "Y", "dop", "aop"), f64=True,
                         init_from_sky=True)
    omm = o.eye_rotations(pitch, yaw)
    for k, pose in enumerate(route):
        ori = o.view_rotations(pose[3], omm)
        y, p, a = o.sky_render(ths, phs, ori)
        c = o.world_render(world.polygons, world.colours, 2.0, pose[None, :3], ori, init_c=y)
        assert host["rgb"].dtype == torch.float64
        assert np.abs(host["Y"][k].numpy() - y).max() < 1e-9 and np.abs(host["dop"][k].numpy() - p).max() < 1e-9
        assert np.abs(host["rgb"][k].numpy() - c).max() < 1e-7


def test_cone_samples_with_general_orientations(ib):
    """S = 2 cone samples and non yaw-only eye orientations (bins path) against the NumPy composition"""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.RandomState(8)
    N, S = 400, 2
    pitch, yaw = o.fibonacci_eye(N)
    rp = np.clip(pitch[:, None] + 0.02 * rng.randn(N, S), -1.5, 1.5).ravel()
    ry = ((yaw[:, None] + 0.02 * rng.randn(N, S) + np.pi) % (2 * np.pi) - np.pi).ravel()
    world = ib.SimpleWorld()
    r = ib.Renderer(world, ib.EyeLayout.from_angles(rp, ry, samples=S))
    pos = np.array([[5.0, 5.2, 0.01], [3.0, 6.0, 0.02]])
    view = R.from_euler('ZYX', [[0.4, 0.1, -0.2], [-2.0, -0.15, 0.05]])
    out = r.render(pos, quat=view.as_quat()[:, None, :], outputs=("rgb", "hit"), f64=True)
    torch.cuda.synchronize()
    r.check()
    omm = o.eye_rotations(rp, ry)
    for k in range(2):
        ori = view[k] * omm
        c = o.world_render(world.polygons, world.colours, 2.0, pos[k:k + 1], ori).astype(np.float64)
        avg = c.reshape(N, S, 3).mean(1)
        got = out["rgb"][k].cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(avg))
        assert np.nanmax(np.abs(got - avg)) < 1e-6 if np.isfinite(avg).any() else True
        yw, pt, _ = ori.as_euler('ZYX').T
        assert np.array_equal(out["hit"][k].cpu().numpy(), c_oracle.world_hits(world.polygons, 2.0, pos[k], pt, yw)[0])
