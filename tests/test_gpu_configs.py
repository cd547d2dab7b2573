"""
GPU: the other BASELINE.json configurations as parity cases (scaled to what the oracle finishes in seconds).
  config 2: sky polarisation sweep over sun positions for a dorsal (upper hemisphere) eye
  config 4: parallel route familiarity lines (offset routes x headings)
  config 5: synthetic stress -- large vegetation world, many-ommatidia eye with 16 cone samples
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


def test_config2_sky_sweep_sun_per_view(ib):
    """5000-ommatidia dorsal eye, a sub-grid of the 360 x 90 sun sweep, one sun per view, no world"""
    pitch, yaw = o.fibonacci_eye(5000, hemisphere=True)
    empty = ib.WorldBase(np.zeros((0, 3, 3), dtype=np.float32), np.zeros((0, 3), dtype=np.float32))
    r = ib.Renderer(empty, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.Sky(1.0, 0.0))
    az = np.deg2rad(np.arange(0, 360, 45.))
    el = np.deg2rad(np.array([0.5, 20.5, 45.5, 70.5, 89.5]))
    suns = np.array([(np.pi / 2 - e, (a + np.pi) % (2 * np.pi) - np.pi) for a in az for e in el])
    K = suns.shape[0]
    out = _np(r, pos=np.zeros((K, 3)), yaw=np.zeros((K, 1)), sun=suns, outputs=("Y", "dop", "aop", "rgb"))
    out64 = _np(r, pos=np.zeros((K, 3)), yaw=np.zeros((K, 1)), sun=suns, outputs=("Y", "dop", "aop"), f64=True)
    omm = o.eye_rotations(pitch, yaw)
    for k in range(0, K, 3):
        y, p, a = o.sky_render(suns[k, 0], suns[k, 1], omm)
        assert np.abs(out64["Y"][k] - y).max() < 1e-9 and np.abs(out64["dop"][k] - p).max() < 1e-9
        assert circ_diff(out64["aop"][k], a).max() < 1e-9
        assert np.abs(out["Y"][k] - y).max() < 1e-5 and np.abs(out["dop"][k] - p).max() < 1e-5
        assert circ_diff(out["aop"][k], a).max() < 1e-5
    assert np.all(np.isnan(out["rgb"]))          # dorsal rays, empty world: nothing seen (world.py:80)
    # sky outputs only: the element-wise sky kernel (sky_views_kernel) instead of the render kernel; headings too
    yw = np.linspace(-3.1, 3.1, K).reshape(K, 1)
    sky_only = _np(r, pos=np.zeros((K, 3)), yaw=yw, sun=suns, outputs=("Y", "dop", "aop"))
    general = _np(r, pos=np.zeros((K, 3)), yaw=yw, sun=suns, outputs=("Y", "dop", "aop", "rgb"))
    for name in ("Y", "dop", "aop"):
        assert np.array_equal(sky_only[name], general[name], equal_nan=True), name
    for k in range(1, K, 7):
        y, p, a = o.sky_render(suns[k, 0], suns[k, 1], o.view_rotations(np.rad2deg(yw[k, 0]), omm))
        assert np.abs(sky_only["Y"][k] - y).max() < 1e-5 and np.abs(sky_only["dop"][k] - p).max() < 1e-5
        assert circ_diff(sky_only["aop"][k], a).max() < 1e-5
    # one sun for all views, and a uniform sky
    one = _np(r, pos=np.zeros((6, 3)), yaw=yw[:6], outputs=("Y", "aop"))
    y, p, a = o.sky_render(1.0, 0.0, o.view_rotations(np.rad2deg(yw[3, 0]), omm))
    assert np.abs(one["Y"][3] - y).max() < 1e-5 and circ_diff(one["aop"][3], a).max() < 1e-5
    ru = ib.Renderer(empty, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.UniformSky(7.))
    uni = _np(ru, pos=np.zeros((4, 3)), yaw=yw[:4], outputs=("Y", "dop", "aop"))
    assert np.all(uni["Y"] == 7.) and np.all(uni["dop"] == 0.) and np.all(np.isnan(uni["aop"]))


def test_config4_parallel_lines_slice(ib):
    from invertsy_b200 import views
    world = ib.Seville2009()
    route = ib.Seville2009.load_routes(degrees=True)["path"][0][::90]
    pos, yaw_deg = views.parallel_views(route, nb_orientations=72, nb_parallel=3, disposition_step=0.02)
    pitch, yaw = o.fibonacci_eye(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.UniformSky(10.))
    out = _np(r, pos=pos, yaw=np.deg2rad(yaw_deg), outputs=("hit", "Y", "dop", "aop"))
    assert np.all(out["Y"] == 10.) and np.all(out["dop"] == 0.) and np.all(np.isnan(out["aop"]))
    omm = o.eye_rotations(pitch, yaw)
    for b in range(0, pos.shape[0] * 72, 97):
        p, h = divmod(b, 72)
        ori = o.view_rotations(yaw_deg[p, h], omm)
        yw, pt, _ = ori.as_euler('ZYX').T
        ref = c_oracle.world_hits(world.polygons, 2.0, pos[p], pt, yw)[0]
        assert np.array_equal(out["hit"][b], ref)


def test_config5_synthetic_stress_slice(ib):
    """20 k grass blades at Seville density, 2000-ommatidia eye x 16 cone samples (32 000 rays per view)"""
    rng = np.random.RandomState(2021)
    T, side = 20000, 20.
    base = np.c_[rng.uniform(0, side, T), rng.uniform(0, side, T), np.zeros(T)]
    d = rng.uniform(-.08, .08, (T, 2))
    tri = np.stack([base + np.c_[d, np.zeros(T)], base - np.c_[d, np.zeros(T)],
                    base + np.c_[rng.uniform(-.05, .05, (T, 2)), rng.lognormal(np.log(.18), .6, T)]], axis=1)
    tri = tri.astype(np.float32)
    col = np.c_[np.zeros(T), rng.uniform(.07, .98, T), np.zeros(T)].astype(np.float32)
    world = ib.WorldBase(tri, col, horizon=2.)
    N, S = 2000, 16
    pitch, yaw = o.fibonacci_eye(N)
    rho = np.deg2rad(4.)
    dp = rho / 2 * rng.randn(N, S)
    dy = rho / 2 * rng.randn(N, S)
    rp = np.clip(pitch[:, None] + dp, -1.5, 1.5).ravel()
    ry = ((yaw[:, None] + dy + np.pi) % (2 * np.pi) - np.pi).ravel()
    r = ib.Renderer(world, ib.EyeLayout.from_angles(rp, ry, samples=S), sky=ib.Sky(np.deg2rad(60), np.pi))
    P = 3
    pos = np.c_[rng.uniform(2, side - 2, P), rng.uniform(2, side - 2, P), np.full(P, 0.01)]
    yaw_rad = rng.uniform(-np.pi, np.pi, (P, 2))
    out = _np(r, pos=pos, yaw=yaw_rad, outputs=("hit", "rgb", "Y"))
    assert out["hit"].shape == (P * 2, N * S) and out["rgb"].shape == (P * 2, N, 3)
    for p in range(P):
        for h in range(2):
            gy = ry + yaw_rad[p, h]
            gy = np.where(gy > np.pi, gy - 2 * np.pi, np.where(gy <= -np.pi, gy + 2 * np.pi, gy))
            ref = c_oracle.world_hits(tri, 2.0, pos[p], rp, gy)[0]
            assert (out["hit"][p * 2 + h] != ref).sum() == 0
    assert (out["hit"] >= 0).mean() > 0.05


def test_spectrum_influence_per_view_in_batches(ib):
    """`sky(ori, irgbu=...)` weights the luminance of each call (sky.py:342-344, :671-703); a batched render does the same
    per view on the device (isy_spectrum_influence_views), float and double outputs, one irgbu row or one per ommatidium"""
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(1000)
    ths, phs = np.deg2rad(55.), -1.0
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.Sky(ths, phs))
    omm = o.eye_rotations(pitch, yaw)
    rng = np.random.RandomState(8)
    pos = np.c_[rng.uniform(2, 8, 5), rng.uniform(2, 8, 5), np.full(5, 0.01)]
    yw = rng.uniform(-180, 180, (5, 3))
    for irgbu in (np.array([[0., 0., 1., 0., 0.]]), np.array([[.2, .1, .8, .3, 1.]]), rng.uniform(0, 1, (1000, 5))):
        a = _np(r, pos=pos, yaw=np.deg2rad(yw), outputs=("Y", "dop"), irgbu=irgbu)
        b = _np(r, pos=pos, yaw=np.deg2rad(yw), outputs=("Y",), irgbu=irgbu, f64=True)
        for v in (0, 7, 14):
            y, p, _ = o.sky_render(ths, phs, o.view_rotations(yw[v // 3, v % 3], omm), irgbu=irgbu)
            assert np.abs(b["Y"][v] - y).max() < 1e-9 * max(1., np.abs(y).max())
            assert np.abs(a["Y"][v] - y).max() < 2e-5 * max(1., np.abs(y).max())
            assert np.abs(a["dop"][v] - p).max() < 1e-5
