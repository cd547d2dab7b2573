"""
GPU: one compact pass over every rasteriser / final-pass variant of the yaw-only kernels (small, large and
cone-sampled eyes; scans, single headings, irregular headings; one sun, a sun per view, uniform sky; fp32 and fp64;
brute force), each checked against the C oracle on sampled views and against the brute-force kernel everywhere.
"""
import numpy as np
import pytest
import torch

import c_oracle
import isy_oracle as o
from conftest import circ_diff

pytestmark = pytest.mark.gpu

P = 150     # more positions than SMs: a position keeps all its headings (scan plans)


@pytest.fixture(scope="module")
def ib():
    import invertsy_b200
    assert torch.cuda.is_available()
    return invertsy_b200


def _positions(n, seed):
    rng = np.random.RandomState(seed)
    return np.c_[rng.uniform(1, 9, n), rng.uniform(1, 9, n), np.full(n, 0.01)]


def _gy(ry, h):
    gy = ry + h
    return np.where(gy > np.pi, gy - 2 * np.pi, np.where(gy <= -np.pi, gy + 2 * np.pi, gy))


def _eyes(ib):
    rng = np.random.RandomState(5)
    p, y = o.fibonacci_eye(400)
    yield "small", p, y, 1
    p, y = o.fibonacci_eye(1500)
    yield "large", p, y, 1
    N, S = 120, 16
    p, y = o.fibonacci_eye(N)
    rp = np.clip(p[:, None] + 0.04 * rng.randn(N, S), -1.5, 1.5).ravel()
    ry = ((y[:, None] + 0.04 * rng.randn(N, S) + np.pi) % (2 * np.pi) - np.pi).ravel()
    yield "cone", rp, ry, S


def _headings(kind):
    if kind == "scan":
        return np.deg2rad(np.arange(6) * 60. - 180.)
    if kind == "single":
        return np.deg2rad(np.array([37.]))
    return np.sort(np.random.RandomState(9).uniform(-3.1, 3.1, 7))


@pytest.mark.parametrize("kind", ["scan", "single", "irregular"])
def test_every_eye_class_and_heading_plan_matches_the_oracle(ib, kind):
    world = ib.Seville2009()
    sun = (np.deg2rad(55.), 2.5)
    heads = _headings(kind)
    H = heads.size
    pos = _positions(P, 21)
    for name, rp, ry, S in _eyes(ib):
        r = ib.Renderer(world, ib.EyeLayout.from_angles(rp, ry, samples=S), ib.Sky(*sun))
        out = r.render(pos, np.tile(heads, (P, 1)), outputs=("hit", "depth", "Y"))
        torch.cuda.synchronize()
        r.check()
        hit = out["hit"].cpu().numpy()
        # the brute-force kernel on a few positions: same rule, no acceleration structure
        bf = r.render(pos[:4], np.tile(heads, (4, 1)), outputs=("hit", "Y"), brute_force=True)
        assert np.array_equal(hit[:4 * H], bf["hit"].cpu().numpy()), (name, kind)
        assert torch.equal(out["Y"][:4 * H], bf["Y"]), (name, kind)
        for p_, h_ in ((0, 0), (71, H - 1), (P - 1, H // 2)):
            ref, depth, _ = c_oracle.world_hits(world.polygons, 2.0, pos[p_], rp, _gy(ry, heads[h_]))
            assert np.array_equal(hit[p_ * H + h_], ref), (name, kind, p_, h_)
            m = ref >= 0
            assert np.allclose(out["depth"][p_ * H + h_].cpu().numpy()[m], depth[m], rtol=1e-6)


def test_sun_per_view_uniform_sky_and_f64_on_large_and_cone_eyes(ib):
    world = ib.Seville2009()
    rng = np.random.RandomState(2)
    heads = _headings("scan")
    H = heads.size
    pos = _positions(P, 22)
    yw = np.tile(heads, (P, 1))
    for name, rp, ry, S in _eyes(ib):
        N = rp.size // S
        eye = ib.EyeLayout.from_angles(rp, ry, samples=S)
        r = ib.Renderer(world, eye, ib.Sky(1.0, 0.0))
        suns = np.c_[rng.uniform(0.1, 1.4, P * H), rng.uniform(-3, 3, P * H)]
        out = r.render(pos, yw, sun=suns, outputs=("Y", "dop", "aop"))
        torch.cuda.synchronize()
        r.check()
        samp = o.eye_rotations(rp, ry)
        for b in (0, 433, P * H - 1):
            y, dop, aop = o.sky_render(suns[b, 0], suns[b, 1], o.view_rotations(np.rad2deg(heads[b % H]), samp))
            assert np.abs(out["Y"][b].cpu().numpy() - y.reshape(N, S).mean(1)).max() < 1e-5, name
            assert np.abs(out["dop"][b].cpu().numpy() - dop.reshape(N, S).mean(1)).max() < 1e-5, name
            circ = np.arctan2(np.sin(aop).reshape(N, S).mean(1), np.cos(aop).reshape(N, S).mean(1))
            assert circ_diff(out["aop"][b].cpu().numpy(), circ).max() < 1e-5, name
        # uniform sky painted behind the world: 19 * 10 + (13, 135, 201) clipped to 255 where nothing is hit (world.py:479)
        ru = ib.Renderer(world, eye, ib.UniformSky(10.))
        ou = ru.render(pos, yw, outputs=("rgb", "hit", "Y", "dop", "aop"), init_from_sky=True)
        torch.cuda.synchronize()
        assert torch.all(ou["Y"] == 10.) and torch.all(ou["dop"] == 0.) and torch.all(torch.isnan(ou["aop"]))
        if S == 1:
            rgb, hit = ou["rgb"].cpu().numpy(), ou["hit"].cpu().numpy()
            up = (hit < 0) & (np.tile(rp, (P * H, 1)) < 0)
            assert np.allclose(rgb[up], [203., 255., 255.])
        # fp64 outputs take the reference-order path: equal hits, sky to 1e-9
        o64 = r.render(pos[:5], yw[:5], outputs=("hit", "Y"), f64=True)
        o32 = r.render(pos[:5], yw[:5], outputs=("hit", "Y"))
        assert torch.equal(o64["hit"], o32["hit"])
        assert (o64["Y"].double() - o32["Y"].double()).abs().max() < 1e-5
