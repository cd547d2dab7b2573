"""CPU: host-side logic of the drop-in (loaders, view generators, eta mask, C-ABI surface, loud failure)."""
import ast
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import isy_oracle as o
from conftest import ROOT, golden


def test_library_exports_every_declared_symbol():
    from invertsy_b200 import _lib
    header = open(os.path.join(ROOT, "include", "isy_b200.h")).read()
    declared = set(re.findall(r"\b(isy_[a-z_0-9]+)\s*\(", header))
    declared -= {"isy_status", "isy_sky_kind"}
    assert declared == set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().isy_abi_version() == _lib.ABI_VERSION == 2
    assert _lib.launch_count() >= 0


def test_flag_constants_match_the_header():
    from invertsy_b200 import _lib
    header = open(os.path.join(ROOT, "include", "isy_b200.h")).read()
    flags = {m.group(1): 1 << int(m.group(2)) for m in re.finditer(r"\b(ISY_FLAG_[A-Z_0-9]+)\s*=\s*1\s*<<\s*(\d+)", header)}
    assert flags == {"ISY_FLAG_INIT_FROM_SKY": _lib.FLAG_INIT_FROM_SKY, "ISY_FLAG_BRUTE_FORCE": _lib.FLAG_BRUTE_FORCE,
                     "ISY_FLAG_SUN_PER_VIEW": _lib.FLAG_SUN_PER_VIEW, "ISY_FLAG_OUT_F64": _lib.FLAG_OUT_F64,
                     "ISY_FLAG_PER_VIEW_SKY": _lib.FLAG_PER_VIEW_SKY, "ISY_FLAG_NO_TILES": _lib.FLAG_NO_TILES}


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "invertsy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "isy_oracle" not in src and "c_oracle" not in src and "ref_loader" not in src, f


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_a_gpu():
    from scipy.spatial.transform import Rotation as R
    import invertsy_b200 as ib
    from invertsy_b200._lib import IsyError
    ori = R.from_euler('Z', [[0.], [1.]])
    with pytest.raises(IsyError):
        ib.Seville2009()(np.zeros((1, 3)), ori)
    with pytest.raises(IsyError):
        ib.Sky(1., 0.)(ori)
    with pytest.raises(RuntimeError):
        ib.Renderer(ib.Seville2009(), ib.EyeLayout.from_angles(*ib.fibonacci_lattice(8)))


def test_world_call_error_behaviour_matches_reference():
    import invertsy_b200 as ib
    with pytest.raises(IndexError):          # world.py:80 -> np.shape(None)[0]
        ib.Seville2009()(np.zeros((1, 3)), None)


def test_sky_scalars_and_assertions():
    import invertsy_b200 as ib
    g = golden("sky")
    sky = ib.Sky(np.deg2rad(60), np.pi)
    assert np.array_equal([sky.A, sky.B, sky.C, sky.D, sky.E], g["coef"])
    assert sky.tau_L == float(g["tau_L"]) and sky.Y_z == float(g["Y_z"]) and sky.M_p == float(g["M_p"])
    with pytest.raises(AssertionError):
        _ = sky.Y
    with pytest.raises(AssertionError):
        sky.tau_L = 0.5
    sky.tau_L = 3.
    assert abs(sky.tau_L - 3.) < 1e-12
    assert np.allclose([sky.A, sky.B, sky.C, sky.D, sky.E], o.sky_coefficients(3.))
    s2 = ib.Sky(30., 90., degrees=True)
    assert abs(s2.theta_s - np.pi / 6) < 1e-15 and abs(s2.phi_s - np.pi / 2) < 1e-15
    c = sky.copy()
    assert abs(c.tau_L - sky.tau_L) < 1e-12 and c.theta_s == sky.theta_s


def test_loaders_and_lattices():
    import invertsy_b200 as ib
    pol, col = ib.Seville2009.load_world()
    assert pol.shape == (5000, 3, 3) and pol.dtype == np.float32 and col.shape == (5000, 3)
    assert np.all(col[:, [0, 2]] == 0) and col[:, 1].min() > 0.06 and col[:, 1].max() < 0.98
    sp, sc = ib.SimpleWorld.load_world()
    assert sp.shape == (3222, 3, 3) and np.allclose(sc, [0.8, 1.0, 0.8])
    routes = ib.Seville2009.load_routes(degrees=True)
    assert len(routes["path"]) == 133 and routes["path"][0].shape == (812, 4)
    assert np.allclose(routes["path"][0][0], [8.45, 6.30, 0.01, -139.653564], atol=1e-5)
    rad = ib.Seville2009.load_routes()["path"][0]
    assert np.allclose(np.rad2deg(rad[:, 3]), routes["path"][0][:, 3])
    p, y = ib.fibonacci_lattice(1000)
    p2, y2 = o.fibonacci_eye(1000)
    assert np.array_equal(p, p2) and np.array_equal(y, y2)
    q = ib.render.angles_to_quat(p, y)
    q2 = o.eye_rotations(p, y).as_quat()
    assert np.minimum(np.abs(q - q2).max(1), np.abs(q + q2).max(1)).max() < 1e-15


def test_view_generators_follow_the_reference_iteration_order():
    from invertsy_b200 import views
    rows, cols, oris = 7, 5, 16
    pos, yaw = views.grid_views(rows, cols, oris)
    assert pos.shape == (rows * cols, 3) and yaw.shape == (rows * cols, oris)
    for j, (row, col, ori) in enumerate(np.ndindex(rows, cols, oris)):     # simulation.py:2620, 2677
        p, h = divmod(j, oris)
        assert pos[p, 0] == np.float32(col * 10.) / np.float32(cols)
        assert pos[p, 1] == np.float32(row * 10.) / np.float32(rows)
        two_pi = np.float32(360.)
        assert yaw[p, h] == (np.float32(ori) * two_pi / np.float32(oris) + two_pi / 2) % two_pi - two_pi / 2
    assert yaw.min() >= -180 and yaw.max() < 180
    import invertsy_b200 as ib
    route = ib.Seville2009.load_routes(degrees=True)["path"][0][::40]
    L, K, H = route.shape[0], 5, 8
    pos, yaw = views.parallel_views(route, H, K, 0.02)
    span = route[-1, :2] - route[0, :2]
    unit = span / np.linalg.norm(span)
    disp = np.array([-unit[1], unit[0]])
    for j in range(L * K * H):                                              # simulation.py:2683-2701
        m, k, ori = (j // H) % L, j // (L * H), j % H
        r = 0.02 * float((k % 2 - (k + 1) % 2) * ((k + 1) // 2))
        p, h = divmod(j, H)
        assert np.allclose(pos[p], [route[m, 0] + r * disp[0], route[m, 1] + r * disp[1], route[m, 2]], atol=0, rtol=0)
        d_yaw = views.ori2yaw(ori, H)
        assert yaw[p, h] == (route[m, 3] + d_yaw + 180) % 360 - 180
    # sharding: contiguous, disjoint, complete
    for n, ws in [(40000, 8), (13, 4), (3, 8)]:
        blocks = [views.shard_positions(n, r, ws) for r in range(ws)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_eta_mask_semantics():
    from invertsy_b200.env.occlusion import add_noise
    assert add_noise(noise=0., shape=(5, 3)).shape == () and not add_noise(noise=0., shape=(5, 3))
    c = np.ones((5, 3))
    c[add_noise(noise=0., shape=c.shape)] = 0.
    assert c.sum() == 15
    for kw in [dict(noise=.4, shape=(10, 3)), dict(noise=.4, shape=(10,)), dict(noise=np.array([1, 0, 1, 0]), shape=(4,)),
               dict(noise=np.array([1, 0]), shape=(4,))]:
        a = add_noise(rng=np.random.RandomState(3), **kw)
        b = o.add_noise(rng=np.random.RandomState(3), **kw)
        assert np.array_equal(a, b) and a.dtype == b.dtype


def test_sky_table_construction_is_accurate_enough():
    """G(c) = exp(D acos c) from a 2 x 257 cubic-Hermite table in u = sqrt(1 - |c|) (csrc: sky_table_kernel,
    g_table_eval): the same construction in NumPy stays within 1e-9 of the exact value"""
    M = 256
    u = np.linspace(0, 1, M + 1)
    h = 1.0 / M
    rng = np.random.RandomState(0)
    c = np.clip(np.concatenate([rng.uniform(-1, 1, 400000), 1 - 10 ** rng.uniform(-12, 0, 50000),
                                -1 + 10 ** rng.uniform(-12, 0, 50000)]), -1, 1)
    uu = np.sqrt(1 - np.abs(c))
    idx = np.minimum((uu * M).astype(int), M - 1)
    t = uu * M - idx
    for D in (-3.5, -2.3359, -0.5, 1.0):
        g = 2 * np.arcsin(u / np.sqrt(2))
        dg = 2 / np.sqrt(2 - u * u)
        tabs = [(np.exp(D * g), np.exp(D * g) * D * dg * h), (np.exp(D * (np.pi - g)), -np.exp(D * (np.pi - g)) * D * dg * h)]
        out = np.empty_like(c)
        for half, (f, d) in enumerate(tabs):
            f0, f1, d0, d1 = f[idx], f[idx + 1], d[idx], d[idx + 1]
            c2 = 3 * (f1 - f0) - 2 * d0 - d1
            c3 = 2 * (f0 - f1) + d0 + d1
            val = f0 + t * (d0 + t * (c2 + t * c3))
            sel = (c < 0) == bool(half)
            out[sel] = val[sel]
        assert np.abs(out - np.exp(D * np.arccos(c))).max() < 1e-9


def test_ephemeris_matches_reference_goldens():
    """vectorised NOAA sun position == the reference's Sun.compute (goldens made from the reference)"""
    from datetime import datetime
    import invertsy_b200 as ib
    from invertsy_b200 import ephemeris
    g = golden("ephemeris")
    rows = g["rows"]
    dates = [datetime(*[int(v) for v in r[:6]]) for r in rows]
    alt, az = ephemeris.sun_position(rows[:, 6], rows[:, 7], dates, rows[:, 8])
    assert np.abs(alt - rows[:, 9]).max() < 1e-12 and np.abs(az - rows[:, 10]).max() < 1e-12
    sky = ib.Sky.from_observer()                       # Seville, 21/06/2021 10:00 (sky.py:597-609)
    if datetime.now().astimezone().utcoffset().total_seconds() == 0:   # the reference's tz rule reads the local clock
        assert np.allclose([sky.theta_s, sky.phi_s], g["default_sky"], atol=1e-12)
    from scipy.spatial.transform import Rotation as R
    s2 = ib.Sky.from_observer(ori=R.from_euler('Z', 0.5))
    assert abs(((s2.phi_s - (sky.phi_s - 0.5) + np.pi) % (2 * np.pi)) - np.pi) < 1e-12
