"""CPU: the oracle (NumPy + C restatements) against the golden fixtures made from the unmodified reference."""
import numpy as np
import pytest

import c_oracle
import isy_oracle as o
from conftest import circ_diff, golden


def _queries(pose, omm):
    ori = o.view_rotations(pose[3], omm)
    yaw, pitch, _ = ori.as_euler('ZYX').T
    return ori, pitch, yaw


@pytest.mark.parametrize("name,world", [("world_seville", "seville_arrays"), ("world_sparse", "sparse_arrays")])
def test_numpy_oracle_matches_reference_rgb_and_hits(name, world, omm1000, request):
    polygons, colours = request.getfixturevalue(world)
    g = golden(name)
    _, _, omm = omm1000
    step = 4 if name == "world_seville" else 3
    for k in range(0, g["poses"].shape[0], step):
        pose = g["poses"][k]
        ori, pitch, yaw = _queries(pose, omm)
        rgb = o.world_render(polygons, colours, 2.0, pose[None, :3], ori)
        assert rgb.dtype == np.float32
        assert np.array_equal(rgb, g["rgb"][k], equal_nan=True)
        hit, depth, nvis = o.world_hits(polygons, 2.0, pose[:3], pitch, yaw)
        assert nvis == g["nvis"][k]
        assert np.array_equal(hit, g["hit"][k])


@pytest.mark.parametrize("name,world", [("world_seville", "seville_arrays"), ("world_sparse", "sparse_arrays")])
def test_c_oracle_matches_reference_hits(name, world, omm1000, request):
    polygons, colours = request.getfixturevalue(world)
    g = golden(name)
    _, _, omm = omm1000
    mism = 0
    for k in range(g["poses"].shape[0]):
        pose = g["poses"][k]
        _, pitch, yaw = _queries(pose, omm)
        if k < g["query"].shape[0]:   # SciPy's Euler angles are what they were when the goldens were made
            assert np.allclose(np.c_[pitch, yaw], g["query"][k], atol=1e-14, rtol=0)
        hit, depth, nvis = c_oracle.world_hits(polygons, 2.0, pose[:3], pitch, yaw)
        assert nvis == g["nvis"][k]
        mism += int((hit != g["hit"][k]).sum())
        ok = hit >= 0
        assert np.all(np.isnan(depth[~ok])) and np.all(depth[ok] <= 2.0)
        # rgb from hits: colours where hit, ground below the horizon, NaN elsewhere
        rgb = np.full((1000, 3), np.nan, dtype=np.float32)
        rgb[pitch >= 0] = np.array(o.GROUND_COLOUR) / 255.
        rgb[ok] = colours[hit[ok]]
        if not (hit != g["hit"][k]).any():
            assert np.array_equal(rgb, g["rgb"][k], equal_nan=True)
    assert mism == 0


def test_c_oracle_dense_band(seville_arrays):
    polygons, _ = seville_arrays
    g = golden("world_band")
    omm = o.eye_rotations(g["pitch"], g["yaw"])
    for k, pose in enumerate(g["poses"]):
        _, pitch, yaw = _queries(pose, omm)
        hit, _, _ = c_oracle.world_hits(polygons, 2.0, pose[:3], pitch, yaw)
        assert (hit != g["hit"][k]).sum() == 0


def test_oracle_init_c_and_uniform_sky(seville_arrays, omm1000):
    polygons, colours = seville_arrays
    g = golden("world_initc")
    _, _, omm = omm1000
    for k, pose in enumerate(g["poses"]):
        ori = o.view_rotations(pose[3], omm)
        rgb = o.world_render(polygons, colours, 2.0, pose[None, :3], ori, init_c=g["init_c"][k])
        assert rgb.dtype == np.float64 and np.array_equal(rgb, g["rgb"][k], equal_nan=True)
    y, p, a = o.uniform_sky_render(10., o.view_rotations(g["poses"][0][3], omm))
    assert np.array_equal(y, g["uniform_Y"]) and np.array_equal(p, g["uniform_P"])
    assert np.all(np.isnan(a)) and np.all(np.isnan(g["uniform_A"]))


def test_sky_oracles_match_reference(omm1000):
    from scipy.spatial.transform import Rotation as R
    g = golden("sky")
    _, _, omm = omm1000
    coef = o.sky_coefficients(2.)
    tau = o.sky_tau_from_coefficients(*coef)
    assert np.allclose(coef, g["coef"], rtol=0, atol=0) and tau == float(g["tau_L"])
    for k, (ths, phs, heading) in enumerate(g["suns"]):
        ori = o.view_rotations(heading, omm)
        y, p, a = o.sky_render(ths, phs, ori)
        assert np.abs(y - g["Y"][k]).max() < 1e-12 and np.abs(p - g["P"][k]).max() < 1e-12
        assert circ_diff(a, g["A"][k]).max() < 1e-12
        yi, _, _ = o.sky_render(ths, phs, ori, irgbu=g["irgbu"])
        assert np.abs(yi - g["Y_irgbu"][k]).max() < 1e-10
        sxy = R.from_euler('ZY', [phs, np.pi / 2 - ths]).apply([1, 0, 0])[:2]
        yc, pc, ac = c_oracle.sky(ori.apply([1, 0, 0]), ths, phs, coef, tau, .6, 4., sxy)
        assert np.abs(yc - g["Y"][k]).max() < 1e-10 and np.abs(pc - g["P"][k]).max() < 1e-12
        assert circ_diff(ac, g["A"][k]).max() < 1e-12
    hp, hy = o.fibonacci_eye(5000, hemisphere=True)
    y, p, a = o.sky_render(np.deg2rad(60), np.pi, o.eye_rotations(hp, hy))
    assert np.abs(y - g["hemi_Y"]).max() < 1e-12 and np.abs(p - g["hemi_P"]).max() < 1e-12
    assert circ_diff(a, g["hemi_A"]).max() < 1e-12
    assert abs(o.sky_zenith_luminance(tau, np.deg2rad(60)) - float(g["Y_z"])) < 1e-14
    assert abs(o.sky_max_dop(tau) - float(g["M_p"])) < 1e-15


def test_float32_positions_are_a_documented_deviation(seville_arrays, omm1000):
    """With float32 `pos` the reference evaluates vertex angles in float32 (world.py:94, 109-110);
    the float64 rule differs from it only on a handful of edge-grazing rays."""
    polygons, _ = seville_arrays
    g = golden("world_seville_pos32")
    _, _, omm = omm1000
    total = mism = 0
    for k, pose in enumerate(g["poses"]):
        _, pitch, yaw = _queries(pose, omm)
        pos32 = pose[:3].astype(np.float32).astype(np.float64)
        hit, _, _ = c_oracle.world_hits(polygons, 2.0, pos32, pitch, yaw)
        mism += int((hit != g["hit"][k]).sum())
        total += hit.size
    assert mism <= 2e-3 * total
