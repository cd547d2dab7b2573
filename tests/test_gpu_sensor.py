"""GPU: the CompoundEye-compatible sensor (N1) and the batched dataset writer (N2)."""
import os

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation as R

import isy_oracle as o

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import invertsy_b200
    assert torch.cuda.is_available()
    return invertsy_b200


def _oracle_responses(eye, world, sky_args, pos, ori, uniform=None):
    """reference world()/sky() on every sample direction (the oracle port), then the eye's documented transform stated
    here in NumPy: sample colours on the 0..1 scale (sky background = luminance2skycolour / 255), weighted cone
    average, clip(colour * c_sensitive[RGB] * omm_res, 0, 1)"""
    lay = eye._eye_layout()
    glob = ori * R.from_quat(lay.quat)
    c = o.world_render(world.polygons, world.colours, 2.0, pos.reshape(1, 3), glob).astype(np.float64)   # NaN = nothing seen
    bg = np.zeros_like(c)
    if sky_args is not None:
        y, _, _ = o.sky_render(sky_args[0], sky_args[1], glob)
        bg = o.luminance2skycolour(y) / 255.
    elif uniform is not None:
        bg = o.luminance2skycolour(np.full(c.shape[0], uniform)) / 255.
    c = np.where(np.isnan(c), bg, c).reshape(eye.nb_ommatidia, eye.nb_samples, 3)
    w = lay.weights.astype(np.float64).reshape(eye.nb_ommatidia, eye.nb_samples) if lay.weights is not None \
        else np.full((eye.nb_ommatidia, eye.nb_samples), 1. / eye.nb_samples)
    avg = (c * w[:, :, None]).sum(1)
    return np.clip(avg * eye.c_sensitive[1:4] * eye.omm_res, 0., 1.)


@pytest.mark.parametrize("nb_samples,with_sky", [(1, False), (4, True), (16, True)])
def test_compound_eye_matches_reference_composition(ib, nb_samples, with_sky):
    world = ib.Seville2009()
    sky_args = (np.deg2rad(50.), 2.0) if with_sky else None
    sky = ib.Sky(*sky_args) if with_sky else None
    eye = ib.CompoundEye(nb_input=600, omm_rho=np.deg2rad(6.), omm_res=1.5, c_sensitive=[0, 0.2, 1., 0.1, 0],
                         nb_samples=nb_samples)
    route = ib.Seville2009.load_routes(degrees=True)["path"][0]
    seen = []
    for k in (0, 350, 700):
        eye.xyz = route[k, :3]
        eye.ori = R.from_euler('ZYX', [route[k, 3], 5., -3.], degrees=True)    # not yaw-only: the bins path
        r = eye(sky=sky, scene=world, callback=seen.append)
        assert r.shape == (600, 3) and r.dtype == np.float32 and r.min() >= 0 and r.max() <= 1
        ref = _oracle_responses(eye, world, sky_args, route[k, :3], eye.ori)
        assert np.abs(r - ref).max() < 1e-5
        assert np.allclose(np.clip(r.mean(axis=1), 0, 1), np.clip(ref.mean(axis=1), 0, 1), atol=1e-5)   # agent.py:744
        # the same pose as a yaw-only orientation (what the reference's simulations set, simulation.py:2706): the
        # call takes the rasteriser for yaw-only views and must agree with the reference composition too
        eye.ori = R.from_euler('Z', route[k, 3], degrees=True)
        r = eye(sky=sky, scene=world)
        ref = _oracle_responses(eye, world, sky_args, route[k, :3], eye.ori)
        assert np.abs(r - ref).max() < 1e-5
        # ... and the one-channel PN output of the device is clip(mean over channels) of the three (agent.py:744)
        pn = eye.render_batch(world, sky, route[k:k + 1, :3], np.array([[route[k, 3]]]), channels=1)
        assert pn.shape == (1, 600) and np.abs(pn[0] - np.clip(ref.mean(axis=1), 0, 1)).max() < 1e-5
    assert len(seen) == 3 and seen[0] is eye
    assert abs(eye.yaw_deg - route[700, 3]) < 1e-9 and eye.x == route[700, 0]


def test_noise_blanks_ommatidia_like_the_reference_environment(ib):
    world = ib.Seville2009()
    clean = ib.CompoundEye(nb_input=500, omm_res=2.)
    noisy = ib.CompoundEye(nb_input=500, omm_res=2., noise=0.2, rng=np.random.RandomState(5))
    for e in (clean, noisy):
        e.xyz = [5., 5., 0.01]
    a, b = clean(scene=world), noisy(scene=world)
    rng = np.random.RandomState(5)
    eta = np.argsort(np.absolute(rng.randn(500, 3)))[:int(0.2 * 500)]          # env/_helpers.py:27-29
    want = a.copy()
    want[eta] = 0.
    assert np.array_equal(b, want)


def test_dataset_writer_matches_per_view_calls_and_schema(ib, tmp_path):
    from invertsy_b200 import datasets, views
    world = ib.Seville2009()
    sky = ib.UniformSky(10.)                                   # simulation.py:2602-2603
    eye = ib.CompoundEye(nb_input=300, omm_rho=np.deg2rad(4.), omm_res=10., c_sensitive=[0, 0., 1., 0., 0.], nb_samples=4)
    route = ib.Seville2009.load_routes(degrees=True)["path"][0][::100]
    name = datasets.dataset_name(4, 3, 2, 1, 1, world, nb_ommatidia=300)
    assert name == "dataset-scan4-rows3-cols2-ant1-route1-seville2009-omm300"
    stats = datasets.collect_familiarity_dataset(route, eye, world, sky, method="grid", nb_orientations=4, nb_rows=3,
                                                 nb_cols=2, filename=str(tmp_path / name))
    L = route.shape[0]
    assert stats["ommatidia_out"].shape == (L, 300, 3) and stats["xyz_out"].shape == (L, 4)
    assert stats["ommatidia"].shape == (3 * 2 * 4, 300, 3) and stats["xyz"].shape == (24, 4)
    with np.load(str(tmp_path / name) + ".npz") as f:
        assert sorted(f.keys()) == ["ommatidia", "ommatidia_out", "xyz", "xyz_out"]
        assert np.array_equal(f["ommatidia"], stats["ommatidia"])
    # the same views one at a time, as the reference's loop does (simulation.py:2660-2707)
    pos, yaw = views.grid_views(3, 2, 4, z=route[0, 2])
    for j in (0, 5, 23):
        p, h = divmod(j, 4)
        eye.xyz = pos[p]
        eye.ori = R.from_euler('Z', yaw[p, h], degrees=True)
        r = eye(sky=sky, scene=world)
        assert np.abs(r - stats["ommatidia"][j]).max() < 1e-6
        assert np.allclose(stats["xyz"][j], [pos[p, 0], pos[p, 1], pos[p, 2], yaw[p, h]])
    par = datasets.collect_familiarity_dataset(route, eye, world, sky, method="parallel", nb_orientations=3, nb_parallel=3)
    assert par["ommatidia"].shape == (3 * L * 3, 300, 3)
    blk = datasets.collect_familiarity_dataset(route, eye, world, sky, method="parallel", nb_orientations=3, nb_parallel=3,
                                               block=(2, 5))
    assert np.array_equal(blk["ommatidia"], par["ommatidia"][2 * 3:5 * 3])
    # streamed in small chunks straight into the file, nothing kept in memory; one-channel PN variant
    st = datasets.collect_familiarity_dataset(route, eye, world, sky, method="grid", nb_orientations=4, nb_rows=3, nb_cols=2,
                                              filename=str(tmp_path / "streamed"), chunk_views=8, keep=False, compressed=False)
    assert st["ommatidia"] is None
    with np.load(str(tmp_path / "streamed.npz")) as f:
        assert np.array_equal(f["ommatidia"], stats["ommatidia"]) and np.array_equal(f["xyz"], stats["xyz"])
    pn = datasets.collect_familiarity_dataset(route, eye, world, sky, method="grid", nb_orientations=4, nb_rows=3, nb_cols=2,
                                              channels=1)
    assert pn["ommatidia"].shape == (24, 300)
    assert np.abs(pn["ommatidia"] - np.clip(stats["ommatidia"].mean(axis=2), 0, 1)).max() < 1e-6
    # against the reference composition under the uniform sky of the data-collection simulations
    eye.xyz, eye.ori = pos[2], R.from_euler('Z', yaw[2, 1], degrees=True)
    ref = _oracle_responses(eye, world, None, pos[2], eye.ori, uniform=10.)
    assert np.abs(stats["ommatidia"][2 * 4 + 1] - ref).max() < 1e-5


def test_familiarity_of_a_small_grid_on_the_gpu(ib):
    """N3 end to end on the device: dataset (N2) -> PN responses -> PCA whitening -> perfect memory -> heat map."""
    from invertsy_b200 import familiarity as fam
    world = ib.Seville2009()
    eye = ib.CompoundEye(nb_input=400, omm_rho=np.deg2rad(4.), omm_res=5., c_sensitive=[0, 0., 1., 0., 0.])
    route = ib.Seville2009.load_routes(degrees=True)["path"][0][::8]      # 102 views: a well-conditioned calibration set
    stats = ib.datasets.collect_familiarity_dataset(route, eye, world, sky=ib.UniformSky(10.), method="grid",
                                                    nb_orientations=4, nb_rows=5, nb_cols=5)
    views = torch.as_tensor(stats["ommatidia"]).cuda()
    rviews = torch.as_tensor(stats["ommatidia_out"]).cuda()
    wh = fam.Whitening(nb_output=8).fit(fam.pn_responses(rviews))
    f = fam.evaluate(views, rviews, whitening=wh)
    assert f.is_cuda and f.shape == (5 * 5 * 4,) and torch.isfinite(f).all() and (f <= 1 + 1e-6).all()
    # the same on the CPU in float64
    ref = fam.evaluate(views.cpu().double(), rviews.cpu().double(),
                       whitening=fam.Whitening(nb_output=8, dtype=torch.float64).fit(fam.pn_responses(rviews.cpu().double())))
    assert np.abs(f.cpu().numpy() - ref.numpy()).max() < 5e-3, np.abs(f.cpu().numpy() - ref.numpy()).max()
    # a route view is fully familiar with itself
    assert np.allclose(fam.evaluate(rviews, rviews, whitening=wh).cpu().numpy(), 1.0, atol=1e-3)
    assert fam.familiarity_map(f, 5, 5, 4).shape == (5, 5, 4)


def _pol_term(rho, p, a, phi):
    return rho * (np.sin(a - phi) ** 2 + np.cos(a - phi) ** 2 * (1. - p) ** 2) + (1. - rho)


@pytest.mark.parametrize("yaw_only", [True, False])
def test_polarisation_sensor_and_sky_only_eye_calls(ib, yaw_only):
    """`pol_sensor(sky=, scene=)` (agent.py:879) and `eye(sky=sky)` (examples/test_sky.py:22): the device's responses
    against the NumPy statement of the documented photoreceptor model on the oracle's sky (Y, DoP, AoP)"""
    ths, phs = np.deg2rad(40.), 0.7
    sky = ib.Sky(ths, phs)
    ori = R.from_euler('Z', 33., degrees=True) if yaw_only else R.from_euler('ZYX', [33., 6., -4.], degrees=True)
    # --- dorsal-rim polarisation sensor, sky alone and under a world (Seville: nothing reaches that high)
    pol = ib.PolarisationSensor(nb_input=60, field_of_view=56, degrees=True, omm_res=0.8)
    pol.ori = ori
    pol.xyz = [5., 5., 0.01]
    r = pol(sky=sky)
    assert r.shape == (60, 1) and r.dtype == np.float32
    glob = ori * pol.omm_ori
    y, p, a = o.sky_render(ths, phs, glob)
    d = glob.apply([1, 0, 0])
    phi = np.arctan2(d[:, 1], d[:, 0])
    want = 0.8 * np.sqrt(y * _pol_term(1.0, p, a, phi))
    assert np.abs(r[:, 0] - want).max() < 1e-5
    assert np.abs(pol(sky=sky, scene=ib.Seville2009())[:, 0] - want).max() < 1e-5
    assert r.std() > 0.01                                      # the filter does modulate the responses
    # --- a polarisation-sensitive compound eye looking at the sky alone: colour responses times the term
    eye = ib.CompoundEye(nb_input=400, omm_pol_op=0.6, omm_res=1.2, c_sensitive=[0, 0.3, 1., 0.5, 0])
    eye.ori = ori
    r = eye(sky=sky)
    glob = ori * eye.omm_ori
    y, p, a = o.sky_render(ths, phs, glob)
    d = glob.apply([1, 0, 0])
    phi = np.arctan2(d[:, 1], d[:, 0])
    yaw_g, pitch_g, _ = glob.as_euler('ZYX').T
    col = o.luminance2skycolour(y) / 255. * _pol_term(0.6, p, a, phi)[:, None]
    col[pitch_g >= 0] = np.array(o.GROUND_COLOUR) / 255.      # looking down: the ground colour (world.py:90), no sky light
    want = np.clip(col * eye.c_sensitive[1:4] * 1.2, 0, 1)
    assert r.shape == (400, 3) and np.abs(r - want).max() < 1e-5
    # batches (rows mode under shared headings) give what the single calls give
    if yaw_only:
        pos = np.tile([[5., 5., 0.01]], (130, 1))
        yw = np.tile(np.arange(36) * 10. - 180., (130, 1))
        pn = eye.render_batch(ib.sense._empty_world(0), sky, pos, yw, channels=1)
        eye.ori = R.from_euler('Z', yw[0, 7], degrees=True)
        single = eye(sky=sky)
        assert np.abs(pn[7] - np.clip(single.mean(axis=1), 0, 1)).max() < 1e-6
        assert np.abs(pn[36 * 129 + 7] - pn[7]).max() == 0
