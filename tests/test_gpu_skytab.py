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


@pytest.mark.parametrize("sky", ["perez", "uniform", "none"])
@pytest.mark.parametrize("H,R", [(36, 1000), (5, 999)])
def test_photoreceptor_rows_match_the_per_view_passes_and_the_reference_composition(ib, sky, H, R):
    """the `resp` output (PN inputs, one channel) in rows mode -- template / table rows + patches -- against the
    per-view final pass, the three-channel output, the fp64 generic pass, and the NumPy statement of the transform
    on the oracle's world / sky"""
    from invertsy_b200 import _lib
    world = ib.Seville2009()
    pitch, yaw = o.fibonacci_eye(R)
    ths, phs = np.deg2rad(60.), np.pi
    sk = {"perez": ib.Sky(ths, phs), "uniform": ib.UniformSky(10.), "none": None}[sky]
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sk)
    P = max(150, -(-4200 // H))
    pos = _positions(P, 36)
    heads = (np.arange(H) * 2 * np.pi / H + 0.2 + np.pi) % (2 * np.pi) - np.pi
    yw = np.tile(heads, (P, 1))
    w_rgb, gain = (0.2, 1.0, 0.1), 1.5
    s1, s3 = _lib.SensorParams.make(w_rgb, gain, 1), _lib.SensorParams.make(w_rgb, gain, 3)
    outs = ("rgb", "hit", "resp") + (("Y",) if sk is not None else ())
    a = r.render(pos, yw, outputs=outs, sensor=s1)                          # rows mode
    b = r.render(pos, yw, outputs=outs, sensor=s1, per_view_sky=True)       # per-view fast pass
    c = r.render(pos, yw, outputs=("resp",), sensor=s3)                     # three channels (per-view pass)
    d = r.render(pos[:20], yw[:20], outputs=("resp",), sensor=s3, f64=True) # generic pass, fp64
    torch.cuda.synchronize()
    r.check()
    ra, rb, rc, rd = (x["resp"].cpu().numpy() for x in (a, b, c, d))
    assert ra.shape == (P * H, R) and rc.shape == (P * H, R, 3) and rd.dtype == np.float64
    assert np.array_equal(a["hit"].cpu().numpy(), b["hit"].cpu().numpy())
    assert np.abs(ra - rb).max() < 1e-6
    assert np.abs(ra - np.clip(rc.mean(axis=2), 0, 1)).max() < 1e-6
    assert np.abs(rc[:20 * H] - rd).max() < 1e-6
    # reference composition on a few views
    omm = o.eye_rotations(pitch, yaw)
    for v in (0, (P // 2) * H + H // 3, P * H - 1):
        ori = o.view_rotations(np.rad2deg(heads[v % H]), omm)
        col = o.world_render(world.polygons, world.colours, 2.0, pos[v // H].reshape(1, 3), ori).astype(np.float64)
        bg = np.zeros_like(col)
        if sky == "perez":
            bg = o.luminance2skycolour(o.sky_render(ths, phs, ori)[0]) / 255.
        elif sky == "uniform":
            bg = o.luminance2skycolour(np.full(R, 10.)) / 255.
        col = np.where(np.isnan(col), bg, col)
        want = np.clip(col * np.asarray(w_rgb) * gain, 0, 1)
        assert np.abs(rc[v] - want).max() < 1e-5
        assert np.abs(ra[v] - np.clip(want.mean(axis=1), 0, 1)).max() < 1e-5


def test_status_of_a_workspace_is_sticky_across_launches(ib):
    """an overflow raised by one launch survives later launches on the same workspace until it is read
    (include/isy_b200.h, isy_workspace_status); reading clears it"""
    rng = np.random.RandomState(7)
    T = 135000                                           # slab: min(2 T, 1 << 17) = 131 072 copies per position
    base = np.c_[rng.uniform(4.5, 5.5, T), rng.uniform(4.5, 5.5, T), np.zeros(T)]
    tri = np.stack([base, base + [0.01, 0., 0.], base + [0., 0.01, 0.05]], axis=1).astype(np.float32)
    world = ib.WorldBase(tri, np.full((T, 3), 0.5, dtype=np.float32), horizon=3.)
    pitch, yaw = o.fibonacci_eye(64)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw))
    seam = np.array([[5.0, 5.0, 0.01]])                  # sees all 135 000 triangles: more than the slab holds
    calm = np.array([[-40.0, 5.0, 0.01]])                # sees nothing
    r.render(seam, np.zeros((1, 1)), outputs=("hit",))
    r.render(calm, np.zeros((1, 1)), outputs=("hit",))   # a later, clean launch on the same workspace
    with pytest.raises(_lib_error(ib)):
        r.check()
    r.check()                                            # read = cleared
    r.render(calm, np.zeros((1, 1)), outputs=("hit",))
    r.check()


def _lib_error(ib):
    from invertsy_b200 import _lib
    return _lib.IsyError
