"""
CPU, build container only: pins the oracle, the loaders and the view generators to the
UNMODIFIED reference imported from /root/reference (skipped where it is absent, e.g. the GPU box;
the committed goldens carry the same evidence there).
"""
import ast
import os

import numpy as np
import pytest

import isy_oracle as o
from ref_loader import DEFAULT_REFERENCE_ROOT, load_reference, reference_available, reference_file_hashes

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="reference checkout not present")]


@pytest.fixture(scope="module")
def ref():
    return load_reference()


def test_golden_meta_matches_reference_files():
    import json
    from conftest import GOLDEN
    meta = json.load(open(os.path.join(GOLDEN, "meta.json")))
    assert meta["reference_sha256"] == reference_file_hashes()


def test_loaders_equal_reference_loaders(ref):
    import invertsy_b200 as ib
    refworld, _ = ref
    for mine, theirs in ((ib.Seville2009, refworld.Seville2009), (ib.SimpleWorld, refworld.SimpleWorld)):
        p1, c1 = mine.load_world()
        p2, c2 = theirs.load_world()
        assert p1.dtype == p2.dtype and np.array_equal(p1, p2) and np.array_equal(c1, c2)
    for deg in (True, False):
        r1 = ib.Seville2009.load_routes(degrees=deg)
        r2 = refworld.Seville2009.load_routes(degrees=deg)
        assert r1["ant_no"] == r2["ant_no"] and r1["route_no"] == r2["route_no"]
        assert all(np.array_equal(a, b) for a, b in zip(r1["path"], r2["path"]))


def test_oracle_bit_identical_on_fresh_random_poses(ref):
    refworld, refsky = ref
    rng = np.random.RandomState(7)
    pitch, yaw = o.fibonacci_eye(600)
    omm = o.eye_rotations(pitch, yaw)
    for world in (refworld.Seville2009(), refworld.SimpleWorld()):
        for _ in range(3):
            pos = np.array([[rng.uniform(0, 10), rng.uniform(0, 10), 0.01]])
            ori = o.view_rotations(rng.uniform(-180, 180), omm)
            a = world(pos, ori)
            b = o.world_render(world.polygons, world.colours, world.horizon, pos, ori)
            assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True)
            ths, phs = rng.uniform(0.05, 1.5), rng.uniform(-np.pi, np.pi)
            sky = refsky.Sky(ths, phs)
            for irgbu in (None, np.array([[0., 0., 1., 0., 0.]]), np.array([[.1, .2, .3, -1., .4]])):
                ya, pa, aa = sky(ori, irgbu=irgbu)
                yb, pb, ab = o.sky_render(ths, phs, ori, irgbu=irgbu)
                assert np.array_equal(ya, yb) and np.array_equal(pa, pb) and np.array_equal(aa, ab, equal_nan=True)
            eta = rng.rand(600) < 0.2
            ya, pa, aa = sky(ori, eta=eta.copy())
            yb, pb, ab = o.sky_render(ths, phs, ori, eta=eta.copy())
            assert np.array_equal(ya, yb) and np.array_equal(pa, pb) and np.array_equal(aa, ab, equal_nan=True)
            c = world(pos, ori, init_c=ya, eta=np.repeat(eta[:, None], 3, 1))
            d = o.world_render(world.polygons, world.colours, world.horizon, pos, ori, init_c=yb,
                               eta=np.repeat(eta[:, None], 3, 1))
            assert np.array_equal(c, d, equal_nan=True)


def _reference_functions(relpath, names):
    """exec selected top-level functions of a reference file that cannot be imported as a module"""
    src = open(os.path.join(DEFAULT_REFERENCE_ROOT, relpath)).read()
    tree = ast.parse(src)
    ns = {"np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), relpath, "exec"), ns)
    return ns


def test_view_generators_equal_reference_index_maths():
    from invertsy_b200 import views
    ns = _reference_functions("src/invertsy/sim/_helpers.py", {"col2x", "row2y", "ori2yaw"})
    for n in (100, 200, 7):
        for i in range(0, n, max(1, n // 13)):
            assert views.col2x(i, n) == ns["col2x"](i, nb_cols=n, max_meters=10.)
            assert views.row2y(i, n) == ns["row2y"](i, nb_rows=n, max_meters=10.)
    for n in (16, 36, 72):
        for i in range(n):
            assert views.ori2yaw(i, n, degrees=True) == ns["ori2yaw"](i, nb_oris=n, degrees=True)
            assert views.ori2yaw(i, n, degrees=False) == ns["ori2yaw"](i, nb_oris=n, degrees=False)
