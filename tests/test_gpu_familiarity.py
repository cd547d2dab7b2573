"""
GPU: the perfect-memory familiarity kernels (csrc/isy_fam.cuh: TF32 tcgen05 filter + exact fp32 re-check) through the
C ABI, against the NumPy float64 statement of the definition (`familiarity.reference_familiarity`):
familiarity(q) = 1 - min_m sqrt(mean_k (q_k - m_k)^2).  Tolerance: 2e-6 absolute on the familiarity (fp32 sums of up to
a few thousand squared differences + one sqrt); the nearest index must be the true one wherever the runner-up is further
than 1e-6 in mean squared difference.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-6


@pytest.fixture(scope="module")
def fam():
    from invertsy_b200 import familiarity
    assert torch.cuda.is_available()
    return familiarity


def _check(fam, q, m, via="cuda"):
    ref, ref_i = fam.reference_familiarity(q, m)
    mem = fam.PerfectMemory().store(m)
    if via == "cuda":
        f, i = mem.familiarity(torch.as_tensor(q, dtype=torch.float32).cuda(), nearest=True)
        torch.cuda.synchronize()
        assert 0 <= mem.unresolved() <= q.shape[0]        # (diagnostic count: four candidates inside the error margin)
    else:
        f, i = mem.familiarity(q, nearest=True)
    f, i = f.cpu().numpy(), i.cpu().numpy()
    assert np.abs(f - ref).max() < TOL, np.abs(f - ref).max()
    bad = np.nonzero(i != ref_i)[0]
    if bad.size:   # only genuine near-ties may differ
        d = ((q[bad].astype(np.float64) - m[i[bad]].astype(np.float64)) ** 2).mean(1)
        d0 = ((q[bad].astype(np.float64) - m[ref_i[bad]].astype(np.float64)) ** 2).mean(1)
        assert np.abs(d - d0).max() < 1e-6
    return f, i


@pytest.mark.parametrize("Q,M,K", [(1000, 812, 1000), (128, 256, 32), (1, 1, 4), (129, 257, 36), (300, 3, 1000),
                                    (5000, 1500, 600), (77, 812, 1001)])
def test_random_vectors_match_the_definition(fam, Q, M, K):
    rng = np.random.RandomState(Q + M + K)
    m = rng.uniform(0, 1, (M, K)).astype(np.float32)
    q = rng.uniform(0, 1, (Q, K)).astype(np.float32)
    n = min(Q, M)
    q[:n:5] = m[:n:5]                                   # some queries are stored vectors: familiarity exactly 1
    f, i = _check(fam, q, m)
    assert np.all(f[:n:5] == 1.0) and np.array_equal(i[:n:5], np.arange(0, n, 5))


def test_route_like_memory_with_near_ties(fam):
    """stored vectors that differ little from their neighbours (consecutive route views) and queries between them"""
    rng = np.random.RandomState(3)
    M, K, Q = 812, 1000, 4096
    base = rng.uniform(0, 1, K)
    walk = np.cumsum(rng.normal(0, 0.01, (M, K)), axis=0)
    m = np.clip(base + walk, 0, 1).astype(np.float32)
    j = rng.randint(0, M - 1, Q)
    t = rng.uniform(0, 1, (Q, 1))
    q = np.clip((1 - t) * m[j] + t * m[j + 1] + rng.normal(0, 0.02, (Q, K)), 0, 1).astype(np.float32)
    _check(fam, q, m)


def test_host_buffers_and_strided_device_queries(fam):
    rng = np.random.RandomState(4)
    m = rng.uniform(0, 1, (100, 250)).astype(np.float32)
    q = rng.uniform(0, 1, (333, 250)).astype(np.float32)
    a, _ = _check(fam, q, m, via="host")
    mem = fam.PerfectMemory().store(m)
    wide = torch.zeros((333, 256), dtype=torch.float32, device="cuda")     # row stride 256 floats, 250 used
    wide[:, :250] = torch.as_tensor(q)
    b = mem.familiarity(wide[:, :250]).cpu().numpy()
    c = mem.familiarity(torch.as_tensor(q, dtype=torch.float64).cuda()).cpu().numpy()   # converted copy
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_rendered_views_end_to_end(fam):
    """N2 -> N3 on the device: PN inputs written by the render kernel (one-channel `resp`) straight into the memory"""
    import invertsy_b200 as ib
    world = ib.Seville2009()
    eye = ib.CompoundEye(nb_input=1000, omm_res=5., c_sensitive=[0, 0., 1., 0., 0.])
    route = ib.Seville2009.load_routes(degrees=True)["path"][0][::4]
    r = ib.Renderer(world, eye._eye_layout(), sky=ib.UniformSky(10.))
    s1 = eye.sensor_params(1)
    stored = r.render(route[:, :3], np.deg2rad(route[:, 3:4]), outputs=("resp",), sensor=s1)["resp"]
    pos, yaw = ib.views.grid_views(6, 6, 8)
    views = r.render(pos, np.deg2rad(yaw), outputs=("resp",), sensor=s1)["resp"]
    mem = fam.PerfectMemory().store(stored)
    f, i = mem.familiarity(views, nearest=True)
    torch.cuda.synchronize()
    ref, ref_i = fam.reference_familiarity(views.cpu().numpy(), stored.cpu().numpy())
    assert np.abs(f.cpu().numpy() - ref).max() < TOL
    assert (i.cpu().numpy() != ref_i).mean() < 0.01
    assert np.allclose(mem.familiarity(stored).cpu().numpy(), 1.0, atol=1e-6)
    assert fam.familiarity_map(f, 6, 6, 8).shape == (6, 6, 8)


def test_duplicate_stored_vectors_are_kept_once_and_ties_go_to_the_lowest_index(fam):
    rng = np.random.RandomState(9)
    base = rng.uniform(0, 1, (40, 64)).astype(np.float32)
    m = base[rng.randint(0, 40, 500)]                       # 500 stored vectors, at most 40 different ones
    q = np.clip(base[rng.randint(0, 40, 300)] + rng.normal(0, 0.01, (300, 64)), 0, 1).astype(np.float32)
    f, i = _check(fam, q, m)
    ref, ref_i = fam.reference_familiarity(q, m)
    assert np.array_equal(i, ref_i)                          # np.argmin: the first of the copies
    mem = fam.PerfectMemory().store(m)
    assert mem.unique_rows().size == np.unique(m, axis=0).shape[0] <= 40
    mem.familiarity(torch.as_tensor(q).cuda())
    torch.cuda.synchronize()
    assert mem.unresolved() == 0


@pytest.mark.parametrize("nb_pn,nb_kc,fan_in,sparse", [(1000, 40000, 8, 0.01), (64, 2000, 5, 0.05), (300, 4099, 3, 0.001)])
def test_willshaw_memory_matches_its_numpy_statement(fam, nb_pn, nb_kc, fan_in, sparse):
    rng = np.random.RandomState(nb_kc)
    stored = rng.uniform(0, 1, (60, nb_pn)).astype(np.float32)
    queries = rng.uniform(0, 1, (90, nb_pn)).astype(np.float32)
    queries[:20] = stored[:20]                                        # learnt views: familiarity exactly 1
    queries[20:40] = np.clip(stored[20:40] + rng.normal(0, 0.02, (20, nb_pn)), 0, 1)   # near a learnt view
    queries[40:50] = np.round(queries[40:50] * 4) / 4                 # coarse values: many tied KC inputs
    mem = fam.WillshawMemory(nb_pn, nb_kc, fan_in=fan_in, sparseness=sparse, seed=7).store(stored)
    got = mem.familiarity(queries).cpu().numpy()
    ref = fam.reference_willshaw(stored, queries, nb_kc, fan_in, sparse, seed=7)
    assert mem.nb_active == max(1, round(sparse * nb_kc))
    assert np.abs(got - ref).max() < 1e-6, np.abs(got - ref).max()
    assert np.all(got[:20] == 1.0) and got[50:].mean() < got[20:40].mean() < 1.0
    # storing in two calls is the same as in one; reset forgets
    two = fam.WillshawMemory(nb_pn, nb_kc, fan_in=fan_in, sparseness=sparse, seed=7).store(stored[:25]).store(stored[25:])
    assert np.array_equal(two.familiarity(queries).cpu().numpy(), got)
    assert np.all(mem.reset().familiarity(queries[:5]).cpu().numpy() == 0.0)


def test_heat_map_call_single_gpu(fam):
    """`familiarity.HeatMap` (the call bench.py times as e2e): host poses in -> render -> perfect memory -> map out, against
    the float64 statement on the rendered responses; one process = the whole grid, no collective"""
    import invertsy_b200 as ib
    from invertsy_b200 import _lib
    world = ib.Seville2009()
    pitch, yaw = ib.fibonacci_lattice(1000)
    r = ib.Renderer(world, ib.EyeLayout.from_angles(pitch, yaw), sky=ib.Sky(np.deg2rad(60.), np.pi))
    sensor = _lib.SensorParams.make((0., 1., 0.), 1., 1)
    route = ib.Seville2009.load_routes(degrees=True)["path"][0][::5]
    stored = r.render(route[:, :3], np.deg2rad(route[:, 3:4]), outputs=("resp",), sensor=sensor)["resp"]
    memory = fam.PerfectMemory().store(stored)
    pos, yaw_deg = ib.views.grid_views(12, 11, 36)                 # 132 positions x 36 headings: rows mode
    hm = fam.HeatMap(r, memory, sensor, nb_positions=len(pos), nb_headings=36)
    mine = hm.positions()
    assert np.array_equal(mine, np.arange(len(pos)))
    out = hm(pos[mine], np.deg2rad(yaw_deg[mine]))
    assert out.shape == (len(pos) * 36,) and out.is_pinned()
    views = hm.views["resp"].cpu().numpy()
    ref, _ = fam.reference_familiarity(views, stored.cpu().numpy())
    assert np.abs(out.numpy() - ref).max() < TOL
    # the views it rendered are the per-view renders
    single = r.render(pos[5:6], np.deg2rad(yaw_deg[5:6]), outputs=("resp",), sensor=sensor, per_view_sky=True)["resp"].cpu().numpy()
    assert np.abs(views[5 * 36:6 * 36] - single).max() < 1e-6
    assert fam.familiarity_map(out, 12, 11, 36).shape == (12, 11, 36)


def test_bad_arguments_are_refused(fam):
    from invertsy_b200 import _lib
    mem = fam.PerfectMemory().store(np.zeros((3, 8), dtype=np.float32))
    with pytest.raises(ValueError):
        mem.familiarity(torch.zeros((2, 7), device="cuda"))
    with pytest.raises(_lib.IsyError):
        fam.WillshawMemory(100000, 100000)._handle()              # does not fit one SM's shared memory
    with pytest.raises(_lib.IsyError):
        fam.WillshawMemory(10, 100, sparseness=0.)._handle()
    empty = mem.familiarity(torch.zeros((0, 8), device="cuda"))
    assert empty.shape == (0,)
