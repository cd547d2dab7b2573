"""Measures the TF32 filter's score error on rendered views (real data) and random vectors: mean / std / max, against
the round-to-nearest and truncation models."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import invertsy_b200 as ib
from invertsy_b200 import familiarity as fam, _lib

def report(name, q, m):
    mem = fam.PerfectMemory().store(m)
    f = mem.familiarity(torch.as_tensor(q).cuda()); torch.cuda.synchronize()
    cand, score = mem.candidates()
    m = m[mem.unique_rows()]
    print('   unique stored vectors: %d' % m.shape[0])
    q64, m64 = q.astype(np.float64), m.astype(np.float64)
    mu = m64.mean(0); mc = m64 - mu
    exact = (m64[cand] ** 2).sum(-1) - 2 * np.einsum('qk,qck->qc', q64, mc[cand])
    err = score.astype(np.float64) - exact
    s = np.einsum('qk,qck->qc', q64, mc[cand])
    sig = 2 * 2.0 ** -11 * np.sqrt(2 / 3.) * np.sqrt(np.einsum('qk,qck->qc', q64 ** 2, mc[cand] ** 2))
    gap = np.diff(exact, axis=1)
    print(name, "err mean %.3e std %.3e max %.3e | model sigma(RN) median %.3e | err/sigma std %.2f max %.2f | bias vs 2^-9 s: corr %.3f slope %.3e | exact gaps of sorted candidates: median %.3e, p1 %.3e"
          % (err.mean(), err.std(), np.abs(err).max(), np.median(sig), (err / sig).std(), np.abs(err / sig).max(),
             np.corrcoef(err.ravel(), s.ravel())[0, 1], np.polyfit(s.ravel(), err.ravel(), 1)[0], np.median(np.abs(gap)), np.percentile(np.abs(gap), 1)))
    ref, ref_i = fam.reference_familiarity(q[:3000], m)
    print("   max |fam - ref| on 3000 queries: %.2e ; unresolved %d of %d" % (np.abs(f.cpu().numpy()[:3000] - ref).max(), mem.unresolved(), q.shape[0]))

rng = np.random.RandomState(0)
report("random", rng.uniform(0, 1, (20000, 1000)).astype(np.float32), rng.uniform(0, 1, (812, 1000)).astype(np.float32))
world = ib.Seville2009()
eye = ib.EyeLayout.from_angles(*ib.fibonacci_lattice(1000))
r = ib.Renderer(world, eye, sky=ib.Sky(np.deg2rad(60.), np.pi))
sensor = _lib.SensorParams.make((0., 1., 0.), 5., 1)
route = ib.Seville2009.load_routes(degrees=True)["path"][0]
stored = r.render(route[:, :3], np.deg2rad(route[:, 3:4]), outputs=("resp",), sensor=sensor)["resp"].cpu().numpy()
pos, yaw = ib.views.grid_views(200, 200, 36)
sel = rng.choice(pos.shape[0], 600, replace=False)
views = r.render(pos[sel], np.deg2rad(yaw[sel]), outputs=("resp",), sensor=sensor)["resp"].cpu().numpy()
report("rendered", views, stored)
