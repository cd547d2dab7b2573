"""
Measures, on a sample of the heat-map grid (BASELINE.json configs[2]), how often the reference's answer changes when the
caller passes float32 `pos` -- the reference then evaluates world.py:94-123 in float32 (`polygons - nanmean(pos)` keeps
the dtype), while this library always promotes the position to float64.  Uses the NumPy port of the reference with
index-encoded colours (the painter's winner per ray), float32 positions vs the same positions in float64.

    python tools/float32_pos_mismatch.py [--positions 64] [--procs 8]   ->  JSON line (profiles/r02_float32_pos.json)
"""
import argparse
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def work(args):
    import isy_oracle as o
    from invertsy_b200.env.world import Seville2009
    pos_list, seed = args
    polygons, _ = Seville2009.load_world()
    T = polygons.shape[0]
    idx_col = np.repeat((np.arange(T) + 1)[:, None], 3, 1).astype(np.float32)     # colour = triangle index + 1
    pitch, yaw = o.fibonacci_eye(1000)
    omm = o.eye_rotations(pitch, yaw)
    mism = total = hits = 0
    for (x, y, z) in pos_list:
        for h in range(0, 36, 3):                                                   # 12 of the 36 headings
            ori = o.view_rotations(h * 10. - 180., omm)
            p64 = np.array([[np.float32(x), np.float32(y), np.float32(z)]], dtype=np.float64)   # the same numbers ...
            p32 = p64.astype(np.float32)                                                        # ... as float32
            a = o.world_render(polygons, idx_col, 2.0, p64, ori)[:, 0]
            b = o.world_render(polygons, idx_col, 2.0, p32, ori)[:, 0]
            a, b = np.nan_to_num(a, nan=-1.), np.nan_to_num(b, nan=-1.)
            mism += int((a != b).sum())
            hits += int((a >= 1).sum())
            total += a.size
    return mism, total, hits


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    from invertsy_b200 import views
    pos, _ = views.grid_views(200, 200, 36)
    rng = np.random.RandomState(2021)
    sel = pos[rng.choice(pos.shape[0], a.positions, replace=False)]
    chunks = [(sel[i::a.procs].tolist(), i) for i in range(a.procs)]
    with mp.get_context("fork").Pool(a.procs) as pool:
        res = pool.map(work, chunks)
    mism, total, hits = (sum(r[k] for r in res) for k in range(3))
    print(json.dumps({"positions": a.positions, "headings_per_position": 12, "rays": total, "rays_hitting_vegetation": hits,
                      "rays_whose_winner_changes_with_float32_pos": mism, "fraction_of_all_rays": mism / total,
                      "fraction_of_hit_rays": mism / max(hits, 1)}))
