"""
TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference
(``invertsy.env.world`` / ``invertsy.env.sky`` imported from /root/reference through
oracle/ref_loader.py).  Run in the build container only:

    python oracle/gen_golden.py

The fixtures hold inputs + reference outputs + the sha256 of the reference files they came
from, so the GPU box (no /root/reference) can check parity against the real reference.
Hit indices are obtained without touching reference code by rendering with index-encoded
colours: colours[t] = t+1, so rgb[:,0]-1 is the painter's winning triangle.
"""
import json
import os
import sys

import numpy as np
from scipy.spatial.transform import Rotation as R

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import isy_oracle as o  # noqa: E402
from ref_loader import load_reference, reference_file_hashes  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
HEADINGS = (0., 33.3, -139.65, 179.9)


def ref_hits(refworld, world, pos, ori):
    T = world.polygons.shape[0]
    idxcol = np.repeat((np.arange(T) + 1)[:, None], 3, 1).astype('float32')
    wi = refworld.WorldBase(world.polygons, idxcol, world.horizon)
    c = wi(pos, ori)
    first = c[:, 0]
    return np.where(np.isnan(first) | (first < 1), -1, first - 1).astype(np.int32)


def n_visible(world, pos):
    xyz = world.polygons - np.nanmean(pos, axis=0)
    return int((np.min(np.linalg.norm(xyz, axis=2), axis=1) <= world.horizon).sum())


def world_set(refworld, world, poses, omm_ori, pos_dtype=np.float64):
    rgb, hit, nvis, q = [], [], [], []
    for x, y, z, yaw_deg in poses:
        pos = np.array([[x, y, z]], dtype=pos_dtype)
        ori = o.view_rotations(yaw_deg, omm_ori)
        rgb.append(world(pos, ori))
        hit.append(ref_hits(refworld, world, pos, ori))
        nvis.append(n_visible(world, pos))
        yw, pt, _ = ori.as_euler('ZYX').T
        q.append(np.c_[pt, yw])
    return dict(poses=np.asarray(poses, dtype=np.float64), rgb=np.asarray(rgb), hit=np.asarray(hit),
                nvis=np.asarray(nvis), query=np.asarray(q)[:4])


def main():
    os.makedirs(OUT, exist_ok=True)
    refworld, refsky = load_reference()
    rng = np.random.RandomState(2021)

    seville = refworld.Seville2009()
    sparse = refworld.SimpleWorld()
    route = refworld.Seville2009.load_routes(degrees=True)['path'][0]

    pitch, yaw = o.fibonacci_eye(1000)
    omm = o.eye_rotations(pitch, yaw)

    # ---- poses: route steps, arena corners / edges, densest and emptiest spots ----
    poses = [tuple(route[k]) for k in range(0, 812, 40)]
    k = 0
    for x in (0., 5., 10.):
        for y in (0., 5., 10.):
            poses.append((x, y, 0.01, HEADINGS[k % 4]))
            k += 1
    gx, gy = np.meshgrid(np.linspace(0, 10, 41), np.linspace(0, 10, 41))
    nv = np.array([n_visible(seville, np.array([[a, b, 0.01]])) for a, b in zip(gx.ravel(), gy.ravel())])
    for idx in (nv.argmax(), nv.argmin()):
        poses.append((gx.ravel()[idx], gy.ravel()[idx], 0.01, HEADINGS[k % 4]))
        k += 1
    for _ in range(5):
        poses.append((rng.uniform(0, 10), rng.uniform(0, 10), 0.01, rng.uniform(-180, 180)))

    g = world_set(refworld, seville, poses, omm)
    np.savez_compressed(os.path.join(OUT, "world_seville.npz"), n_omm=1000, **g)
    print("seville", g["rgb"].shape, "nvis", g["nvis"].min(), g["nvis"].max(), "hits", (g["hit"] >= 0).mean())

    sposes = [(rng.uniform(0.5, 9.5), rng.uniform(0.5, 9.5), 0.01, rng.uniform(-180, 180)) for _ in range(10)]
    sposes.append((0.1, 0.1, 0.01, 12.))      # V = 0 corner
    g = world_set(refworld, sparse, sposes, omm)
    np.savez_compressed(os.path.join(OUT, "world_sparse.npz"), n_omm=1000, **g)
    print("sparse", g["rgb"].shape, "nvis", g["nvis"].min(), g["nvis"].max(), "hits", (g["hit"] >= 0).mean())

    # ---- float32 positions (reference then computes vertex angles in float32: documented deviation) ----
    g = world_set(refworld, seville, poses[:4], omm, pos_dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "world_seville_pos32.npz"), n_omm=1000, **g)

    # ---- dense elevation band: 20000 rays between -20 deg and +45 deg elevation (many edge-near rays) ----
    nb = 20000
    i = np.arange(nb)
    byaw = (np.pi * (1 + np.sqrt(5.)) * (i + .5)) % (2 * np.pi) - np.pi
    z = np.sin(np.deg2rad(-20.)) + (np.sin(np.deg2rad(45.)) - np.sin(np.deg2rad(-20.))) * (i + .5) / nb
    bpitch = -np.arcsin(z)
    bomm = o.eye_rotations(bpitch, byaw)
    bposes = [tuple(route[0]), tuple(route[400]), poses[-3]]
    hits = []
    for x, y, zz, yaw_deg in bposes:
        pos = np.array([[x, y, zz]])
        hits.append(ref_hits(refworld, seville, pos, o.view_rotations(yaw_deg, bomm)))
    np.savez_compressed(os.path.join(OUT, "world_band.npz"), poses=np.asarray(bposes), pitch=bpitch, yaw=byaw,
                        hit=np.asarray(hits))
    print("band hits", (np.asarray(hits) >= 0).mean())

    # ---- sky: 1000-direction sphere eye, 4 x 3 suns, one heading each ----
    suns, Y, P, A, Yi = [], [], [], [], []
    k = 0
    irgbu = np.array([[0., 0., 1., 0., 0.]])
    for ths in (5., 30., 60., 85.):
        for phs in (0., 90., 180.):
            sky = refsky.Sky(np.deg2rad(ths), np.deg2rad(phs))
            ori = o.view_rotations(HEADINGS[k % 4], omm)
            y, p, a = sky(ori)
            yi, _, _ = sky(ori, irgbu=irgbu)
            suns.append((np.deg2rad(ths), np.deg2rad(phs), HEADINGS[k % 4]))
            Y.append(y), P.append(p), A.append(a), Yi.append(yi)
            k += 1
    # 5000-direction dorsal (upper hemisphere) eye as in examples/test_sky.py
    hp, hy = o.fibonacci_eye(5000, hemisphere=True)
    homm = o.eye_rotations(hp, hy)
    sky = refsky.Sky(np.deg2rad(60), np.pi)
    hY, hP, hA = sky(homm)
    coef = np.array([sky.A, sky.B, sky.C, sky.D, sky.E])
    np.savez_compressed(os.path.join(OUT, "sky.npz"), suns=np.asarray(suns), Y=np.asarray(Y), P=np.asarray(P),
                        A=np.asarray(A), Y_irgbu=np.asarray(Yi), irgbu=irgbu, hemi_Y=hY, hemi_P=hP, hemi_A=hA,
                        coef=coef, tau_L=sky.tau_L, Y_z=sky.Y_z, M_p=sky.M_p)

    # ---- world painted over the sky background (init_c -> float64 output on the 0..255 scale) ----
    sky = refsky.Sky(np.deg2rad(60), np.pi)
    rgb64, ys = [], []
    for x, y, zz, yaw_deg in poses[:3]:
        ori = o.view_rotations(yaw_deg, omm)
        lum, _, _ = sky(ori)
        rgb64.append(seville(np.array([[x, y, zz]]), ori, init_c=lum))
        ys.append(lum)
    usky = refsky.UniformSky(10.)
    uy, up, ua = usky(o.view_rotations(poses[0][3], omm))
    np.savez_compressed(os.path.join(OUT, "world_initc.npz"), poses=np.asarray(poses[:3]), rgb=np.asarray(rgb64),
                        init_c=np.asarray(ys), uniform_Y=uy, uniform_P=up, uniform_A=ua)

    # ---- sun ephemeris (row N4): reference Sun.compute on random places / times ----
    import importlib
    from datetime import datetime, timedelta
    eph = importlib.import_module('invertsy.env.ephemeris')
    obsm = importlib.import_module('invertsy.env.observer')
    erng = np.random.RandomState(4)
    rows = []
    for _ in range(200):
        d = datetime(2021, 1, 1) + timedelta(seconds=int(erng.uniform(0, 365 * 86400)))
        lon, lat = erng.uniform(-np.pi, np.pi), erng.uniform(-1.4, 1.4)
        ob = obsm.Observer(lon=lon, lat=lat, date=d)
        sun = eph.Sun()
        sun.compute(ob)
        rows.append((d.year, d.month, d.day, d.hour, d.minute, d.second, lon, lat, ob.tzgmt, sun.alt, sun.az))
    dsk = refsky.Sky.from_observer()
    np.savez_compressed(os.path.join(OUT, "ephemeris.npz"), rows=np.asarray(rows, dtype=np.float64),
                        default_sky=np.array([dsk.theta_s, dsk.phi_s]))

    import scipy
    meta = dict(reference_sha256=reference_file_hashes(), numpy=np.__version__, scipy=scipy.__version__,
                note="generated by oracle/gen_golden.py from the unmodified reference")
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
