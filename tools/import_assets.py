"""
Converts the AntWorld assets shipped with the reference (MATLAB .mat files) into compact
.npz files under invertsy_b200/data/, so that the GPU box (which has no /root/reference)
can run the named worlds.  Data only -- no reference source is copied.

  python tools/import_assets.py [--reference /root/reference]

Layout written (the raw .mat arrays, untouched, float64):
  seville2009_world.npz : X, Y, Z (5000,3), colp (5000,3)        <- data/Seville2009_world/world5000_gray.mat
  sparse_world.npz      : X, Y, Z (3222,3)                         <- data/Sparse_world/world_gray.mat
  seville2009_routes.npz: keys "Ant{a}_Route{r}" (n,3) [x_cm, y_cm, phi_deg]   <- AntRoutes.mat
The unit/axis conventions of invertsy.env.world loaders (env/world.py:195-273, 420-462) are
applied at load time by invertsy_b200.env.world, not here.
"""
import argparse
import os

import numpy as np
from scipy.io import loadmat

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "invertsy_b200", "data")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    data = os.path.join(args.reference, "data")
    os.makedirs(OUT, exist_ok=True)

    m = loadmat(os.path.join(data, "Seville2009_world", "world5000_gray.mat"))
    np.savez_compressed(os.path.join(OUT, "seville2009_world.npz"),
                        X=m["X"], Y=m["Y"], Z=m["Z"], colp=m["colp"])

    m = loadmat(os.path.join(data, "Sparse_world", "world_gray.mat"))
    np.savez_compressed(os.path.join(OUT, "sparse_world.npz"), X=m["X"], Y=m["Y"], Z=m["Z"])

    m = loadmat(os.path.join(data, "Seville2009_world", "AntRoutes.mat"))
    routes = {k: np.asarray(v, dtype=np.float64) for k, v in m.items() if k.startswith("Ant")}
    np.savez_compressed(os.path.join(OUT, "seville2009_routes.npz"), **routes)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
