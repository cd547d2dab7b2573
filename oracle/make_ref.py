"""
TEST / BENCH INFRASTRUCTURE ONLY -- ships the *unmodified* reference implementation of the path to where the GPU box
can run it as the CPU baseline.

`/root/reference` exists only in the build container.  `__graft_entry__.build()` calls `build_ref()` there: it copies
the reference's own files for this path (env/world.py, env/sky.py and the four small modules they import, plus the
Seville2009 world file) byte for byte into `oracle/_ref/`, mirroring the reference's directory layout so that its own
relative data paths resolve.  `oracle/_ref/` is git-ignored (never part of the history) but not gpurun-ignored, so it
travels to the GPU box like the built `.so` files.  `bench.py`'s CPU legs import it through `oracle/ref_loader.py`
(three import shims, no edits) and report `cpu_baseline.kind = "reference"`; without it they fall back to the NumPy
port `oracle/isy_oracle.py` (bit-identical to the reference, `tests/test_oracle_vs_reference.py`), kind "port".
Nothing under `invertsy_b200/` ever touches this directory.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
FILES = ("src/invertsy/__helpers.py", "src/invertsy/env/world.py", "src/invertsy/env/sky.py",
         "src/invertsy/env/_helpers.py", "src/invertsy/env/observer.py", "src/invertsy/env/ephemeris.py",
         "data/Seville2009_world/world5000_gray.mat")


def build_ref(src_root="/root/reference", dst=REF_DST):
    """copies the files if the reference checkout is present; returns the destination or None"""
    if not os.path.isfile(os.path.join(src_root, FILES[1])):
        return dst if ref_available(dst) else None
    manifest = {}
    for rel in FILES:
        a, b = os.path.join(src_root, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(b), exist_ok=True)
        if os.path.exists(b):
            os.chmod(b, 0o644)
        shutil.copyfile(a, b)
        with open(b, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": "InsectRobotics/InvertSy (unmodified files)", "sha256": manifest}, f, indent=1)
    return dst


def ref_available(dst=REF_DST):
    return all(os.path.isfile(os.path.join(dst, rel)) for rel in FILES)


if __name__ == "__main__":
    print(build_ref())
