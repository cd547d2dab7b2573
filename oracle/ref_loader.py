"""
TEST INFRASTRUCTURE ONLY -- loader for the *unmodified* reference modules.

Imports ``invertsy.env.world`` and ``invertsy.env.sky`` from a read-only checkout of
InsectRobotics/InvertSy (``/root/reference`` in the build container) without pulling in
InvertPy / matplotlib / pytz, which are not installed.  Used only by
``oracle/gen_golden.py`` and by the ``reference``-marked CPU tests to (a) pin the oracle
restatement in ``oracle/isy_oracle.py`` and (b) generate the fixtures in ``tests/golden``.

The reference tree does not exist on the GPU box: nothing in ``-m gpu`` tests, ``smoke()``
or ``bench.py`` calls this module.

Shims:
  1. ``matplotlib.cm.get_cmap`` stub          (needed by world.py:23,29 at import)
  2. ``pytz.timezone`` stub                   (needed by observer.py:16,50, ephemeris.py:17)
  3. bare ``invertsy`` / ``invertsy.env`` package objects so that invertsy/__init__.py:14-16
     (which imports InvertPy) never runs.
"""
import importlib
import os
import sys
import types
import warnings

DEFAULT_REFERENCE_ROOT = os.environ.get("ISY_REFERENCE_ROOT", "/root/reference")


def reference_available(root=DEFAULT_REFERENCE_ROOT):
    return os.path.isfile(os.path.join(root, "src", "invertsy", "env", "world.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference(root=DEFAULT_REFERENCE_ROOT):
    """Returns (world_module, sky_module) of the unmodified reference."""
    if not reference_available(root):
        raise FileNotFoundError("reference checkout not found under %r" % root)
    src = os.path.join(root, "src")

    try:
        import matplotlib  # noqa: F401
    except ImportError:
        cm = _stub("matplotlib.cm", get_cmap=lambda name=None: None)
        _stub("matplotlib", cm=cm)
    try:
        import pytz  # noqa: F401
    except ImportError:
        _stub("pytz", timezone=lambda name: None, utc=None)

    if "invertsy.env.world" not in sys.modules:
        pkg = types.ModuleType("invertsy")
        pkg.__path__ = [os.path.join(src, "invertsy")]
        sys.modules["invertsy"] = pkg
        env = types.ModuleType("invertsy.env")
        env.__path__ = [os.path.join(src, "invertsy", "env")]
        sys.modules["invertsy.env"] = env
        pkg.env = env

    warnings.filterwarnings("ignore", category=DeprecationWarning)
    world = importlib.import_module("invertsy.env.world")
    sky = importlib.import_module("invertsy.env.sky")
    return world, sky


def reference_file_hashes(root=DEFAULT_REFERENCE_ROOT):
    import hashlib
    out = {}
    for rel in ("src/invertsy/env/world.py", "src/invertsy/env/sky.py", "src/invertsy/env/_helpers.py",
                "src/invertsy/__helpers.py", "data/Seville2009_world/world5000_gray.mat",
                "data/Seville2009_world/AntRoutes.mat", "data/Sparse_world/world_gray.mat"):
        with open(os.path.join(root, rel), "rb") as f:
            out[rel] = hashlib.sha256(f.read()).hexdigest()
    return out
