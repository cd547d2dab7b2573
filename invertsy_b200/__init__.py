"""
invertsy_b200 -- B200-native (sm_100a) renderer for InvertSy's compound-eye views.

A drop-in for ONE hot path of InsectRobotics/InvertSy: ``invertsy.env.world`` (what a set of
viewing directions sees of a triangle world) fused with ``invertsy.env.sky`` (Perez luminance,
degree and angle of polarisation).  The compute lives in hand-written CUDA behind the C ABI of
``include/isy_b200.h``; this package is the Python mirror of the reference interface plus a
batched API.  There is no CPU fallback.
"""
from . import env  # noqa: F401
from .env.world import WorldBase, Seville2009, SimpleWorld  # noqa: F401
from .env.sky import Sky, UniformSky  # noqa: F401
from .render import EyeLayout, Renderer, fibonacci_lattice  # noqa: F401
from .sense import CompoundEye, PolarisationSensor  # noqa: F401
from . import datasets, distributed, familiarity, views  # noqa: F401

__version__ = "0.1.0"
