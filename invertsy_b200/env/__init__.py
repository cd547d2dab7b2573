"""Mirror of ``invertsy.env`` for the rendering hot path (reference: src/invertsy/env/__init__.py:18-21)."""
from .world import WorldBase, Seville2009, SimpleWorld  # noqa: F401
from .sky import Sky, UniformSky  # noqa: F401
from .occlusion import add_noise, RNG, eps  # noqa: F401
