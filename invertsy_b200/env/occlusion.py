"""
Host-side occlusion ("eta") mask of the reference's environment callables, plus the shared
constants of the package.

``add_noise`` keeps the call signature and the observable behaviour of the reference's
``invertsy.env._helpers.add_noise`` (reference src/invertsy/env/_helpers.py:10-36):

  * ``noise`` is an ndarray  -> it *is* the mask (cast to bool); if its size differs from
    ``sum(shape)`` (sic) it is written into the head of an all-False mask of ``shape``;
  * ``noise`` > 0            -> the mask is an *index* array: the positions of the
    ``int(noise * shape[0])`` smallest ``|randn(*shape)|`` along the last axis (argsort);
  * otherwise                -> ``zeros_like(v, bool)``, i.e. a 0-d ``False`` when ``v`` is None,
    which makes ``c[eta] = 0`` a no-op.

The default generator and ``eps`` follow reference src/invertsy/__helpers.py:6-10.
"""
import os

import numpy as np

RNG = np.random.RandomState(2021)
eps = np.finfo(float).eps
__data__ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


def _mask_from_array(noise, shape):
    expected = None if shape is None else np.sum(shape)
    if expected is None or noise.size == expected:
        return np.array(noise, dtype=bool)
    mask = np.zeros(shape, dtype=bool)
    mask[:noise.size] = noise
    return mask


def add_noise(v=None, noise=0., shape=None, fill_value=0, rng=RNG):
    if v is not None and shape is None:
        shape = v.shape
    if isinstance(noise, np.ndarray):
        eta = _mask_from_array(noise, shape)
    elif noise > 0 and shape is not None:
        ranked = np.argsort(np.absolute(rng.randn(*shape)))
        eta = ranked[:int(noise * shape[0])]
    elif noise > 0:
        eta = rng.randn()
    else:
        eta = np.zeros_like(v, dtype=bool)
    if v is not None:
        v[eta] = fill_value
    return eta
