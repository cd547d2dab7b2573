"""
TEST INFRASTRUCTURE ONLY -- CPU (NumPy) restatement of the InvertSy eye-rendering hot path.

This module is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path (``invertsy_b200``) never imports anything from ``oracle/`` and has no CPU
fallback.

Parity status: **pinned**.  ``tests/test_oracle_vs_reference.py`` (runs wherever the
reference checkout is present) and the committed fixtures under ``tests/golden/`` (made
by ``oracle/gen_golden.py`` from the unmodified reference) check every function here
bit-for-bit against ``/root/reference/src/invertsy/env/{world,sky}.py``.

Each function cites the reference lines it follows.  The arithmetic keeps the reference's
operation order and dtypes (NumPy float64 unless the reference itself drops to float32),
and deliberately keeps the reference's *structure* where it determines run time (the
per-triangle Python loop with six 2-vector ``np.cross`` calls), so that timing this port
is representative of timing the reference (measured here: within a few % of the
reference, see DESIGN.md "CPU baseline").
"""
import warnings

import numpy as np
from scipy.spatial.transform import Rotation as R

eps = np.finfo(float).eps                       # invertsy/__helpers.py:10
GROUND_COLOUR = (229, 183, 90)                  # env/world.py:36
SKY_COLOUR = (13, 135, 201)                     # env/world.py:37
WAVELENGTHS = (1200, 715, 535, 475, 350)        # env/sky.py:686

# turbidity -> Perez coefficients, env/sky.py:25-29
T_L = np.array([[0.1787, -1.4630],
                [-0.3554, 0.4275],
                [-0.0227, 5.3251],
                [0.1206, -2.5771],
                [-0.0670, 0.3703]])


# ----------------------------------------------------------------------------------------
# env/_helpers.py:10-36  (occlusion / "eta" mask)
# ----------------------------------------------------------------------------------------
def add_noise(v=None, noise=0., shape=None, fill_value=0, rng=None):
    """Restates ``add_noise`` (env/_helpers.py:10-36), including its odd ``np.sum(shape)`` size rule."""
    if shape is None and v is not None:
        shape = v.shape
    if shape is not None:
        size = np.sum(shape)
    elif v is not None:
        size = v.size
    else:
        size = None
    if isinstance(noise, np.ndarray):
        if size is None or noise.size == size:
            eta = np.array(noise, dtype=bool)
        else:
            eta = np.zeros(shape, dtype=bool)
            eta[:noise.size] = noise
    elif noise > 0:
        if shape is not None:
            eta = np.argsort(np.absolute(rng.randn(*shape)))[:int(noise * shape[0])]
        else:
            eta = rng.randn()
    else:
        eta = np.zeros_like(v, dtype=bool)      # v is None -> 0-d array(False): a no-op mask
    if v is not None:
        v[eta] = fill_value
    return eta


# ----------------------------------------------------------------------------------------
# env/world.py
# ----------------------------------------------------------------------------------------
def luminance2skycolour(y):
    """env/world.py:465-479 -- blue shade on the 0..255 scale."""
    return np.clip(19. * y[..., np.newaxis] + np.array(SKY_COLOUR).reshape((1, -1)), 0, 255)


def _cross2(u, v):
    # the reference calls np.cross on 2-vectors (env/world.py:526), which NumPy evaluates as
    # u0*v1 - u1*v0 with separately rounded products; np.cross is kept for timing fidelity.
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)
        return np.cross(u, v)


def same_side(p1, p2, a, b):
    """env/world.py:506-526."""
    return np.multiply(_cross2(b - a, p1 - a), _cross2(b - a, p2 - a)) >= 0


def same_side_technique(points, triangle):
    """env/world.py:482-503 -- 2-D point-in-triangle in the (pitch, yaw) plane."""
    a, b, c = triangle[..., 0, :], triangle[..., 1, :], triangle[..., 2, :]
    return same_side(points, a, b, c) & same_side(points, b, a, c) & same_side(points, c, a, b)


def world_render(polygons, colours, horizon, pos, ori, init_c=None, noise=0., eta=None, rng=None,
                 dtype='float32'):
    """
    Restates ``WorldBase.__call__`` (env/world.py:55-129; duplicated at :344-418).

    polygons (T,3,3) float32, colours (T,3) float32, ``ori`` a SciPy Rotation of length N.
    Returns (N,3): float32 if ``init_c`` is None else float64.
    """
    if init_c is None:                                                   # :79-82
        c = np.full((np.shape(ori)[0], 3), np.nan, dtype=dtype)
    else:
        c = luminance2skycolour(init_c)

    yaw, pitch, roll = ori.as_euler('ZYX', degrees=False).T            # :84
    c[pitch >= 0, :] = np.array(GROUND_COLOUR) / 255.                   # :90

    xyz = polygons - np.nanmean(pos, axis=0)                            # :94
    dist = np.min(np.linalg.norm(xyz, axis=2), axis=1)                  # :97
    visible = dist <= horizon                                           # :98
    tris = xyz[visible]
    keep = ~np.isnan(np.linalg.norm(tris, axis=(1, 2)))                 # :102-104
    cols = colours[visible][keep]
    tris = tris[keep]
    order = np.argsort(dist[visible])[::-1]                             # :106 far -> near

    poi = np.array([pitch, yaw]).T                                      # :112 (loop invariant)
    for tri, col in zip(tris[order], cols[order]):                      # :108
        phi = np.arctan2(tri[..., 1], tri[..., 0])                      # :109
        theta = np.arctan2(np.linalg.norm(tri[..., :2], axis=1), tri[..., 2]) - np.pi / 2  # :110
        pol = np.array([theta, phi]).T                                  # :113
        if phi.max() - phi.min() >= np.pi:                              # :116 seam rule
            p1, p2 = pol.copy(), pol.copy()
            p1[:, 1][phi < 0] += 2 * np.pi
            p2[:, 1][phi > 0] -= 2 * np.pi
            inside = same_side_technique(poi, p1) | same_side_technique(poi, p2)
        else:
            inside = same_side_technique(poi, pol)
        c[inside] = col                                                 # :123

    if eta is None:                                                     # :125-127
        eta = add_noise(noise=noise, shape=c.shape, rng=rng)
    c[eta] = 0.
    return c


def world_hits(polygons, horizon, pos, pitch, yaw):
    """
    Index oracle: for each query (pitch, yaw) the visible triangle that the painter's loop
    (env/world.py:106-123) leaves on top = the containing triangle with the smallest
    min-vertex distance.  Returns (hit int32 (N,), -1 = none; depth float64 (N,), nan = none;
    n_visible).  Same arithmetic as ``world_render`` but vectorised over rays per triangle
    without np.cross, and it records which triangle won instead of its colour.
    Equivalence with the reference via index-encoded colours is checked in
    tests/test_oracle_vs_reference.py.
    """
    pitch = np.asarray(pitch, dtype=np.float64)
    yaw = np.asarray(yaw, dtype=np.float64)
    xyz = polygons - np.nanmean(np.atleast_2d(pos), axis=0)
    dist = np.min(np.linalg.norm(xyz, axis=2), axis=1)
    vis = np.flatnonzero(dist <= horizon)
    order = vis[np.argsort(dist[vis])[::-1]]
    hit = np.full(pitch.shape, -1, dtype=np.int32)
    for t in order:
        tri = xyz[t]
        phi = np.arctan2(tri[:, 1], tri[:, 0])
        theta = np.arctan2(np.linalg.norm(tri[:, :2], axis=1), tri[:, 2]) - np.pi / 2
        if phi.max() - phi.min() >= np.pi:
            php = np.where(phi < 0, phi + 2 * np.pi, phi)
            phm = np.where(phi > 0, phi - 2 * np.pi, phi)
            inside = _inside_fast(pitch, yaw, theta, php) | _inside_fast(pitch, yaw, theta, phm)
        else:
            inside = _inside_fast(pitch, yaw, theta, phi)
        hit[inside] = t
    depth = np.where(hit >= 0, dist[np.maximum(hit, 0)].astype(np.float64), np.nan)
    return hit, depth, int(vis.size)


def _inside_fast(p, y, th, ph):
    def ss(ax, ay, bx, by, qx, qy):
        ux, uy = bx - ax, by - ay
        c1 = ux * (y - ay) - uy * (p - ax)
        c2 = ux * (qy - ay) - uy * (qx - ax)
        return c1 * c2 >= 0
    a, b, c = (th[0], ph[0]), (th[1], ph[1]), (th[2], ph[2])
    return (ss(b[0], b[1], c[0], c[1], a[0], a[1]) &
            ss(a[0], a[1], c[0], c[1], b[0], b[1]) &
            ss(a[0], a[1], b[0], b[1], c[0], c[1]))


# ----------------------------------------------------------------------------------------
# env/sky.py
# ----------------------------------------------------------------------------------------
def sky_coefficients(tau_L=2.):
    """env/sky.py:523-533 (A..E = T_L . [tau, 1])."""
    return tuple(T_L.dot(np.array([tau_L, 1.])))


def sky_tau_from_coefficients(a, b, c, d, e):
    """env/sky.py:535-554 (pseudo-inverse round trip that Sky performs after every coefficient update)."""
    T_T = np.linalg.pinv(T_L)
    tau_L, k = T_T.dot(np.array([a, b, c, d, e]))
    return tau_L / k


def sky_zenith_luminance(tau_L, theta_s):
    """env/sky.py:501-510."""
    chi = (4. / 9. - tau_L / 120.) * (np.pi - 2 * (np.pi / 2 - theta_s))
    return (4.0453 * tau_L - 4.9710) * np.tan(chi) - 0.2155 * tau_L + 2.4192


def sky_max_dop(tau_L, c1=.6, c2=4.):
    """env/sky.py:513-521."""
    return np.exp(-(tau_L - c1) / (c2 + eps))


def perez_L(chi, z, A, B, C, D, E):
    """env/sky.py:348-374."""
    z = np.array(z)
    i = z < (np.pi / 2)
    f = np.zeros_like(z)
    if z.ndim > 0:
        f[i] = (1. + A * np.exp(B / (np.cos(z[i]) + eps)))
    elif i:
        f = (1. + A * np.exp(B / (np.cos(z) + eps)))
    phi = (1. + C * np.exp(D * chi) + E * np.square(np.cos(chi)))
    return f * phi


def sky_coordinates(ori):
    """env/sky.py:99-113 -- (theta = zenith angle, phi = azimuth) of ori.apply([1,0,0])."""
    xyz = np.clip(ori.apply([1, 0, 0]), -1, 1)
    phi = np.arctan2(xyz[..., 1], xyz[..., 0])
    theta = np.arccos(xyz[..., 2])
    phi = (phi + np.pi) % (2 * np.pi) - np.pi
    return theta, phi


def spectrum_influence(v, irgbu):
    """env/sky.py:671-703."""
    wl = np.array(WAVELENGTHS, dtype='float32')
    v = v[..., np.newaxis]
    irgbu = np.vstack([irgbu] + [irgbu for _ in range(v.shape[0] // irgbu.shape[0] - 1)])
    l1 = 10.0 * irgbu * np.power(wl / 1000., 8) * np.square(v) / float(v.size)
    l2 = 0.001 * irgbu * np.power(1000. / wl, 8) * np.nansum(np.square(v)) / float(v.size)
    v_max = np.nanmax(v)
    w_max = np.nanmax(v + l1 + l2)
    w = v_max * (v + l1 + l2) / w_max
    if isinstance(irgbu, np.ndarray):
        if irgbu.shape[0] == 1 and w.shape[0] > irgbu.shape[0]:
            irgbu = np.vstack([irgbu] * w.shape[0])
        w[irgbu < 0] = np.hstack([v] * irgbu.shape[1])[irgbu < 0]
    elif irgbu < 0:
        w = v
    return w


def sky_render(theta_s, phi_s, ori, tau_L=2., c1=.6, c2=4., irgbu=None, noise=0., eta=None, rng=None,
               return_raw=False):
    """
    Restates ``Sky.__call__`` (env/sky.py:269-346).  Returns (Y, DoP, AoP), each (N,) float64.
    ``tau_L`` goes through the same T_L / pinv round trip the reference's constructor does
    (env/sky.py:259-261 -> :523-554), so the coefficients are bit-identical.
    """
    A, B, C, D, E = sky_coefficients(tau_L)
    tau_L = sky_tau_from_coefficients(A, B, C, D, E)
    theta, phi = sky_coordinates(ori)

    gamma = np.arccos(np.cos(theta) * np.cos(theta_s) +
                      np.sin(theta) * np.sin(theta_s) * np.cos(phi - phi_s))        # :304-305
    i_prez = perez_L(gamma, theta, A, B, C, D, E)                                     # :308
    i_00 = perez_L(0., theta_s, A, B, C, D, E)                                        # :309
    i_90 = perez_L(np.pi / 2, np.absolute(theta_s - np.pi / 2), A, B, C, D, E)        # :310
    i = (1. / (i_prez + eps) - 1. / (i_00 + eps)) * i_00 * i_90 / (i_00 - i_90 + eps)  # :312
    y = np.maximum(sky_zenith_luminance(tau_L, theta_s) * i_prez / (i_00 + eps), 0.)  # :313

    lp = np.square(np.sin(gamma)) / (1 + np.square(np.cos(gamma)))                    # :316
    p = np.clip(2. / np.pi * sky_max_dop(tau_L, c1, c2) * lp *
                (theta * np.cos(theta) + (np.pi / 2 - theta) * i), 0., 1.)            # :317

    ori_s = R.from_euler('ZY', [phi_s, np.pi / 2 - theta_s], degrees=False)           # :320
    x_s, y_s, _ = ori_s.apply([1, 0, 0]).T
    x_p, y_p, _ = ori.apply([1, 0, 0]).T
    a_x = np.arctan2(y_p - y_s, x_p - x_s) + np.pi / 2                                # :323
    a = (a_x + np.pi) % (2 * np.pi) - np.pi                                           # :324

    if eta is None:                                                                   # :327-328
        eta = add_noise(noise=noise, shape=y.shape, rng=rng)
    y[eta] = 0.
    p[eta] = 0.
    a[eta] = np.nan
    y[gamma < np.pi / 60] = 17                                                        # :334 sun disc

    if return_raw:
        return y, p, a, gamma, theta, phi
    if irgbu is not None:                                                             # :343-344
        y = spectrum_influence(y, irgbu).sum(axis=1)
    return y, p, a


def uniform_sky_render(luminance, ori, irgbu=None, noise=0., eta=None, rng=None):
    """Restates ``UniformSky.__call__`` (env/sky.py:47-97)."""
    theta, phi = sky_coordinates(ori)
    y = np.full_like(phi, luminance)
    p = np.full_like(phi, 0.)
    a = np.full_like(phi, np.nan)
    if eta is None:
        eta = add_noise(noise=noise, shape=y.shape, rng=rng)
    y[eta] = 0.
    if irgbu is not None:
        y = spectrum_influence(y, irgbu).sum(axis=1)
    return y, p, a


# ----------------------------------------------------------------------------------------
# Eye lattices and view composition used by goldens, tests and bench
# ----------------------------------------------------------------------------------------
def fibonacci_eye(n, hemisphere=False):
    """
    Deterministic ommatidia lattice: phi_i = (pi(1+sqrt5)(i+1/2)) mod 2pi - pi,
    z_i = 1 - 2(i+1/2)/n (full sphere) or 1 - (i+1/2)/n (upper hemisphere).
    Returns (pitch, yaw) with pitch = arccos(z) - pi/2 (negative = looking up, Appendix C).
    """
    i = np.arange(n, dtype=np.float64)
    yaw = (np.pi * (1. + np.sqrt(5.)) * (i + .5)) % (2 * np.pi) - np.pi
    z = 1. - (1. if hemisphere else 2.) * (i + .5) / n
    pitch = np.arccos(z) - np.pi / 2
    return pitch, yaw


def eye_rotations(pitch, yaw):
    """Ommatidia orientations in the eye frame: direction (cos p cos y, cos p sin y, -sin p)."""
    return R.from_euler('ZY', np.c_[yaw, pitch], degrees=False)


def view_rotations(heading_deg, omm_ori):
    """Global ommatidia orientations for a yaw-only heading (sim/simulation.py:836, 2706)."""
    return R.from_euler('Z', heading_deg, degrees=True) * omm_ori
