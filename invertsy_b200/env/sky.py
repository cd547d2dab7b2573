"""
GPU drop-in for ``invertsy.env.sky`` (reference src/invertsy/env/sky.py).

``UniformSky`` / ``Sky`` keep the reference's constructor, call signature
``sky(ori, irgbu, noise, eta, rng) -> (Y, DoP, AoP)``, the cached ``Y / DOP / AOP / theta / phi /
eta`` properties with their "not generated yet" assertions, and the turbidity <-> Perez
coefficient plumbing.  The per-direction maths runs on the GPU (``isy_sky_call_host``,
``isy_spectrum_influence_host`` in include/isy_b200.h); scalars stay on the host.

Out of scope: ``Sky.from_type`` (unreachable in the reference: it opens a
path that does not exist, sky.py:646) and the matplotlib ``visualise_*`` helpers.
"""
import ctypes

import numpy as np
from scipy.spatial.transform import Rotation as R

from .. import _lib
from .occlusion import RNG, add_noise, eps

# turbidity -> Perez coefficients (sky.py:25-29)
T_L = np.array([[0.1787, -1.4630],
                [-0.3554, 0.4275],
                [-0.0227, 5.3251],
                [0.1206, -2.5771],
                [-0.0670, 0.3703]])


def spectrum_influence(v, irgbu, device=0):
    """
    Luminance split into the (IR, R, G, B, UV) channels for the given sensitivities
    (sky.py:671-703); returns (n, 5).  Its nansum / nanmax reductions run over the whole call.
    """
    v = _lib.as_f64(v).reshape(-1)
    irgbu = np.atleast_2d(_lib.as_f64(irgbu))
    out = np.empty((v.size, 5), dtype=np.float64)
    _lib.check(_lib.lib().isy_spectrum_influence_host(_lib.ptr(v), v.size, _lib.ptr(irgbu), irgbu.shape[0],
                                                      int(device), None, _lib.ptr(out)))
    return out


def _spectrum_sum(v, irgbu, device):
    out = np.empty_like(v)
    _lib.check(_lib.lib().isy_spectrum_influence_host(_lib.ptr(v), v.size, _lib.ptr(irgbu), irgbu.shape[0],
                                                      int(device), _lib.ptr(out), None))
    return out


class UniformSky(object):
    """Constant-luminance, unpolarised sky (reference ``UniformSky``, sky.py:32-241)."""

    def __init__(self, luminance=1., name="uniform-sky", device=0):
        self.__luminance = luminance
        self.__y = np.full(1, np.nan)
        self.__aop = np.full(1, np.nan)
        self.__dop = np.full(1, np.nan)
        self.__eta = np.full(1, False)
        self.__theta = np.full(1, np.nan)
        self.__phi = np.full(1, np.nan)
        self._is_generated = False
        self.__name = name
        self.device = device
        self._last_quat = None

    # ---- GPU plumbing shared with Sky ----
    def _params(self):
        sp = _lib.SkyParams()
        sp.kind = _lib.SKY_UNIFORM
        sp.luminance = float(self.__luminance)
        return sp

    def _sun(self):
        return 0., 0.

    def _device_call(self, quat, eta_mask, want_coordinates=True):
        n = quat.shape[0]
        y, p, a = (np.empty(n, dtype=np.float64) for _ in range(3))
        th = np.empty(n, dtype=np.float64) if want_coordinates else None
        ph = np.empty(n, dtype=np.float64) if want_coordinates else None
        sp = self._params()
        ths, phs = self._sun()
        _lib.check(_lib.lib().isy_sky_call_host(ctypes.byref(sp), float(ths), float(phs), _lib.ptr(quat), n,
                                                _lib.ptr(eta_mask), int(self.device), _lib.ptr(y), _lib.ptr(p),
                                                _lib.ptr(a), _lib.ptr(th), _lib.ptr(ph)))
        return y, p, a, th, ph

    def _render(self, ori, irgbu, noise, eta, rng):
        if ori is not None:
            quat = _lib.as_f64(ori.as_quat()).reshape(-1, 4)
            n = quat.shape[0]
        else:
            # sky.py:110-111 re-uses the cached coordinates: theta/phi stay, the orientation is rebuilt
            # from them with theta as the Y rotation (which does not round-trip; replicated as written)
            n = self.__phi.shape[0]
            quat = None
        if eta is None:
            eta = add_noise(noise=noise, shape=(n,), rng=rng)
        mask = np.zeros(n, dtype=bool)
        mask[eta] = True
        mask8 = np.ascontiguousarray(mask, dtype=np.uint8) if mask.any() else None

        if quat is not None:
            y, p, a, th, ph = self._device_call(quat, mask8)
            self.theta = th
            self.phi = ph
            self._last_quat = quat
        else:
            if self._last_quat is None:
                raise ValueError("sky called with ori=None before any orientations were given")
            y, p, _, _, _ = self._device_call(self._last_quat, mask8, want_coordinates=False)
            rebuilt = R.from_euler('ZY', np.vstack([self.__phi, self.__theta]).T, degrees=False)
            _, _, a, _, _ = self._device_call(_lib.as_f64(rebuilt.as_quat()).reshape(-1, 4), mask8,
                                              want_coordinates=False)
        self.__y = y
        self.__dop = p
        self.__aop = a
        self.__eta = eta
        self._is_generated = True
        if irgbu is not None:
            irgbu = np.atleast_2d(_lib.as_f64(irgbu))
            y = _spectrum_sum(np.ascontiguousarray(y), irgbu, self.device)
        return y, p, a

    def __call__(self, ori=None, irgbu=None, noise=0., eta=None, rng=RNG):
        """
        Skylight for the given orientations (sky.py:47-97): Y = luminance, DoP = 0, AoP = NaN,
        occluded rays (eta) get Y = 0; with ``irgbu`` Y is the spectrally weighted sum.
        """
        return self._render(ori, irgbu, noise, eta, rng)

    # ---- cached results (sky.py:115-200) ----
    @property
    def Y(self):
        assert self._is_generated, "Sky is not generated yet. In order to generate the env, use the call function."
        return self.__y

    @property
    def DOP(self):
        assert self._is_generated, "Sky is not generated yet. In order to generate the env, use the call function."
        return self.__dop

    @property
    def AOP(self):
        assert self._is_generated, "Sky is not generated yet. In order to generate the env, use the call function."
        return self.__aop

    @property
    def theta(self):
        assert self._is_generated, "Sky is not generated yet. In order to generate the env, use the call function."
        return self.__theta

    @theta.setter
    def theta(self, value):
        self.__theta = value
        self._is_generated = False

    @property
    def phi(self):
        assert self._is_generated, "Sky is not generated yet. In order to generate the env, use the call function."
        return self.__phi

    @phi.setter
    def phi(self, value):
        self.__phi = value
        self._is_generated = False

    @property
    def eta(self):
        assert self._is_generated, "Sky is not generated yet. In order to generate the env, use the call function."
        return self.__eta

    @eta.setter
    def eta(self, value):
        self.__eta = value
        self._is_generated = False

    @property
    def luminance(self):
        return self.__luminance

    @property
    def name(self):
        return self.__name


class Sky(UniformSky):
    """Analytic clear sky: Perez luminance + single-scattering polarisation (sky.py:243-668)."""

    def __init__(self, theta_s=0., phi_s=0., degrees=False, name="sky", device=0):
        """
        theta_s is used as the sun's *zenith distance* in the maths (sky.py:304, :609) although
        the reference docstring calls it elevation; phi_s is the sun azimuth.
        """
        super(Sky, self).__init__(name=name, device=device)
        self.__a, self.__b, self.__c, self.__d, self.__e = 0., 0., 0., 0., 0.
        self.__tau_L = 2.
        self._update_luminance_coefficients(self.__tau_L)
        self.__c1 = .6
        self.__c2 = 4.
        self.__theta_s = np.deg2rad(theta_s) if degrees else theta_s
        self.__phi_s = np.deg2rad(phi_s) if degrees else phi_s

    def _params(self):
        sp = _lib.SkyParams()
        sp.kind = _lib.SKY_PEREZ
        sp.luminance = 0.
        sp.A, sp.B, sp.C, sp.D, sp.E = (float(x) for x in (self.A, self.B, self.C, self.D, self.E))
        sp.tau_L = float(self.tau_L)
        sp.c1, sp.c2 = float(self.c1), float(self.c2)
        return sp

    def _sun(self):
        return self.theta_s, self.phi_s

    def __call__(self, ori=None, irgbu=None, noise=0., eta=None, rng=RNG):
        """Y (Perez luminance, 17 inside the sun disc), DoP and AoP per orientation (sky.py:269-346)."""
        return self._render(ori, irgbu, noise, eta, rng)

    # ---- sun position ----
    @property
    def theta_s(self):
        return self.__theta_s

    @theta_s.setter
    def theta_s(self, value):
        self.__theta_s = value
        self._is_generated = False

    @property
    def phi_s(self):
        return self.__phi_s

    @phi_s.setter
    def phi_s(self, value):
        self.__phi_s = value
        self._is_generated = False

    # ---- Perez coefficients <-> turbidity (sky.py:376-554) ----
    def _set_coeff(self, which, value):
        setattr(self, "_Sky__" + which, value)
        self._update_turbidity(self.A, self.B, self.C, self.D, self.E)
        self._is_generated = False

    A = property(lambda self: self.__a, lambda self, v: self._set_coeff("a", v),
                 doc="darkening or brightening of the horizon")
    B = property(lambda self: self.__b, lambda self, v: self._set_coeff("b", v),
                 doc="luminance gradient near the horizon")
    C = property(lambda self: self.__c, lambda self, v: self._set_coeff("c", v),
                 doc="relative intensity of the circumsolar region")
    D = property(lambda self: self.__d, lambda self, v: self._set_coeff("d", v),
                 doc="width of the circumsolar region")
    E = property(lambda self: self.__e, lambda self, v: self._set_coeff("e", v),
                 doc="relative backscattered light")

    @property
    def c1(self):
        return self.__c1

    @property
    def c2(self):
        return self.__c2

    @property
    def tau_L(self):
        return self.__tau_L

    @tau_L.setter
    def tau_L(self, value):
        assert value >= 1., "Turbidity must be greater or eaqual to 1."
        self._is_generated = self.__tau_L == value and self._is_generated
        self._update_luminance_coefficients(value)

    @property
    def Y_z(self):
        """Zenith luminance in K cd/m^2 (sky.py:501-510)."""
        chi = (4. / 9. - self.tau_L / 120.) * (np.pi - 2 * (np.pi / 2 - self.theta_s))
        return (4.0453 * self.tau_L - 4.9710) * np.tan(chi) - 0.2155 * self.tau_L + 2.4192

    @property
    def M_p(self):
        """Maximum degree of polarisation (sky.py:513-521)."""
        return np.exp(-(self.tau_L - self.c1) / (self.c2 + eps))

    def _update_luminance_coefficients(self, tau_L):
        self.__a, self.__b, self.__c, self.__d, self.__e = T_L.dot(np.array([tau_L, 1.]))
        self._update_turbidity(self.A, self.B, self.C, self.D, self.E)

    def _update_turbidity(self, a, b, c, d, e):
        tau_L, k = np.linalg.pinv(T_L).dot(np.array([a, b, c, d, e]))
        self.__tau_L = tau_L / k

    def copy(self):
        """Copy with the same sun position and turbidity (sky.py:556-578)."""
        sky = Sky(device=self.device)
        sky.tau_L = self.tau_L
        sky.theta_s = self.theta_s
        sky.phi_s = self.phi_s
        sky._is_generated = False
        return sky

    @staticmethod
    def from_observer(obs=None, date=None, ori=None):
        """
        Sky for an observer on Earth (sky.py:580-611).  The sun ephemeris is a scalar host-side
        set-up outside the rendering hot path.
        """
        from ..ephemeris import sun_position_for_observer
        alt, az = sun_position_for_observer(obs, date)
        if ori is not None:
            yaw, pitch, roll = ori.as_euler('ZYX', degrees=False)
        else:
            yaw = 0.
        theta_s, phi_s = np.pi / 2 - alt, (az - yaw + np.pi) % (2 * np.pi) - np.pi
        return Sky(theta_s=theta_s, phi_s=phi_s)
