"""
CompoundEye-compatible sensor on top of the GPU renderer.

The reference takes its eye from the third-party package InvertPy (`from invertpy.sense import
CompoundEye`, reference agent/agent.py:20, sim/simulation.py:24), which is neither vendored nor pinned;
its acceptance-cone sampler and photoreceptor transform are therefore **unpinned** (SURVEY.md 0.3, 8c).
This class keeps the *surface* the reference uses -- constructor keywords (agent.py:588,
simulation.py:796, 2238, 2599, examples/create_familiarity_map_data.py:30), `eye(sky=, scene=,
callback=)` returning `(nb_ommatidia, channels)` responses whose consumers take `.mean(axis=1)` and clip
to [0, 1] (agent.py:744, 788; simulation.py:816, 838), and the attributes read elsewhere (`omm_ori`,
`nb_ommatidia`, `ori`, `xyz`, `x/y/z`, `yaw_deg`, `responses`, `_xyz`, `_ori`) -- and defines the
unpinned parts explicitly:

  * acceptance cone: `nb_samples` (1, 2, 4, 8, 16, 32) directions per ommatidium: the axis, then rings
    of points at 0.5 * omm_rho and omm_rho/... (see `cone_offsets`), Gaussian weights with
    sigma = omm_rho / 2, normalised per ommatidium;
  * photoreceptor transform (`isy_sensor_params`, include/isy_b200.h -- evaluated ON THE DEVICE by the final
    pass of the render kernel, so only responses leave the GPU): per cone sample the colour on the 0..1 scale
    is the hit triangle's colour, else the ground colour where the sample looks down, else
    `luminance2skycolour(Y) / 255` under a sky, else 0; the samples are averaged with the cone weights;
    responses (N, 3) = clip(mean colour * c_sensitive[R, G, B] * omm_res, 0, 1).  `transform_reference` states
    the same in NumPy (float64) for tests and documentation;
  * `omm_pol_op` = rho > 0: a sample that sees the sky is multiplied by rho (sin^2(A - phi) + cos^2(A - phi)(1 - P)^2)
    + (1 - rho), P / A the sky's degree / angle of polarisation in its direction and phi the azimuth of the sample's
    viewing direction (the microvilli of the photoreceptor lie along it); `PolarisationSensor` (the dorsal-rim sensor
    of the path-integration agents, agent.py:835, :879) returns one such photoreceptor per ommatidium:
    omm_res * sqrt(cone average of Y * that term), the sky's luminance seen through a polarisation filter;
  * `eye(sky=sky)` without a scene (examples/test_sky.py:22) renders against an empty world: the sky alone;
  * `noise` > 0: a fraction of the ommatidia is blanked per call exactly as the reference's environment does for
    its own outputs (`env/_helpers.py:10-36` add_noise -> `c[eta] = 0`, world.py:125-127), with the eye's `rng`.

Everything the eye computes per ray goes through `Renderer` (one fused launch per call or per batch); a
yaw-only eye orientation -- every pose the reference's simulations set (`R.from_euler('Z', yaw)`,
simulation.py:836, 2706) -- takes the rasteriser for yaw-only views, any other orientation the general one.
"""
import numpy as np
from scipy.spatial.transform import Rotation as R

from . import _lib
from .env.occlusion import RNG, add_noise
from .render import EyeLayout, Renderer, fibonacci_lattice


def cone_offsets(nb_samples, rho):
    """
    Deterministic sample pattern of an acceptance cone of full angle `rho` around +x: returns
    (pitch_off, yaw_off, weight), sample 0 on the axis, the others on a ring of radius rho/2 (and a
    second ring at rho/4 from 16 samples up), Gaussian weights with sigma = rho/2.
    """
    if nb_samples == 1:
        return np.zeros(1), np.zeros(1), np.ones(1)
    k = np.arange(1, nb_samples)
    rad = np.full(k.size, rho / 2.)
    if nb_samples >= 16:
        rad[k % 2 == 0] = rho / 4.
    ang = 2 * np.pi * k / (nb_samples - 1) + (0.5 if nb_samples >= 16 else 0.) * (k % 2)
    d_pitch = np.concatenate([[0.], rad * np.sin(ang)])
    d_yaw = np.concatenate([[0.], rad * np.cos(ang)])
    r2 = d_pitch ** 2 + d_yaw ** 2
    w = np.exp(-r2 / (2 * (rho / 2.) ** 2)) if rho > 0 else np.ones(nb_samples)
    return d_pitch, d_yaw, w / w.sum()


def transform_reference(sample_rgb01, weights, w_rgb, gain):
    """
    NumPy (float64) statement of the photoreceptor transform the device evaluates: sample_rgb01 (..., N, S, 3)
    colours on the 0..1 scale, weights (N, S) cone weights summing to 1 per ommatidium -> responses (..., N, 3).
    """
    c = (np.asarray(sample_rgb01, dtype=np.float64) * np.asarray(weights, dtype=np.float64)[..., None]).sum(axis=-2)
    return np.clip(c * np.asarray(w_rgb, dtype=np.float64) * float(gain), 0., 1.)


def yaw_only(rot, tol=0.0):
    """yaw (rad) if `rot` is a pure rotation about +Z (quaternion x = y = 0 exactly), else None."""
    x, y, z, w = np.asarray(rot.as_quat(), dtype=np.float64).reshape(4)
    if abs(x) > tol or abs(y) > tol:
        return None
    return 2.0 * np.arctan2(z, w) if w >= 0 else 2.0 * np.arctan2(-z, -w)


_EMPTY = {}


def _empty_world(device):
    """a world without triangles (sky-only calls)"""
    from .env.world import WorldBase
    if device not in _EMPTY:
        _EMPTY[device] = WorldBase(np.zeros((0, 3, 3), dtype=np.float32), np.zeros((0, 3), dtype=np.float32), device=device)
    return _EMPTY[device]


def dorsal_rim_layout(nb_input=60, field_of_view=56., degrees=True):
    """
    Ommatidia of a dorsal-rim polarisation sensor: `nb_input` viewing directions on a Fibonacci lattice over the cap of
    full opening angle `field_of_view` around the zenith (InvertPy's own layout is not vendored: this is the library's,
    documented here).  Returns the orientations as a SciPy Rotation (direction = ori.apply([1, 0, 0])).
    """
    fov = np.deg2rad(field_of_view) if degrees else float(field_of_view)
    i = np.arange(nb_input, dtype=np.float64)
    yaw = (np.pi * (1. + np.sqrt(5.)) * (i + .5)) % (2 * np.pi) - np.pi
    z = 1. - (1. - np.cos(fov / 2.)) * (i + .5) / nb_input          # uniform in solid angle over the cap
    pitch = np.arccos(z) - np.pi / 2                                   # negative pitch looks up (world.py:84-90)
    return R.from_euler('ZY', np.c_[yaw, pitch])


class CompoundEye(object):
    def __init__(self, nb_input=1000, omm_ori=None, omm_rho=np.deg2rad(5.), omm_res=1., omm_pol_op=0.,
                 c_sensitive=(0., 0., 1., 0., 0.), noise=0., xyz=None, ori=None, nb_samples=1, dtype='float32',
                 name="compound_eye", device=0, rng=RNG):
        if omm_ori is None:
            pitch, yaw = fibonacci_lattice(nb_input)
            omm_ori = R.from_euler('ZY', np.c_[yaw, pitch])
        self._omm_ori = omm_ori
        self.nb_ommatidia = len(omm_ori)
        self.omm_rho = float(np.max(omm_rho))
        self.omm_res = omm_res
        self.omm_pol_op = omm_pol_op
        self.c_sensitive = np.asarray(c_sensitive, dtype=np.float64).reshape(-1)
        self.noise = noise
        self.rng = rng
        self.nb_samples = int(nb_samples)
        self.dtype = dtype
        self.name = name
        self.device = device
        self._xyz = np.zeros(3) if xyz is None else np.asarray(xyz, dtype=np.float64)
        self._ori = R.from_euler('Z', 0.) if ori is None else ori
        self.responses = np.zeros((self.nb_ommatidia, 3), dtype=dtype)
        self._layout = None
        self._renderers = {}
        self._host_out = {}          # the call's host result buffer, re-used from step to step
        self._sensor3 = None

    # ---- geometry ----
    @property
    def omm_ori(self):
        return self._omm_ori

    def _eye_layout(self):
        if self._layout is None:
            yaw, pitch, _ = self._omm_ori.as_euler('ZYX').T
            dp, dy, w = cone_offsets(self.nb_samples, self.omm_rho)
            p = np.clip(pitch[:, None] + dp[None, :], -np.pi / 2 + 1e-6, np.pi / 2 - 1e-6)
            y = (yaw[:, None] + dy[None, :] / np.maximum(np.cos(pitch), 0.05)[:, None] + np.pi) % (2 * np.pi) - np.pi
            self._layout = EyeLayout.from_angles(p.ravel(), y.ravel(), samples=self.nb_samples,
                                                 weights=np.tile(w, self.nb_ommatidia))
        return self._layout

    def _renderer(self, scene, sky):
        key = (id(scene), id(sky))
        if key not in self._renderers:
            self._renderers[key] = Renderer(scene, self._eye_layout(), sky=sky, device=self.device)
        return self._renderers[key]

    @property
    def w_rgb(self):
        """spectral sensitivity to (R, G, B): entries 1..3 of the IRGBU vector `c_sensitive`"""
        return self.c_sensitive[1:4] if self.c_sensitive.size >= 4 else np.ones(3)

    def sensor_params(self, channels=3):
        """the eye's photoreceptor transform as the C ABI takes it (isy_sensor_params)"""
        return _lib.SensorParams.make(self.w_rgb, float(np.max(self.omm_res)), channels, float(np.max(self.omm_pol_op)))

    nb_channels = 3      # values per ommatidium in `responses`

    # ---- pose (reference: eye.xyz = ..., eye.ori = R.from_euler('Z', yaw, degrees=True)) ----
    @property
    def xyz(self):
        return self._xyz

    @xyz.setter
    def xyz(self, v):
        self._xyz = np.asarray(v, dtype=np.float64)

    @property
    def ori(self):
        return self._ori

    @ori.setter
    def ori(self, v):
        self._ori = v

    x = property(lambda self: self._xyz[0])
    y = property(lambda self: self._xyz[1])
    z = property(lambda self: self._xyz[2])

    @property
    def yaw(self):
        return self._ori.as_euler('ZYX')[0]

    @property
    def yaw_deg(self):
        return np.rad2deg(self.yaw)

    # ---- sensing ----
    def __call__(self, sky=None, scene=None, callback=None):
        if scene is None:
            if sky is None:
                raise ValueError("the eye needs a sky and / or a scene to look at")
            scene = _empty_world(self.device)          # eye(sky=sky): the sky alone (examples/test_sky.py:22)
        r = self._renderer(scene, sky)
        yaw = yaw_only(self._ori)
        if self._sensor3 is None:
            self._sensor3 = self.sensor_params(3 if self.nb_channels == 3 else 0)
        if yaw is not None:     # the pose of every simulation step of the reference: rasteriser for yaw-only views
            out = r.render_host(self._xyz.reshape(1, 3), yaw=np.array([[yaw]]), outputs=("resp",),
                                sensor=self._sensor3, out=self._host_out)
        else:
            out = r.render_host(self._xyz.reshape(1, 3), quat=self._ori.as_quat().reshape(1, 1, 4), outputs=("resp",),
                                sensor=self._sensor3, out=self._host_out)
        resp = out["resp"].numpy()[0].astype(self.dtype, copy=True).reshape(self.nb_ommatidia, -1)
        if self.noise:
            resp[add_noise(noise=self.noise, shape=resp.shape, rng=self.rng)] = 0.      # world.py:125-127
        self.responses = resp
        if callback is not None:
            callback(self)
        return self.responses

    def render_batch(self, scene, sky, pos, yaw_deg, channels=3, out=None):
        """
        Batched equivalent of calling the eye at P positions x H yaw-only headings: (P*H, N, 3) responses, or
        (P*H, N) PN inputs with channels=1 (agent.py:744), transformed on the device; `out` = a pinned host
        tensor to fill (re-used by the chunked dataset writer).
        """
        r = self._renderer(scene, sky)
        res = {} if out is None else {"resp": out}
        res = r.render_host(pos, np.deg2rad(yaw_deg), outputs=("resp",), sensor=self.sensor_params(channels), out=res)
        return res["resp"].numpy()


class PolarisationSensor(CompoundEye):
    """
    The dorsal-rim polarisation sensor of the reference's path-integration agents (`PolarisationSensor(nb_input=60,
    field_of_view=56, degrees=True, noise=, rng=)`, agent/agent.py:835, :1385; called as `pol_sensor(sky=, scene=)`,
    :879).  One polarisation-sensitive photoreceptor per ommatidium (`omm_pol_op` = 1 by default):
        responses (N, 1) = omm_res * sqrt(Y * (rho (sin^2(A - phi) + cos^2(A - phi)(1 - P)^2) + 1 - rho))
    evaluated on the device from the sky model's Y / DoP / AoP in each viewing direction (include/isy_b200.h,
    isy_sensor_params, channels = 0); a sample that a triangle or the ground hides contributes nothing.
    """
    nb_channels = 1

    def __init__(self, nb_input=60, field_of_view=56., degrees=True, omm_ori=None, omm_pol_op=1., omm_rho=np.deg2rad(5.4),
                 omm_res=1., c_sensitive=(0., 0., 0., 0., 1.), **kwargs):
        if omm_ori is None:
            omm_ori = dorsal_rim_layout(nb_input, field_of_view, degrees)
        self.field_of_view = np.deg2rad(field_of_view) if degrees else float(field_of_view)
        CompoundEye.__init__(self, nb_input=nb_input, omm_ori=omm_ori, omm_rho=omm_rho, omm_res=omm_res,
                             omm_pol_op=omm_pol_op, c_sensitive=c_sensitive, **kwargs)
        self.responses = np.zeros((self.nb_ommatidia, 1), dtype=self.dtype)
