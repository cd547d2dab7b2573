"""
Batched GPU rendering of compound-eye views: the B200-native entry point of the package.

    world = Seville2009()                       # drop-in world (env/world.py)
    eye = EyeLayout.from_angles(*fibonacci_lattice(1000))
    sky = Sky(np.deg2rad(60), np.pi)            # drop-in sky (env/sky.py)
    r = Renderer(world, eye, sky)
    out = r.render(pos, yaw)                    # pos (P,3), yaw (P,H) radians -> dict of CUDA tensors

A *view* is one (position, orientation) pair; P positions x H orientations per position give
B = P*H views ordered b = p*H + h.  The cull / angular projection of the world
(reference world.py:94-110) is done once per position and shared by its H orientations, which is
the shape of the familiarity-grid workloads (reference sim/simulation.py:2675-2707).

PyTorch is used only for device memory and streams; all compute is in libisy_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .env.sky import Sky, UniformSky

_OUT_RAYS = ("hit", "depth")
_OUT_OMM = ("rgb", "Y", "dop", "aop")
_OUT_ALL = _OUT_RAYS + _OUT_OMM + ("resp",)


def fibonacci_lattice(n, hemisphere=False):
    """
    Deterministic ommatidia directions: azimuth phi_i = (pi (1+sqrt 5)(i+1/2)) mod 2pi - pi,
    z_i = 1 - 2(i+1/2)/n (full sphere) or 1 - (i+1/2)/n (upper hemisphere).
    Returns (pitch, yaw); pitch = arccos(z) - pi/2, negative pitch looks up (reference world.py:84-90).
    """
    i = np.arange(n, dtype=np.float64)
    yaw = (np.pi * (1. + np.sqrt(5.)) * (i + .5)) % (2 * np.pi) - np.pi
    z = 1. - (1. if hemisphere else 2.) * (i + .5) / n
    return np.arccos(z) - np.pi / 2, yaw


def angles_to_quat(pitch, yaw):
    """Quaternions (x,y,z,w) of R.from_euler('ZY', [yaw, pitch]): direction (cos p cos y, cos p sin y, -sin p)."""
    pitch = np.asarray(pitch, dtype=np.float64)
    yaw = np.asarray(yaw, dtype=np.float64)
    sy, cy = np.sin(yaw / 2), np.cos(yaw / 2)
    sp, cp = np.sin(pitch / 2), np.cos(pitch / 2)
    # q = qz(yaw) * qy(pitch)
    return np.stack([-sy * sp, cy * sp, sy * cp, cy * cp], axis=-1)


class EyeLayout(object):
    """
    Eye-frame orientations of the rays of a compound eye: N ommatidia x S cone samples, ray
    r = n*S + s.  This is the `omm_ori` a CompoundEye composes with its own orientation before
    calling scene()/sky() (reference agent.py:742 -> world.py:55; examples/test_sky.py:25).
    """

    def __init__(self, quat, samples=1, weights=None):
        quat = np.ascontiguousarray(quat, dtype=np.float64).reshape(-1, 4)
        if quat.shape[0] % samples:
            raise ValueError("number of rays must be a multiple of samples")
        self.quat = quat
        self.samples = int(samples)
        self.nb_ommatidia = quat.shape[0] // self.samples
        self.weights = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32).reshape(-1)
        if self.weights is not None and self.weights.size != quat.shape[0]:
            raise ValueError("one weight per ray expected")
        self._handles = {}

    @staticmethod
    def from_rotation(ori, samples=1, weights=None):
        return EyeLayout(ori.as_quat(), samples, weights)

    @staticmethod
    def from_angles(pitch, yaw, samples=1, weights=None):
        return EyeLayout(angles_to_quat(pitch, yaw), samples, weights)

    @property
    def nb_rays(self):
        return self.quat.shape[0]

    def _device_eye(self, device):
        h = self._handles.get(device)
        if h is None:
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().isy_eye_create(_lib.ptr(self.quat), _lib.ptr(self.weights), self.nb_ommatidia,
                                                 self.samples, int(device), ctypes.byref(h)))
            self._handles[device] = h
        return h

    def __del__(self):
        for h in self._handles.values():
            try:
                _lib.lib().isy_eye_destroy(h)
            except Exception:
                pass
        self._handles = {}


def _sky_params(sky):
    if sky is None:
        sp = _lib.SkyParams()
        sp.kind = _lib.SKY_NONE
        return sp, None
    sp = sky._params()
    sun = (float(sky.theta_s), float(sky.phi_s)) if isinstance(sky, Sky) else None
    return sp, sun


class Renderer(object):
    """Renders batches of views of one world through one eye layout on one GPU."""

    def __init__(self, world, eye, sky=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("invertsy_b200.Renderer needs a CUDA device (no CPU fallback)")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.world = world
        self.eye = eye
        self.sky = sky
        self._w = world._device_world(self.device)   # (the world keeps one handle per device)
        self._e = eye._device_eye(self.device)
        self._workspace = None
        self._inflight = None

    # ---- helpers ----
    def _dev(self, a, shape):
        if a is None:
            return None
        if not torch.is_tensor(a):
            a = torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
        a = a.to(device="cuda:%d" % self.device, dtype=torch.float64).contiguous()
        if tuple(a.shape) != tuple(shape):
            a = a.reshape(shape)
        return a

    def _ws(self, P, H):
        need = int(_lib.lib().isy_render_workspace_bytes(self._w, self._e, P, H))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device="cuda:%d" % self.device)
        return self._workspace

    def render(self, pos, yaw=None, quat=None, sun=None, outputs=("rgb", "Y", "dop", "aop"), init_from_sky=False,
               brute_force=False, f64=False, out=None, sky=None, per_view_sky=False, sensor=None, irgbu=None,
               no_tiles=False):
        """
        irgbu: (rows, 5) spectral weights: the luminance of every view goes through spectrum_influence(y, irgbu).sum(axis=1)
        as `Sky.__call__(ori, irgbu=...)` does per call (sky.py:342-344); rows = 1 or a divisor of the ommatidia count.
        per_view_sky: evaluate the sky for every view even when all positions share their headings (cross-check).
        no_tiles: rasterise evenly spaced scans copy by copy instead of tile by tile (cross-check).
        sensor: _lib.SensorParams for the "resp" output -- (B,N,3) photoreceptor responses, or (B,N) PN inputs with
        channels=1 (include/isy_b200.h, isy_sensor_params).
        pos (P,3) metres; yaw (P,H) radians about +Z *or* quat (P,H,4) full eye orientations;
        sun: None (take the sky's), (2,) or (P,H,2) of (theta_s, phi_s).
        Returns {name: CUDA tensor}: rgb (B,N,3), Y/dop/aop (B,N), hit/depth (B,R), B = P*H.
        Asynchronous on the current torch stream.
        """
        sky = self.sky if sky is None else sky
        sp, sky_sun = _sky_params(sky)
        pos = self._dev(pos, (-1, 3))
        P = pos.shape[0]
        if quat is not None:
            H = quat.shape[1] if getattr(quat, "ndim", 0) == 3 else 1
            quat = self._dev(quat, (P, H, 4))
        else:
            yaw = np.zeros((P, 1)) if yaw is None else yaw
            H = yaw.shape[1] if getattr(yaw, "ndim", 0) == 2 else 1
            yaw = self._dev(yaw, (P, H))
        B = P * H
        flags = 0
        if sp.kind == _lib.SKY_PEREZ:
            sun = sky_sun if sun is None else sun
            sun_t = self._dev(sun, (-1, 2))
            if sun_t.shape[0] == B and B > 1:
                flags |= _lib.FLAG_SUN_PER_VIEW
            elif sun_t.shape[0] != 1:
                raise ValueError("sun must be (2,) or one (theta_s, phi_s) per view")
        else:
            sun_t = None
        if sp.kind == _lib.SKY_NONE:
            outputs = tuple(o for o in outputs if o not in ("Y", "dop", "aop"))
        if init_from_sky:
            flags |= _lib.FLAG_INIT_FROM_SKY
        if brute_force:
            flags |= _lib.FLAG_BRUTE_FORCE
        if per_view_sky:
            flags |= _lib.FLAG_PER_VIEW_SKY
        if no_tiles:
            flags |= _lib.FLAG_NO_TILES
        if f64:
            flags |= _lib.FLAG_OUT_F64
        fdt = torch.float64 if f64 else torch.float32
        N, R = self.eye.nb_ommatidia, self.eye.nb_rays
        dev = "cuda:%d" % self.device
        res = {} if out is None else out
        o = _lib.Outputs()
        for name in outputs:
            if name not in res:
                if name == "rgb":
                    res[name] = torch.empty((B, N, 3), dtype=fdt, device=dev)
                elif name == "hit":
                    res[name] = torch.empty((B, R), dtype=torch.int32, device=dev)
                elif name == "depth":
                    res[name] = torch.empty((B, R), dtype=fdt, device=dev)
                elif name in _OUT_OMM:
                    res[name] = torch.empty((B, N), dtype=fdt, device=dev)
                elif name == "resp":
                    if sensor is None:
                        raise ValueError("the 'resp' output needs sensor=SensorParams")
                    res[name] = torch.empty((B, N, 3) if sensor.channels == 3 else (B, N), dtype=fdt, device=dev)
                else:
                    raise ValueError("unknown output %r" % name)
            setattr(o, name, res[name].data_ptr())
        ws = self._ws(P, H)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().isy_render(self._w, self._e, P, H, pos.data_ptr(),
                                         None if yaw is None or quat is not None else yaw.data_ptr(),
                                         None if quat is None else quat.data_ptr(),
                                         None if sun_t is None else sun_t.data_ptr(),
                                         ctypes.byref(sp), None if sensor is None else ctypes.byref(sensor), flags,
                                         ctypes.byref(o), ws.data_ptr(), ws.numel(), ctypes.c_void_p(stream)))
        if irgbu is not None and "Y" in res and "Y" in outputs:
            w = torch.as_tensor(np.ascontiguousarray(irgbu, dtype=np.float64).reshape(-1, 5), device=dev)
            _lib.check(_lib.lib().isy_spectrum_influence_views(res["Y"].data_ptr(), B, N, w.data_ptr(), w.shape[0],
                                                               _lib.FLAG_OUT_F64 if f64 else 0, self.device,
                                                               ctypes.c_void_p(stream)))
            sun_t = (sun_t, w)
        # keep the input tensors alive until the stream has consumed them
        self._inflight = (pos, yaw, quat, sun_t)
        return res

    def check(self):
        """Synchronises and raises if a position overflowed the workspace slab."""
        if self._workspace is not None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().isy_workspace_status(self._workspace.data_ptr(), ctypes.c_void_p(stream)))

    def phase_cycles(self):
        """Profiling aid: (project, bin, rays, positions) SM-cycle sums of the last render()."""
        out = np.zeros(4, dtype=np.uint64)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().isy_workspace_phase_cycles(self._workspace.data_ptr(), ctypes.c_void_p(stream),
                                                         _lib.ptr(out)))
        return out

    def render_host(self, pos, yaw=None, quat=None, sun=None, outputs=("rgb", "Y", "dop", "aop"),
                    init_from_sky=False, f64=False, out=None, sky=None, per_view_sky=False, sensor=None):
        """
        End-to-end variant with HOST buffers (isy_render_host): inputs are NumPy / CPU tensors,
        results land in (pinned) host tensors; returns {name: CPU tensor}. Blocking.
        """
        sky = self.sky if sky is None else sky
        sp, sky_sun = _sky_params(sky)
        pos = _lib.as_f64(pos).reshape(-1, 3)
        P = pos.shape[0]
        if quat is not None:
            quat = _lib.as_f64(quat).reshape(P, -1, 4)
            H = quat.shape[1]
            yaw = None
        else:
            yaw = _lib.as_f64(np.zeros((P, 1)) if yaw is None else yaw).reshape(P, -1)
            H = yaw.shape[1]
        B = P * H
        flags = 0
        sun_a = None
        if sp.kind == _lib.SKY_PEREZ:
            sun_a = _lib.as_f64(sky_sun if sun is None else sun).reshape(-1, 2)
            if sun_a.shape[0] == B and B > 1:
                flags |= _lib.FLAG_SUN_PER_VIEW
            elif sun_a.shape[0] != 1:
                raise ValueError("sun must be (2,) or one (theta_s, phi_s) per view")
        if sp.kind == _lib.SKY_NONE:
            outputs = tuple(o for o in outputs if o not in ("Y", "dop", "aop"))
        if init_from_sky:
            flags |= _lib.FLAG_INIT_FROM_SKY
        if f64:
            flags |= _lib.FLAG_OUT_F64
        if per_view_sky:
            flags |= _lib.FLAG_PER_VIEW_SKY
        fdt = torch.float64 if f64 else torch.float32
        N, R = self.eye.nb_ommatidia, self.eye.nb_rays
        res = {} if out is None else out
        o = _lib.Outputs()
        for name in outputs:
            if name not in res:
                if name not in _OUT_ALL:
                    raise ValueError("unknown output %r" % name)
                if name == "resp" and sensor is None:
                    raise ValueError("the 'resp' output needs sensor=SensorParams")
                shape = (B, N, 3) if name == "rgb" or (name == "resp" and sensor.channels == 3) else \
                    (B, R) if name in _OUT_RAYS else (B, N)
                dt = torch.int32 if name == "hit" else fdt
                res[name] = torch.empty(shape, dtype=dt, pin_memory=True)
            setattr(o, name, res[name].data_ptr())
        _lib.check(_lib.lib().isy_render_host(self._w, self._e, P, H, _lib.ptr(pos), _lib.ptr(yaw), _lib.ptr(quat),
                                              _lib.ptr(sun_a), ctypes.byref(sp),
                                              None if sensor is None else ctypes.byref(sensor), flags, ctypes.byref(o)))
        return res
