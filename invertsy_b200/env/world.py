"""
GPU drop-in for ``invertsy.env.world`` (reference src/invertsy/env/world.py).

``WorldBase.__call__(pos, ori, init_c, noise, eta, rng)`` keeps the reference signature,
argument meaning, output shape / dtype / NaN conventions and error behaviour
(world.py:55-129) but renders on the GPU through the C ABI (``isy_world_call_host``,
include/isy_b200.h).  The class also hands its device-resident triangle soup to the batched
:class:`invertsy_b200.render.Renderer`.

No CPU fallback: if the CUDA library or a GPU is missing the call raises.
"""
import ctypes
import os

import numpy as np

from .. import _lib
from .occlusion import RNG, add_noise, __data__

GRASS_COLOUR = (0, 255, 0)        # world.py:35
GROUND_COLOUR = (229, 183, 90)    # world.py:36
SKY_COLOUR = (13, 135, 201)       # world.py:37


def luminance2skycolour(y):
    """Blue shade of the sky on the 0..255 scale for a luminance array (world.py:465-479)."""
    y = np.asarray(y)
    return np.clip(19. * y[..., np.newaxis] + np.asarray(SKY_COLOUR, dtype=float)[np.newaxis, :], 0, 255)


class WorldBase(object):
    """Triangle world seen from a position (reference ``WorldBase``, world.py:45-167)."""
    __data_dir__ = __data__

    def __init__(self, polygons, colours, horizon=2., dtype='float32', name='my_world', device=0):
        self.__polygons = polygons
        self.__colours = colours
        self.__horizon = horizon
        self.dtype = dtype
        self.name = name
        self.device = device
        self.__handles = {}          # device -> (handle, key of the arrays it was built from)

    # ---- device handles (lazy, one per device; rebuilt if the arrays were replaced) ----
    def _device_world(self, device=None):
        """
        The isy_world handle of this world on `device` (default: self.device).  Every device keeps its own
        handle for as long as the world lives, so several Renderers (one per GPU) can share one world.
        """
        device = int(self.device if device is None else device)
        key = (id(self.__polygons), id(self.__colours), float(self.__horizon))
        have = self.__handles.get(device)
        if have is None or have[1] != key:
            poly = np.asarray(self.__polygons)
            if poly.dtype == np.float64 and not np.array_equal(poly.astype(np.float32).astype(np.float64), poly, equal_nan=True):
                raise ValueError("float64 polygons that are not exactly representable in float32 are not supported: "
                                 "the device copy of the world is float32 (the reference's own worlds are, world.py:226)")
            tri = np.ascontiguousarray(self.__polygons, dtype=np.float32).reshape(-1, 9)
            rgb = np.ascontiguousarray(self.__colours, dtype=np.float32).reshape(-1, 3)
            if tri.shape[0] != rgb.shape[0]:
                raise ValueError("polygons and colours disagree on the number of triangles")
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().isy_world_create(_lib.ptr(tri), _lib.ptr(rgb), tri.shape[0], float(self.__horizon),
                                                   device, ctypes.byref(h)))
            if have is not None:
                # the arrays were replaced: handles built from the old ones are retired, not destroyed -- a Renderer
                # may still hold one -- and freed with the world
                self.__retired = getattr(self, "_WorldBase__retired", []) + [have[0]]
            self.__handles[device] = (h, key)
        return self.__handles[device][0]

    def _release(self):
        handles = [h for h, _ in self.__handles.values()] + getattr(self, "_WorldBase__retired", [])
        self.__handles = {}
        self.__retired = []
        for h in handles:
            try:
                _lib.lib().isy_world_destroy(h)
            except Exception:
                pass

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def __call__(self, pos, ori=None, init_c=None, noise=0., eta=None, rng=RNG):
        """
        RGB seen by each viewing direction (world.py:55-129).

        pos: (k, 3) positions -- only their nan-mean is used (world.py:94)
        ori: scipy Rotation of length N (global ommatidia orientations)
        init_c: optional (N,) background luminance -> sky-blue init on the 0..255 scale (world.py:82)
        Returns (N, 3): ``self.dtype`` with NaN for "nothing seen" when init_c is None, float64 otherwise.
        """
        n = np.shape(ori)[0]          # IndexError for ori=None / a single rotation, as in world.py:80
        quat = _lib.as_f64(ori.as_quat()).reshape(n, 4)
        centre = _lib.as_f64(np.nanmean(np.asarray(pos), axis=0)).reshape(3)
        rgb = np.empty((n, 3), dtype=np.float64)
        lum = None
        if init_c is not None:
            lum = _lib.as_f64(init_c).reshape(-1)
            if lum.size != n:
                raise ValueError("init_c must hold one luminance per orientation")
        _lib.check(_lib.lib().isy_world_call_host(self._device_world(), _lib.ptr(centre), _lib.ptr(quat), n,
                                                  _lib.ptr(lum), _lib.ptr(rgb), None, None))
        c = rgb.astype(self.dtype) if init_c is None else rgb
        if eta is None:
            eta = add_noise(noise=noise, shape=c.shape, rng=rng)
        c[eta] = 0.
        return c

    def hits(self, pos, ori):
        """
        Extra (not in the reference): per direction the index of the triangle the painter's loop
        leaves on top (-1 = none) and its min-vertex distance (NaN = none).
        """
        n = np.shape(ori)[0]
        quat = _lib.as_f64(ori.as_quat()).reshape(n, 4)
        centre = _lib.as_f64(np.nanmean(np.asarray(pos), axis=0)).reshape(3)
        hit = np.empty(n, dtype=np.int32)
        depth = np.empty(n, dtype=np.float64)
        _lib.check(_lib.lib().isy_world_call_host(self._device_world(), _lib.ptr(centre), _lib.ptr(quat), n, None, None,
                                                  _lib.ptr(hit), _lib.ptr(depth)))
        return hit, depth

    @property
    def polygons(self):
        """(T, 3 vertices, 3 coordinates) triangle soup (world.py:131-140)."""
        return self.__polygons

    @property
    def colours(self):
        """(T, 3) RGB per triangle (world.py:142-151)."""
        return self.__colours

    @property
    def horizon(self):
        """Radius (m) beyond which triangles are not rendered (world.py:153-163)."""
        return self.__horizon

    @staticmethod
    def set_data_dir(data_dir):
        WorldBase.__data_dir__ = data_dir


def _load_mat_or_npz(filename, packaged):
    """A ``.mat`` path is read with scipy (as the reference does); otherwise the packaged .npz."""
    if filename is not None and os.path.exists(filename):
        if filename.endswith(".npz"):
            return np.load(filename)
        from scipy.io import loadmat
        return loadmat(filename)
    return np.load(os.path.join(__data__, packaged))


class Seville2009(WorldBase):
    """Seville ant world: 5000 grass triangles + recorded ant routes (world.py:170-319)."""
    __data_dir__ = os.path.join(__data__, "Seville2009_world")
    WORLD_FILENAME = "world5000_gray.mat"
    ROUTES_FILENAME = "AntRoutes.mat"

    def __init__(self, horizon=2., dtype='float32', name="seville_2009", device=0):
        polygons, colours = Seville2009.load_world()
        super().__init__(polygons=polygons, colours=colours, horizon=horizon, dtype=dtype, name=name, device=device)

    @staticmethod
    def load_world(world_filename=WORLD_FILENAME, dtype='float32'):
        """
        (polygons (T,3,3), colours (T,3)) with the reference's conventions (world.py:195-230):
        coordinates stored as (Y, X, Z) of the MATLAB file, colour = green channel of ``colp`` only.
        """
        if not os.path.exists(os.path.dirname(world_filename)):
            world_filename = os.path.join(Seville2009.__data_dir__, world_filename)
        mat = _load_mat_or_npz(world_filename, "seville2009_world.npz")
        polygons = np.stack([mat['Y'], mat['X'], mat['Z']], axis=-1).astype(dtype)
        colours = np.zeros(mat['colp'].shape, dtype=dtype)
        colours[:, 1] = np.asarray(mat['colp'], dtype=dtype)[:, 1]
        return polygons, colours

    @staticmethod
    def load_routes(routes_filename=ROUTES_FILENAME, degrees=False):
        """
        Recorded ant routes as {'ant_no', 'route_no', 'path'} with path rows (y, x, 0.01, yaw)
        in metres and radians (or degrees) -- conventions of world.py:232-273.
        """
        if not os.path.exists(os.path.dirname(routes_filename)):
            routes_filename = os.path.join(Seville2009.__data_dir__, routes_filename)
        mat = _load_mat_or_npz(routes_filename, "seville2009_routes.npz")
        keys = set(mat.keys())
        routes = {"ant_no": [], "route_no": [], "path": []}
        ant = 1
        while "Ant%d_Route1" % ant in keys:
            route = 1
            while "Ant%d_Route%d" % (ant, route) in keys:
                raw = np.array(mat["Ant%d_Route%d" % (ant, route)], dtype=np.float64)
                raw[:, :2] /= 100.                                   # cm -> m
                xs, ys, phis = raw.T
                phis = (90 - phis + 180) % 360 - 180                 # MATLAB heading -> yaw
                path = np.zeros((xs.shape[0], 4))
                path[:, 0], path[:, 1], path[:, 2] = ys, xs, 0.01
                path[:, 3] = phis if degrees else np.deg2rad(phis)
                routes["ant_no"].append(ant)
                routes["route_no"].append(route)
                routes["path"].append(path)
                route += 1
            ant += 1
        return routes

    @staticmethod
    def load_route(name):
        """(path, ant, route) of a saved route file (world.py:275-298)."""
        path = f"{name}.npz"
        if not os.path.exists(os.path.dirname(path)):
            path = os.path.join(__data__, "routes", path)
        data = np.load(path)
        return data["path"], data["ant"], data["route"]

    @staticmethod
    def save_route(name, **rt):
        """Saves route arrays under data/routes (world.py:300-315; written with savez)."""
        path = f"{name}.npz"
        if not os.path.exists(os.path.dirname(path)):
            path = os.path.join(__data__, "routes", path)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez(path, **rt)


class SimpleWorld(WorldBase):
    """Sparse world: 3222 short blades, rescaled into the 10 m arena (world.py:317-462)."""
    __data_dir__ = os.path.join(__data__, "Sparse_world")
    WORLD_FILENAME = "world_gray.mat"

    def __init__(self, horizon=2., dtype='float32', name="sparse_world", device=0):
        polygons, colours = SimpleWorld.load_world()
        super().__init__(polygons=polygons, colours=colours, horizon=horizon, dtype=dtype, name=name, device=device)

    @staticmethod
    def load_world(world_filename=WORLD_FILENAME, dtype='float32'):
        """Scale 0.1 / 0.1 / 0.05 and offset 5 / 5 / 0 of the raw file, constant pale green (world.py:420-462)."""
        if not os.path.exists(os.path.dirname(world_filename)):
            world_filename = os.path.join(SimpleWorld.__data_dir__, world_filename)
        mat = _load_mat_or_npz(world_filename, "sparse_world.npz")
        y = mat['Y'] * 0.1 + 5.
        x = mat['X'] * 0.1 + 5.
        z = mat['Z'] * 0.05 + 0.
        polygons = np.stack([y, x, z], axis=-1).astype(dtype)
        colours = np.tile(np.array([0.8, 1.0, 0.8], dtype=dtype), (polygons.shape[0], 1))
        return polygons, colours
