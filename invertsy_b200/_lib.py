"""
ctypes binding of the C-ABI library (include/isy_b200.h).  There is no CPU fallback: if the
CUDA library is missing or fails, every compute entry point raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISY_B200_LIB", os.path.join(HERE, "libisy_b200.so"))   # override: profiling builds

ABI_VERSION = 2
ISY_OK, ISY_ERR_INVALID, ISY_ERR_CUDA, ISY_ERR_WORKSPACE, ISY_ERR_OVERFLOW = 0, -1, -2, -3, -4
SKY_NONE, SKY_UNIFORM, SKY_PEREZ = 0, 1, 2
FLAG_INIT_FROM_SKY, FLAG_BRUTE_FORCE, FLAG_SUN_PER_VIEW, FLAG_OUT_F64, FLAG_PER_VIEW_SKY, FLAG_NO_TILES = 1, 2, 4, 8, 16, 32

# every symbol include/isy_b200.h declares (tests check that the library exports all of them)
EXPORTS = (
    "isy_abi_version", "isy_last_error", "isy_launch_count", "isy_debug_bounds_violations",
    "isy_world_create", "isy_world_destroy", "isy_world_triangles",
    "isy_eye_create", "isy_eye_destroy", "isy_eye_rays",
    "isy_render_workspace_bytes", "isy_render", "isy_workspace_status", "isy_workspace_phase_cycles",
    "isy_render_host",
    "isy_world_call_host", "isy_sky_call_host", "isy_spectrum_influence_host", "isy_spectrum_influence_views",
    "isy_memory_create", "isy_memory_destroy", "isy_memory_size",
    "isy_familiarity_workspace_bytes", "isy_familiarity", "isy_familiarity_unresolved", "isy_familiarity_candidates",
    "isy_familiarity_host",
    "isy_willshaw_create", "isy_willshaw_destroy", "isy_willshaw_reset", "isy_willshaw_active", "isy_willshaw_train",
    "isy_willshaw_familiarity",
)


class SkyParams(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("reserved", ctypes.c_int32), ("luminance", ctypes.c_double),
                ("A", ctypes.c_double), ("B", ctypes.c_double), ("C", ctypes.c_double), ("D", ctypes.c_double),
                ("E", ctypes.c_double), ("tau_L", ctypes.c_double), ("c1", ctypes.c_double), ("c2", ctypes.c_double)]


class Outputs(ctypes.Structure):
    _fields_ = [("rgb", ctypes.c_void_p), ("hit", ctypes.c_void_p), ("depth", ctypes.c_void_p),
                ("Y", ctypes.c_void_p), ("dop", ctypes.c_void_p), ("aop", ctypes.c_void_p), ("resp", ctypes.c_void_p)]


class SensorParams(ctypes.Structure):
    """isy_sensor_params: the photoreceptor transform behind the `resp` output (include/isy_b200.h)."""
    _fields_ = [("w_rgb", ctypes.c_float * 3), ("gain", ctypes.c_float), ("channels", ctypes.c_int32),
                ("pol_op", ctypes.c_float)]

    @staticmethod
    def make(w_rgb=(1., 1., 1.), gain=1., channels=3, pol_op=0.):
        """channels: 3 = (N, 3) responses, 1 = PN inputs (N,), 0 = polarisation photoreceptors (N,)"""
        sp = SensorParams()
        sp.w_rgb[0], sp.w_rgb[1], sp.w_rgb[2] = (float(v) for v in w_rgb)
        sp.gain, sp.channels, sp.pol_op = float(gain), int(channels), float(pol_op)
        return sp


class IsyError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("isy_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def lib():
    """Loads libisy_b200.so (built by invertsy_b200.build / __graft_entry__.build). Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -m invertsy_b200.build` "
                          "(nvcc, sm_100a). invertsy_b200 has no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_double
    L.isy_abi_version.restype = ctypes.c_int
    L.isy_last_error.restype = ctypes.c_char_p
    L.isy_launch_count.restype = i64
    L.isy_debug_bounds_violations.argtypes = [ctypes.c_int, ctypes.POINTER(i32)]
    L.isy_debug_bounds_violations.restype = i64
    L.isy_world_create.argtypes = [vp, vp, i64, dbl, ctypes.c_int, ctypes.POINTER(vp)]
    L.isy_world_destroy.argtypes = [vp]
    L.isy_world_triangles.argtypes = [vp]
    L.isy_world_triangles.restype = i64
    L.isy_eye_create.argtypes = [vp, vp, i64, i32, ctypes.c_int, ctypes.POINTER(vp)]
    L.isy_eye_destroy.argtypes = [vp]
    L.isy_eye_rays.argtypes = [vp]
    L.isy_eye_rays.restype = i64
    L.isy_render_workspace_bytes.argtypes = [vp, vp, i64, i64]
    L.isy_render_workspace_bytes.restype = ctypes.c_size_t
    L.isy_render.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp, ctypes.POINTER(SkyParams), ctypes.POINTER(SensorParams),
                             ctypes.c_uint32, ctypes.POINTER(Outputs), vp, ctypes.c_size_t, vp]
    L.isy_workspace_status.argtypes = [vp, vp]
    L.isy_workspace_phase_cycles.argtypes = [vp, vp, vp]
    L.isy_workspace_phase_cycles.restype = ctypes.c_int
    L.isy_render_host.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp, ctypes.POINTER(SkyParams),
                                  ctypes.POINTER(SensorParams), ctypes.c_uint32, ctypes.POINTER(Outputs)]
    L.isy_world_call_host.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    L.isy_sky_call_host.argtypes = [ctypes.POINTER(SkyParams), dbl, dbl, vp, i64, vp, ctypes.c_int, vp, vp, vp, vp, vp]
    L.isy_spectrum_influence_host.argtypes = [vp, i64, vp, i64, ctypes.c_int, vp, vp]
    L.isy_spectrum_influence_views.argtypes = [vp, i64, i64, vp, i64, ctypes.c_uint32, ctypes.c_int, vp]
    L.isy_spectrum_influence_views.restype = ctypes.c_int
    L.isy_memory_create.argtypes = [vp, i64, i64, ctypes.c_int, ctypes.POINTER(vp)]
    L.isy_memory_destroy.argtypes = [vp]
    L.isy_memory_size.argtypes = [vp]
    L.isy_memory_size.restype = i64
    L.isy_familiarity_workspace_bytes.argtypes = [vp, i64]
    L.isy_familiarity_workspace_bytes.restype = ctypes.c_size_t
    L.isy_familiarity.argtypes = [vp, vp, i64, i64, vp, vp, vp, ctypes.c_size_t, vp]
    L.isy_familiarity_host.argtypes = [vp, vp, i64, vp, vp]
    L.isy_familiarity_unresolved.argtypes = [vp, i64, vp, vp, ctypes.POINTER(i64)]
    L.isy_familiarity_candidates.argtypes = [vp, i64, vp, vp, vp, vp]
    L.isy_familiarity_candidates.restype = ctypes.c_int
    L.isy_willshaw_create.argtypes = [i64, i64, i32, dbl, ctypes.c_uint32, ctypes.c_int, ctypes.POINTER(vp)]
    L.isy_willshaw_destroy.argtypes = [vp]
    L.isy_willshaw_reset.argtypes = [vp]
    L.isy_willshaw_active.argtypes = [vp]
    L.isy_willshaw_active.restype = i64
    L.isy_willshaw_train.argtypes = [vp, vp, i64, i64, vp]
    L.isy_willshaw_familiarity.argtypes = [vp, vp, i64, i64, vp, vp]
    for name in ("isy_willshaw_create", "isy_willshaw_destroy", "isy_willshaw_reset", "isy_willshaw_train",
                 "isy_willshaw_familiarity"):
        getattr(L, name).restype = ctypes.c_int
    for name in ("isy_memory_create", "isy_memory_destroy", "isy_familiarity", "isy_familiarity_host",
                 "isy_familiarity_unresolved"):
        getattr(L, name).restype = ctypes.c_int
    for name in ("isy_world_create", "isy_world_destroy", "isy_eye_create", "isy_eye_destroy", "isy_render",
                 "isy_workspace_status", "isy_render_host", "isy_world_call_host", "isy_sky_call_host",
                 "isy_spectrum_influence_host"):
        getattr(L, name).restype = ctypes.c_int
    if L.isy_abi_version() != ABI_VERSION:
        raise ImportError("libisy_b200.so ABI version mismatch")
    _lib = L
    return L


def check(code):
    if code != ISY_OK:
        raise IsyError(code, lib().isy_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a C-contiguous NumPy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def launch_count():
    return int(lib().isy_launch_count())


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
