"""
TEST INFRASTRUCTURE ONLY -- builds and wraps oracle/isy_oracle.c (ctypes).

Only tests/, __graft_entry__ (build + smoke check) and bench.py's cpu_baseline legs use it.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "isy_oracle.c")
LIB = os.path.join(HERE, "libisy_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fopenmp",
               SRC, "-o", LIB, "-lm"]
        subprocess.check_call(cmd)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        dp = ctypes.POINTER(ctypes.c_double)
        _lib.isy_oracle_world_hits.restype = ctypes.c_long
        _lib.isy_oracle_world_hits.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_long, ctypes.c_double, dp, dp,
                                               dp, ctypes.c_long, ctypes.POINTER(ctypes.c_int32), dp]
        _lib.isy_oracle_sky.restype = None
        _lib.isy_oracle_sky.argtypes = [dp, ctypes.c_long, ctypes.c_double, ctypes.c_double, dp, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_double, dp, dp, dp, dp]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def world_hits(polygons, horizon, pos, pitch, yaw):
    """(hit int32 (N,), depth float64 (N,), n_visible) -- see isy_oracle.c:isy_oracle_world_hits."""
    tri = np.ascontiguousarray(polygons, dtype=np.float32).reshape(-1)
    pos = np.ascontiguousarray(np.nanmean(np.atleast_2d(np.asarray(pos, dtype=np.float64)), axis=0))
    pitch = np.ascontiguousarray(pitch, dtype=np.float64)
    yaw = np.ascontiguousarray(yaw, dtype=np.float64)
    n = pitch.size
    hit = np.empty(n, dtype=np.int32)
    depth = np.empty(n, dtype=np.float64)
    nv = lib().isy_oracle_world_hits(_p(tri, ctypes.c_float), tri.size // 9, float(horizon), _p(pos, ctypes.c_double),
                                     _p(pitch, ctypes.c_double), _p(yaw, ctypes.c_double), n,
                                     _p(hit, ctypes.c_int32), _p(depth, ctypes.c_double))
    if nv < 0:
        raise MemoryError
    return hit.reshape(pitch.shape), depth.reshape(pitch.shape), int(nv)


def sky(dirs, theta_s, phi_s, coef, tau_L, c1, c2, sun_xy):
    dirs = np.ascontiguousarray(dirs, dtype=np.float64)
    n = dirs.shape[0]
    coef = np.ascontiguousarray(coef, dtype=np.float64)
    sun_xy = np.ascontiguousarray(sun_xy, dtype=np.float64)
    Y, P, A = (np.empty(n) for _ in range(3))
    d = ctypes.c_double
    lib().isy_oracle_sky(_p(dirs, d), n, float(theta_s), float(phi_s), _p(coef, d), float(tau_L), float(c1), float(c2),
                         _p(sun_xy, d), _p(Y, d), _p(P, d), _p(A, d))
    return Y, P, A
