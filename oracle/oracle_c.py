"""ctypes loader for oracle/liboracle.so (the C fp64 oracle; test infrastructure
only -- see oracle.c).  Builds the library with `make -C oracle` on first use."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KINDS = {"bilateral": 0, "photometric": 1, "spatial": 2}


class Params(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("channels", C.c_int), ("kind", C.c_int),
                ("h_loc", C.c_double), ("h_val", C.c_double), ("gain", C.c_double), ("power", C.c_double),
                ("m", C.c_int), ("gram_schmidt", C.c_int), ("row0", C.c_int), ("row1", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "oracle.c")):
            build()
        L = C.CDLL(path)
        L.orc_uniform_sampling.restype = C.c_uint
        L.orc_uniform_sampling.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_void_p, C.c_uint]
        L.orc_random_sampling.restype = C.c_uint
        L.orc_random_sampling.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_uint32, C.c_void_p]
        L.orc_synthetic_image.restype = None
        L.orc_synthetic_image.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p]
        L.orc_symeig.restype = C.c_int
        L.orc_symeig.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_pipeline.restype = C.c_int
        L.orc_pipeline.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p,
                                   C.POINTER(C.c_double), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_affinity_rows.restype = None
        L.orc_affinity_rows.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def uniform_sampling(width, height, requested):
    L = lib()
    cnt = L.orc_uniform_sampling(width, height, requested, None, 0)
    out = np.empty(cnt, dtype=np.uint32)
    L.orc_uniform_sampling(width, height, requested, out.ctypes.data, cnt)
    return out


def random_sampling(width, height, requested, seed):
    out = np.empty(requested, dtype=np.uint32)
    cnt = lib().orc_random_sampling(width, height, requested, seed, out.ctypes.data)
    return out[:cnt]


def synthetic_image(width, height, channels=1, seed=1234):
    out = np.empty((height, width, channels), dtype=np.uint8)
    lib().orc_synthetic_image(width, height, channels, seed, out.ctypes.data)
    return out[:, :, 0] if channels == 1 else out


def symeig(a):
    a = np.array(a, dtype=np.float64, order="C", copy=True)
    d = np.empty(a.shape[0])
    rc = lib().orc_symeig(a.ctypes.data, a.shape[0], d.ctypes.data)
    if rc:
        raise RuntimeError("orc_symeig failed")
    return d, a


def run_pipeline(img, sample_indices, m=-1, kind="bilateral", h_loc=40.0, h_val=30.0, gain=3.0,
                 power=1.0, orthonormalise=False, rows=None):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    s = np.ascontiguousarray(sample_indices, dtype=np.uint32)
    p = len(s)
    mm = m if (m is not None and 0 <= m < p) else p - 1
    r0, r1 = (0, H) if rows is None else rows
    P = Params(W, H, Cn, KINDS[kind], h_loc, h_val, gain, power, mm, int(orthonormalise), r0, r1)
    D = np.empty(p)
    mu = np.empty(mm)
    z = np.empty((H, W, Cn))
    alpha = C.c_double()
    t = np.zeros(8)
    rc = lib().orc_pipeline(img.ctypes.data, C.byref(P), s.ctypes.data, p, D.ctypes.data, C.byref(alpha),
                            mu.ctypes.data, z.ctypes.data, t.ctypes.data)
    if rc:
        raise RuntimeError(f"orc_pipeline failed rc={rc}")
    return dict(D=D, alpha=alpha.value, mu=mu, z=z[:, :, 0] if img.ndim == 2 else z, m=mm, p=p,
                timings=dict(zip(["affinity", "laplacian", "eigensolve", "nystroem", "gram_schmidt", "filter", "total"], t[:7])))


def affinity_rows(img, sample_indices, cols, kind="bilateral", h_loc=40.0, h_val=30.0):
    """K(samples, cols), p x len(cols) fp64 (OpenMP over samples)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape[:2]
    Cn = 1 if img.ndim == 2 else img.shape[2]
    s = np.ascontiguousarray(sample_indices, dtype=np.uint32)
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    P = Params(W, H, Cn, KINDS[kind], h_loc, h_val, 3.0, 1.0, -1, 0, 0, H)
    K = np.empty((len(s), len(cols)))
    lib().orc_affinity_rows(img.ctypes.data, C.byref(P), s.ctypes.data, len(s), cols.ctypes.data, len(cols), K.ctypes.data)
    return K


def num_threads():
    return lib().orc_num_threads()
