"""ctypes binding of oracle/_ref/libicet_ref.so: the REFERENCE's own sources (src/icet.cpp, src/utils.cpp,
src/ThreadPool.cpp under /root/reference) compiled by `make -C oracle ref` around oracle/ref/ref_driver.cpp.

TEST INFRASTRUCTURE ONLY (tests/, tools/pin_against_ref.py, bench.py's CPU legs).  /root/reference exists only in the
build container: on the GPU box the prebuilt oracle/_ref/libicet_ref.so (git-ignored, but it travels with the
snapshot) is used as it is.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("ICET_REFERENCE", "/root/reference")
_FP, _IP, _BP = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)

_FIELDS = [("clusterBounds", _FP, np.float32, lambda d: (d["ncell"], 6)),
           ("cnt1", _IP, np.int32, lambda d: (d["ncell"],)),
           ("cnt2", _IP, np.int32, lambda d: (d["ncell"],)),
           ("has1", _BP, np.uint8, lambda d: (d["ncell"],)),
           ("mu1", _FP, np.float32, lambda d: (d["ncell"], 3)),
           ("sigma1", _FP, np.float32, lambda d: (d["ncell"], 3, 3)),
           ("U", _FP, np.float32, lambda d: (d["ncell"], 3, 3)),
           ("L", _FP, np.float32, lambda d: (d["ncell"], 3, 3)),
           ("points2", _FP, np.float32, lambda d: (3, d["n2"])),
           ("HTWH", _FP, np.float32, lambda d: (6, 6)),
           ("HTWdz", _FP, np.float32, lambda d: (6,)),
           ("testPoints", _FP, np.float32, lambda d: (d["ncell"] * 6, 3))]


class _Out(C.Structure):
    _fields_ = [("X", C.c_float * 6), ("pred_stds", C.c_float * 6), ("n_ellipsoids", C.c_int32)] + \
               [(n, t) for n, t, _, _ in _FIELDS]


def so_path(native: bool = False) -> str:
    return os.path.join(_HERE, "_ref", "libicet_ref_native.so" if native else "libicet_ref.so")


def available(native: bool = False) -> bool:
    """True if the library exists or can be built here (the reference's sources are present)."""
    return os.path.exists(so_path(native)) or os.path.isdir(os.path.join(REFERENCE, "src"))


def build(native: bool = False, eigen_include: str | None = None) -> str:
    """`make -C oracle ref` when the reference's sources are present; otherwise the prebuilt library must exist."""
    target = so_path(native)
    have_src = os.path.isdir(os.path.join(REFERENCE, "src"))
    if have_src:
        deps = [os.path.join(_HERE, f) for f in ("ref/ref_driver.cpp", "eigen_algos.h", "eigen_shim/Eigen/Dense", "Makefile")]
        deps += [os.path.join(REFERENCE, "src", f) for f in ("icet.cpp", "utils.cpp", "ThreadPool.cpp")]
        if (not os.path.exists(target)) or os.path.getmtime(target) < max(map(os.path.getmtime, deps)):
            cmd = ["make", "-C", _HERE, "ref", "REFERENCE=" + REFERENCE]
            if eigen_include:
                cmd.append("EIGEN_INCLUDE=" + eigen_include)
            if native:
                # the reference's CMakeLists.txt:18,38 says -O3 -march=native; its sources exist only in the build
                # container, so the timing build targets the x86-64-v3 level (AVX2 + FMA) every GPU-box host has
                cmd += ["REF_OPT=-O3 -march=x86-64-v3", "REF_LIB=libicet_ref_native.so"]
            subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    if not os.path.exists(target):
        raise RuntimeError("%s is missing and %s is not available to build it" % (target, REFERENCE))
    return target


_LIBS: dict = {}


def lib(native: bool = False) -> C.CDLL:
    if native not in _LIBS:
        L = C.CDLL(build(native))
        L.icet_ref_run.restype = C.c_int
        L.icet_ref_run.argtypes = [_FP, C.c_int32, _FP, C.c_int32, C.c_int32, _FP, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_float, C.c_float, C.POINTER(_Out)]
        L.icet_ref_run_sequence.restype = C.c_double
        L.icet_ref_run_sequence.argtypes = [_FP, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_float, C.c_float, _FP]
        _LIBS[native] = L
    return _LIBS[native]


class RefResult(dict):
    __getattr__ = dict.__getitem__


def run(scan1, scan2, runlen=7, X0=None, bins_phi=24, bins_theta=75, n=25, thresh=0.1, buff=0.1) -> RefResult:
    """`ICET it(scan1, scan2, runlen, X0, num_bins_phi, num_bins_theta, n, thresh, buff)` of the reference
    (include/icet.h:38-40); returns its public members.  scan1 / scan2: float32 [3, N] planes."""
    s1 = np.ascontiguousarray(scan1, np.float32)
    s2 = np.ascontiguousarray(scan2, np.float32)
    assert s1.ndim == 2 and s1.shape[0] == 3 and s2.ndim == 2 and s2.shape[0] == 3
    x0 = np.zeros(6, np.float32) if X0 is None else np.ascontiguousarray(X0, np.float32)
    dims = dict(ncell=bins_phi * bins_theta, n2=s2.shape[1])
    o = _Out()
    arrays = {}
    for name, ct, dt, shp in _FIELDS:
        arrays[name] = np.zeros(shp(dims), dt)
        setattr(o, name, arrays[name].ctypes.data_as(ct))
    rc = lib().icet_ref_run(s1.ctypes.data_as(_FP), s1.shape[1], s2.ctypes.data_as(_FP), s2.shape[1], runlen,
                            x0.ctypes.data_as(_FP), bins_phi, bins_theta, n, thresh, buff, C.byref(o))
    if rc != 0:
        raise RuntimeError("the reference constructor threw")
    return RefResult(X=np.array(o.X[:], np.float32), pred_stds=np.array(o.pred_stds[:], np.float32),
                     n_ellipsoids=int(o.n_ellipsoids), **arrays)


def run_sequence(scans: np.ndarray, nthreads=1, native=True, runlen=7, bins_phi=24, bins_theta=75, n=25, thresh=0.1,
                 buff=0.1):
    """scans: float32 [K+1, 3, N]: the K consecutive pairs with X0 = 0, one ICET object per pair, on `nthreads` host
    threads.  Returns (results [K, 12] = X | pred_stds, elapsed seconds)."""
    scans = np.ascontiguousarray(scans, np.float32)
    k1, three, npts = scans.shape
    assert three == 3 and k1 >= 2
    res = np.zeros((k1 - 1, 12), np.float32)
    dt = lib(native).icet_ref_run_sequence(scans.ctypes.data_as(_FP), npts, k1 - 1, nthreads, runlen, bins_phi,
                                           bins_theta, n, thresh, buff, res.ctypes.data_as(_FP))
    return res, float(dt)
