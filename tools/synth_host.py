"""Host build of the synthetic scan generator (bench / test tooling)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsynth_host.so")
_SRC = [os.path.join(_HERE, "synth_host.cpp"), os.path.join(_HERE, "..", "icet_b200", "csrc", "synth.h")]
_LIB = None


def build():
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(map(os.path.getmtime, _SRC)):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.run([os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", _SO,
                        _SRC[0]], check=True)
    return _SO


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.synth_host_scans.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _LIB.synth_host_motion.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
        _LIB.synth_host_motions.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        _LIB.synth_host_map.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_long]
        _LIB.synth_host_map.restype = C.c_long
    return _LIB


def scans(nscans, first_scan=0, seed=20240, rings=64, azim=2048, nthreads=None) -> np.ndarray:
    """float32 [nscans, 3, rings*azim]"""
    out = np.zeros((nscans, 3, rings * azim), np.float32)
    nthreads = nthreads or os.cpu_count() or 1
    rc = _lib().synth_host_scans(seed, first_scan, nscans, rings, azim, nthreads, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("synth_host_scans failed")
    return out


def motion(k, seed=20240) -> np.ndarray:
    d = np.zeros(6, np.float64)
    _lib().synth_host_motion(seed, k, d.ctypes.data)
    return d


def motions(k0, n, seed=20240) -> np.ndarray:
    """[n, 6] generator motions (dx dy dz roll pitch yaw) of the steps k0 .. k0+n-1: the ground truth of the pairs."""
    d = np.zeros((n, 6), np.float64)
    _lib().synth_host_motions(seed, k0, n, d.ctypes.data)
    return d


def submap(nscans=1240, per_scan=2000, first_scan=20, seed=20240, rings=64, azim=2048, nthreads=None):
    """BASELINE.json configs[4]: (map [3, M] float32, current scan [3, rings*azim]).  The map holds `per_scan` returns of
    each of `nscans` consecutive scans (the simpleMapMaker.cpp:150-160 recipe) expressed in the frame of the scan that
    follows them, which is returned as the scan to match.  1240 x 2000 rays minus the dropped returns and the rays without a hit = 2.0 M points.
    Bit-reproducible (pure function of the arguments, double-precision poses of the generator)."""
    cap = nscans * per_scan
    buf = np.zeros((3, cap), np.float32)
    nthreads = nthreads or os.cpu_count() or 1
    n = _lib().synth_host_map(seed, first_scan, nscans, per_scan, rings, azim, nthreads, buf.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("synth_host_map failed: %d" % n)
    cur = scans(1, first_scan=first_scan + nscans, seed=seed, rings=rings, azim=azim, nthreads=nthreads)[0]
    return np.ascontiguousarray(buf[:, :n]), cur
