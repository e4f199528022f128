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
