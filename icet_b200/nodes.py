"""Host-side mirror of the reference's two ROS nodes, minus ROS:

    OdometryNode   reference src/odometry.cpp:23-214       scan in -> pose / covariance diagonal / twist out
    ScanMatcherNode reference src/scanMatcher.cpp:18-150    scan in -> scan 2 in the frame of scan 1 + snail trail out
    MapMakerNode   reference src/simpleMapMaker.cpp:60-291  scan in -> pose + 600 000-point FIFO map out

Both are thin: every step between two scans (min-range filter, registration, seeding of the next registration, pose
accumulation, map re-expression) runs on the device behind the C ABI (include/icet_b200.h, "callers either side of
the path"); the host only hands scans in and reads the published values back.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api
from .api import Context, IcetError, OdometryParams, Params, Pose, POSE_DTYPE, RESULT_DTYPE, Result, as_planes


class Node:
    """icet_b200_node: device-resident prev_pcl_matrix / X0 / X_homo of one node."""

    def __init__(self, ctx: Context, params: Params, op: OdometryParams, max_points: int, X_homo0=None, X0=None):
        self.ctx, self._L = ctx, ctx._L
        self.params, self.op = params, op
        h = C.c_void_p()
        xh = None if X_homo0 is None else np.ascontiguousarray(X_homo0, np.float32).reshape(16)
        x0 = None if X0 is None else np.ascontiguousarray(X0, np.float32).reshape(6)
        self._h = None
        ctx._check(self._L.icet_b200_node_create(ctx._h, C.byref(params), C.byref(op), max_points,
                                                 None if xh is None else xh.ctypes.data,
                                                 None if x0 is None else x0.ctypes.data, C.byref(h)))
        self._h = h

    def close(self):
        if self._h is not None:
            self._L.icet_b200_node_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push(self, scan):
        """One scan from host memory (N x 3 or [3, N]).  Returns None for the very first scan, else (result, pose)
        as numpy records."""
        s = as_planes(scan)
        res, pose = Result(), Pose()
        n = s.shape[1]
        rc = self.ctx._check(self._L.icet_b200_node_push(self._h, s.ctypes.data, n, n, C.byref(res), C.byref(pose)))
        if rc == 0:
            return None
        return (np.frombuffer(bytes(res), dtype=RESULT_DTYPE)[0], np.frombuffer(bytes(pose), dtype=POSE_DTYPE)[0])

    def push_cloud(self, cloud, divide: float = 0.0):
        """`push` for a cloud in its native layout (see api.cloud_desc): e.g. the bytes of a PointCloud2 message."""
        c, keep = cloud if isinstance(cloud, tuple) else api.cloud_desc(cloud, divide)
        res, pose = Result(), Pose()
        rc = self.ctx._check(self._L.icet_b200_node_push_cloud(self._h, C.byref(c), C.byref(res), C.byref(pose)))
        if rc == 0:
            return None
        return (np.frombuffer(bytes(res), dtype=RESULT_DTYPE)[0], np.frombuffer(bytes(pose), dtype=POSE_DTYPE)[0])

    def push_device(self, scans_ptr: int, nscans: int, n: int, res_ptr: int, pose_ptr: int) -> int:
        """nscans consecutive scans in device memory [nscans, 3, n]; asynchronous.  Returns the number of
        registrations whose result / pose will be written to the device arrays."""
        return self.ctx._check(self._L.icet_b200_node_push_device(self._h, nscans, C.c_void_p(scans_ptr), n,
                                                                  C.c_void_p(res_ptr), C.c_void_p(pose_ptr)))

    def current_scan(self):
        """(device pointer of the filtered current scan planes, device pointer of its row count, leading dim)."""
        p, q, ld = C.c_void_p(), C.c_void_p(), C.c_int32()
        self.ctx._check(self._L.icet_b200_node_current_scan(self._h, C.byref(p), C.byref(q), C.byref(ld)))
        return p.value, q.value, ld.value

    def last_result_ptr(self):
        p = C.c_void_p()
        self.ctx._check(self._L.icet_b200_node_last_result(self._h, C.byref(p)))
        return p.value


class PointMap:
    """icet_b200_map: the reference's EigenQueue (src/simpleMapMaker.cpp:18-58) on the device."""

    def __init__(self, ctx: Context, capacity: int = 600_000):
        self.ctx, self._L, self.capacity = ctx, ctx._L, capacity
        h = C.c_void_p()
        self._h = None
        ctx._check(self._L.icet_b200_map_create(ctx._h, capacity, C.byref(h)))
        self._h = h

    def close(self):
        if self._h is not None:
            self._L.icet_b200_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_scan_device(self, scan_ptr: int, n: int, ld: int, X_ptr: int, idx=None, count: int | None = None,
                        n_dev_ptr: int = 0, guard_trans: float = 0.0, guard_rot: float = 0.0):
        ix = None if idx is None else np.ascontiguousarray(idx, np.int32)
        cnt = (len(ix) if ix is not None else n) if count is None else count
        self.ctx._check(self._L.icet_b200_map_add_scan_device(
            self._h, C.c_void_p(scan_ptr), n, ld, C.c_void_p(n_dev_ptr) if n_dev_ptr else None,
            None if ix is None else ix.ctypes.data, cnt, C.c_void_p(X_ptr), guard_trans, guard_rot))

    def get(self) -> np.ndarray:
        """EigenQueue::getQueue: rows oldest first, as an [n, 3] array."""
        out = np.zeros((3, self.capacity), np.float32)
        n = C.c_int32()
        self.ctx._check(self._L.icet_b200_map_get(self._h, out.ctypes.data, self.capacity, C.byref(n)))
        return np.ascontiguousarray(out[:, : n.value].T)


class OdometryNode:
    """Mirror of OdometryNode (src/odometry.cpp:23-214): `callback(scan)` is pointcloudCallback without the ROS
    plumbing.  Parameters are the node's constants: minD = 2 (:57), run_length 7, 24 x 75 bins (:73-75), X0 seeded with
    the previous solution (:82), 10 Hz twist (:134-139)."""

    def __init__(self, ctx: Context | None = None, max_points: int = 131072, min_range: float = 2.0, runlen: int = 7,
                 num_bins_phi: int = 24, num_bins_theta: int = 75, rate_hz: float = 10.0):
        self.ctx = ctx or api.default_context()
        self.node = Node(self.ctx, api.make_params(runlen, num_bins_phi, num_bins_theta),
                         OdometryParams(min_range, 1, rate_hz, 0.0, 0.0), max_points)
        self.X_homo = np.eye(4, dtype=np.float32)
        self.X0 = np.zeros(6, np.float32)
        self.frameCount = 0

    def callback(self, scan):
        self.frameCount += 1
        out = self.node.push(scan)
        if out is None:
            return None
        res, pose = out
        self.X0 = res["X"].copy()
        self.X_homo = pose["X_homo"].copy()
        return {"X": res["X"].copy(), "pred_stds": res["pred_stds"].copy(), "position": pose["position"].copy(),
                "orientation": pose["orientation"].copy(), "covariance_diag": pose["covariance_diag"].copy(),
                "twist": pose["twist"].copy(), "X_homo": pose["X_homo"].copy(), "n_points": int(pose["n_points"])}


class MapMakerNode:
    """Mirror of MapMakerNode (src/simpleMapMaker.cpp:60-291): min range 0.2 (:100), run_length 12 (:115), X0 = 0 for
    every pair (:124), divergence guard 0.3 / 0.3 (:128-137, :241-242), 2000-point random sample of every scan
    (:150-159) pushed into a 600 000-point FIFO that is re-expressed in the newest sensor frame (EigenQueue, :18-58).
    The sample is `rng.permutation(rows)[:downsample]` (the reference shuffles with a default-seeded std::mt19937;
    the C++ mirror in icet_b200/host/nodes.h uses exactly that)."""

    def __init__(self, ctx: Context | None = None, max_points: int = 131072, min_range: float = 0.2, runlen: int = 12,
                 num_bins_phi: int = 24, num_bins_theta: int = 75, capacity: int = 600_000, downsample: int = 2000,
                 trans_thresh: float = 0.3, rot_thresh: float = 0.3, seed: int = 5489):
        self.ctx = ctx or api.default_context()
        self.trans_thresh, self.rot_thresh, self.downsample = trans_thresh, rot_thresh, downsample
        self.node = Node(self.ctx, api.make_params(runlen, num_bins_phi, num_bins_theta),
                         OdometryParams(min_range, 0, 10.0, trans_thresh, rot_thresh), max_points)
        self.q = PointMap(self.ctx, capacity)
        self.rng = np.random.RandomState(seed)
        self.X_homo = np.eye(4, dtype=np.float32)
        self.frameCount = 0

    def callback(self, scan):
        self.frameCount += 1
        out = self.node.push(scan)
        if out is None:
            return None
        res, pose = out
        rows = int(pose["n_points"])
        idx = self.rng.permutation(rows)[: min(self.downsample, rows)].astype(np.int32)
        scan_ptr, n_ptr, ld = self.node.current_scan()
        self.q.add_scan_device(scan_ptr, rows, ld, self.node.last_result_ptr(), idx=idx, n_dev_ptr=n_ptr,
                               guard_trans=self.trans_thresh, guard_rot=self.rot_thresh)
        self.X_homo = pose["X_homo"].copy()
        return {"X": pose["X"].copy(), "pred_stds": res["pred_stds"].copy(), "X_homo": pose["X_homo"].copy(),
                "guarded": bool(pose["guarded"]), "n_points": rows, "sample": idx}

    def map_points(self) -> np.ndarray:
        return self.q.get()


class ScanMatcherNode:
    """Mirror of ScanMatcherNode (src/scanMatcher.cpp:18-150): every cloud is registered against its predecessor as it
    came (no range filter), X0 = 0 (:58-62), run_length 7, 24 x 75 bins (:55-57); publishes scan 2 re-expressed in
    the frame of scan 1, `(pcl_matrix * rot_mat.inverse()).rowwise() - trans` (:73), and the snail trail (:76-80).
    The re-expression of the cloud runs on the device from the device-resident result."""

    def __init__(self, ctx: Context | None = None, max_points: int = 131072, runlen: int = 7, num_bins_phi: int = 24,
                 num_bins_theta: int = 75):
        import torch
        self.ctx = ctx or api.default_context()
        self.node = Node(self.ctx, api.make_params(runlen, num_bins_phi, num_bins_theta),
                         OdometryParams(-1.0, 0, 10.0, 0.0, 0.0), max_points)
        self._torch = torch
        self._out = torch.empty((3, max_points), dtype=torch.float32, device="cuda")
        self.snailTrail = np.zeros((1, 3), np.float32)   # scanMatcher.cpp:25-26
        self.frameCount = 0

    def callback(self, scan):
        self.frameCount += 1
        s = as_planes(scan)
        if s.shape[1] == 0:           # "Received an empty point cloud" (:41-44)
            return None
        out = self.node.push(s)
        if out is None:               # "Previous point cloud is empty, skipping this frame" (:47-51)
            return None
        res, pose = out
        X = res["X"].copy()
        scan_ptr, n_ptr, ld = self.node.current_scan()
        n = s.shape[1]
        self.ctx.transform_cloud_device(scan_ptr, n, ld, self.node.last_result_ptr(), 0, self._out.data_ptr(),
                                        self._out.shape[1])
        self.ctx.synchronize()
        aligned = self._out[:, :n].cpu().numpy().T.copy()
        # snail trail (:76-80): a handful of rows, host side like the reference
        from numpy.linalg import inv
        R = _rot_R(X[3], X[4], X[5])
        self.snailTrail = ((self.snailTrail @ inv(R.astype(np.float64)).astype(np.float32)) - X[:3]).astype(np.float32)
        self.snailTrail = np.vstack([self.snailTrail, np.zeros((1, 3), np.float32)])
        return {"X": X, "scan2_in_scan1_frame": aligned, "snailTrail": self.snailTrail.copy()}


def _rot_R(phi, theta, psi) -> np.ndarray:
    """utils::R (reference src/utils.cpp:144-152), float32, row-major."""
    f = np.float32
    sph, cph, sth, cth, sps, cps = (np.sin(f(phi)), np.cos(f(phi)), np.sin(f(theta)), np.cos(f(theta)),
                                    np.sin(f(psi)), np.cos(f(psi)))
    return np.array([[cth * cps, sps * cph + sph * sth * cps, sph * sps - sth * cph * cps],
                     [-sps * cth, cph * cps - sph * sth * sps, sph * cps + sth * sps * cph],
                     [sth, -sph * cth, cph * cth]], dtype=f)
