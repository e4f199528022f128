"""TEST INFRASTRUCTURE -- CPU restatement of the reference's two ROS callbacks without ROS.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.
PARITY UNPINNED: the reference has no tests or golden outputs for these callbacks and cannot be built here (ROS,
PCL and Eigen are absent); every function cites the reference lines it restates.  The registration itself is
oracle/pyoracle.run (the restatement of src/icet.cpp).

float32 throughout, like the reference (Eigen::MatrixXf / Matrix4f / Quaternionf).
"""
from __future__ import annotations

import numpy as np

from . import pyoracle as po

F = np.float32


def rot_R(phi, theta, psi) -> np.ndarray:
    """utils::R, reference src/utils.cpp:144-152 (float trig on float angles), row-major 3 x 3."""
    phi, theta, psi = F(phi), F(theta), F(psi)
    sph, cph, sth, cth, sps, cps = np.sin(phi), np.cos(phi), np.sin(theta), np.cos(theta), np.sin(psi), np.cos(psi)
    return np.array([[cth * cps, sps * cph + sph * sth * cps, sph * sps - sth * cph * cps],
                     [-sps * cth, cph * cps - sph * sth * sps, sph * cps + sth * sps * cph],
                     [sth, -sph * cth, cph * cth]], dtype=F)


def min_range_filter(cloud: np.ndarray, min_d: float) -> np.ndarray:
    """odometry.cpp:57-70 / simpleMapMaker.cpp:100-112: rows with row.norm() > minD, order kept.  cloud: N x 3 f32."""
    c = np.asarray(cloud, F)
    d = np.sqrt((c[:, 0] * c[:, 0] + c[:, 1] * c[:, 1]) + c[:, 2] * c[:, 2], dtype=F)
    with np.errstate(invalid="ignore"):
        return c[d > F(min_d)]


def quaternion_from_rotation(m: np.ndarray) -> np.ndarray:
    """Eigen::Quaternionf(Matrix3f) (odometry.cpp:115-116): Eigen 3.3 Geometry/Quaternion.h,
    quaternionbase_assign_impl<Other,3,3> -- trace branch, else the largest diagonal entry.  Returns x, y, z, w."""
    m = np.asarray(m, F)
    q = np.zeros(4, F)
    t = F(m[0, 0] + m[1, 1] + m[2, 2])
    if t > 0:
        t = np.sqrt(F(t + F(1)))
        q[3] = F(0.5) * t
        t = F(0.5) / t
        q[0] = (m[2, 1] - m[1, 2]) * t
        q[1] = (m[0, 2] - m[2, 0]) * t
        q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(F(m[i, i] - m[j, j] - m[k, k] + F(1)))
        q[i] = F(0.5) * t
        t = F(0.5) / t
        q[3] = (m[k, j] - m[j, k]) * t
        q[j] = (m[j, i] + m[i, j]) * t
        q[k] = (m[k, i] + m[i, k]) * t
    return q


def homogeneous(X) -> np.ndarray:
    """X_homo_i of odometry.cpp:93-95: rotation block R(X3,X4,X5), translation column X0..2."""
    H = np.eye(4, dtype=F)
    H[:3, :3] = rot_R(X[3], X[4], X[5])
    H[:3, 3] = np.asarray(X[:3], F)
    return H


class OdometryOracle:
    """OdometryNode::pointcloudCallback, src/odometry.cpp:38-168 (ROS publishing removed)."""

    def __init__(self, min_d=2.0, runlen=7, bins_phi=24, bins_theta=75, chain=True, rate=10.0, guard=None, **kw):
        self.min_d, self.runlen, self.bins_phi, self.bins_theta = min_d, runlen, bins_phi, bins_theta
        self.chain, self.rate, self.guard, self.kw = chain, F(rate), guard, kw
        self.X0 = np.zeros(6, F)                     # :28-29
        self.X_homo = np.eye(4, dtype=F)             # :184
        self.prev = None
        self.frames = 0

    def callback(self, cloud):
        cloud = np.ascontiguousarray(np.asarray(cloud, F))
        if self.prev is None:                        # :47-52: the first cloud is stored as it came
            self.prev = cloud
            return None
        cur = min_range_filter(cloud, self.min_d)    # :57-70
        it = po.run(self.prev, cur, runlen=self.runlen, X0=self.X0, bins_phi=self.bins_phi,
                    bins_theta=self.bins_theta, dumps=None, **self.kw)   # :73-76
        X = np.asarray(it.X, F).copy()
        stds = np.asarray(it.pred_stds, F).copy()
        self.X0 = X.copy() if self.chain else np.zeros(6, F)            # :82 / simpleMapMaker.cpp:124
        guarded = False
        if self.guard is not None:                   # simpleMapMaker.cpp:128-137
            tt, rt = self.guard
            if np.any(np.abs(X[:3]) > F(tt)) or np.any(np.abs(X[3:]) > F(rt)):
                X = np.zeros(6, F)
                guarded = True
        self.prev = cur                              # :89
        self.X_homo = (self.X_homo @ homogeneous(X)).astype(F)         # :93-98
        self.frames += 1
        return {"X": X, "X_raw": np.asarray(it.X, F).copy(), "pred_stds": stds, "X_homo": self.X_homo.copy(),
                "position": self.X_homo[:3, 3].copy(),                 # :110-112
                "orientation": quaternion_from_rotation(self.X_homo[:3, :3]),   # :114-119
                "covariance_diag": stds.copy(),                        # :126-131
                "twist": (self.rate * X).astype(F),                    # :134-139
                "n_points": cur.shape[0], "guarded": guarded, "cur": cur}


class EigenQueueOracle:
    """class EigenQueue, src/simpleMapMaker.cpp:18-58."""

    def __init__(self, max_size=600_000):
        self.max_size = max_size
        self.matrix = np.zeros((max_size, 3), F)     # :21
        self.pos, self.filled = 0, False

    def enqueue(self, row):                          # :24-32
        self.matrix[self.pos] = row
        self.pos = (self.pos + 1) % self.max_size
        if self.pos == 0:
            self.filled = True

    def add_new_scan(self, new_scan, trans, rot_mat):  # :34-42
        for r in np.asarray(new_scan, F):
            self.enqueue(r)
        rinv = np.linalg.inv(rot_mat.astype(np.float64)).astype(F)     # rot_mat.inverse() (float in the reference)
        self.matrix = ((self.matrix - np.asarray(trans, F)) @ rinv).astype(F)   # :41

    def get_queue(self):                             # :44-51
        if not self.filled:
            return self.matrix[: self.pos].copy()
        return np.vstack([self.matrix[self.pos:], self.matrix[: self.pos]])


class MapMakerOracle:
    """MapMakerNode::pointcloudCallback, src/simpleMapMaker.cpp:78-240.  The 2000-row sample (:150-159) is passed
    in by the caller (the reference draws it with std::shuffle on a default-seeded std::mt19937)."""

    def __init__(self, max_size=600_000, min_d=0.2, runlen=12, trans_thresh=0.3, rot_thresh=0.3, **kw):
        self.odo = OdometryOracle(min_d=min_d, runlen=runlen, chain=False, guard=(trans_thresh, rot_thresh), **kw)
        self.q = EigenQueueOracle(max_size)

    def callback(self, cloud, sample_fn):
        out = self.odo.callback(cloud)
        if out is None:
            return None
        cur = out["cur"]
        idx = sample_fn(cur.shape[0])
        X = out["X"]
        self.q.add_new_scan(cur[idx], X[:3], rot_R(X[3], X[4], X[5]))   # :139-160
        out["sample"] = idx
        return out


class ScanMatcherOracle:
    """ScanMatcherNode::pointcloudCallback, src/scanMatcher.cpp:30-112 (ROS publishing removed)."""

    def __init__(self, runlen=7, bins_phi=24, bins_theta=75, **kw):
        self.runlen, self.bins_phi, self.bins_theta, self.kw = runlen, bins_phi, bins_theta, kw
        self.prev = None
        self.snail = np.zeros((1, 3), F)             # :25-26

    def callback(self, cloud):
        cloud = np.ascontiguousarray(np.asarray(cloud, F))
        if cloud.shape[0] == 0:                      # :41-44
            return None
        if self.prev is None:                        # :47-51
            self.prev = cloud
            return None
        it = po.run(self.prev, cloud, runlen=self.runlen, X0=np.zeros(6, F), bins_phi=self.bins_phi,
                    bins_theta=self.bins_theta, dumps=None, **self.kw)   # :55-63
        X = np.asarray(it.X, F).copy()
        self.prev = cloud                            # :68
        rinv = np.linalg.inv(rot_R(X[3], X[4], X[5]).astype(np.float64)).astype(F)
        aligned = ((cloud @ rinv) - X[:3]).astype(F)                     # :73
        self.snail = np.vstack([((self.snail @ rinv) - X[:3]).astype(F), np.zeros((1, 3), F)])   # :76-80
        return {"X": X, "scan2_in_scan1_frame": aligned, "snailTrail": self.snail.copy()}
