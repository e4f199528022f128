"""CPU tests of the callers-layer oracle (oracle/nodes_oracle.py: the reference's ROS callbacks restated in numpy)
against independent implementations (scipy / numpy) and the reference's documented behaviour, and of the C++ host
mirror (include/icet_nodes.h) building and failing loudly without a GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def test_rot_R_is_the_reference_convention():
    """utils::R (src/utils.cpp:144-152) = Rz(psi)^T-like product used as p' = (p + t) * R: orthonormal, det 1, and
    equal to the matrix the registration oracle applies (checked through a tiny registration-free identity)."""
    from oracle import nodes_oracle as no
    R = no.rot_R(0.1, -0.2, 0.3).astype(np.float64)
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-6)
    assert abs(np.linalg.det(R) - 1) < 1e-6
    # closed form at single angles: R(0,0,psi) rotates row vectors by -psi about z
    Rz = no.rot_R(0, 0, 0.5)
    np.testing.assert_allclose(Rz, [[np.cos(.5), np.sin(.5), 0], [-np.sin(.5), np.cos(.5), 0], [0, 0, 1]], atol=1e-7)
    Ry = no.rot_R(0, 0.5, 0)
    np.testing.assert_allclose(Ry, [[np.cos(.5), 0, -np.sin(.5)], [0, 1, 0], [np.sin(.5), 0, np.cos(.5)]], atol=1e-7)


def test_quaternion_matches_scipy_on_both_branches():
    from scipy.spatial.transform import Rotation
    from oracle import nodes_oracle as no
    rng = np.random.RandomState(0)
    for k in range(200):
        m = Rotation.from_rotvec(rng.randn(3) * (3.1 if k % 2 else 0.3)).as_matrix().astype(np.float32)
        q = no.quaternion_from_rotation(m)
        qs = Rotation.from_matrix(m.astype(np.float64)).as_quat()
        assert min(np.abs(q - qs).max(), np.abs(q + qs).max()) < 2e-6
    # trace <= 0 branch explicitly (rotation by pi about y)
    q = no.quaternion_from_rotation(np.diag([-1.0, 1.0, -1.0]))
    np.testing.assert_allclose(np.abs(q), [0, 1, 0, 0], atol=1e-7)


def test_min_range_filter_keeps_order_and_drops_zero_and_nan_rows():
    from oracle import nodes_oracle as no
    c = np.array([[3, 0, 0], [0, 0, 0], [1, 1, 1], [np.nan, 0, 0], [0, 2.0000002, 0], [0, 2, 0], [0, -5, 1]], np.float32)
    np.testing.assert_array_equal(no.min_range_filter(c, 2.0), c[[0, 4, 6]])
    np.testing.assert_array_equal(no.min_range_filter(c, 0.0), c[[0, 2, 4, 5, 6]])


def test_eigen_queue_fifo_and_reexpression():
    """EigenQueue (simpleMapMaker.cpp:18-58): ring order oldest-first after wrapping; every insertion re-expresses
    ALL stored rows, the new ones included; composing the re-expressions equals one rigid transform."""
    from oracle import nodes_oracle as no
    q = no.EigenQueueOracle(5)
    a = np.arange(9, dtype=np.float32).reshape(3, 3)
    q.add_new_scan(a, [0, 0, 0], np.eye(3, dtype=np.float32))
    np.testing.assert_array_equal(q.get_queue(), a)
    b = 100 + np.arange(12, dtype=np.float32).reshape(4, 3)
    t = np.array([1, 2, 3], np.float32)
    q.add_new_scan(b, t, np.eye(3, dtype=np.float32))
    assert q.filled and q.pos == 2
    # 7 rows through a 5-ring: a[0], a[1] were overwritten by b[2], b[3]; oldest first = a[2], b[0..3]
    np.testing.assert_array_equal(q.get_queue(), np.vstack([a[2:], b]) - t)
    R = no.rot_R(0.02, -0.01, 0.3)
    before = q.get_queue()
    q.add_new_scan(np.zeros((0, 3), np.float32), [0.5, 0, 0], R)
    np.testing.assert_allclose(q.get_queue(), (before - np.float32([0.5, 0, 0])) @ np.linalg.inv(R), atol=1e-4)


def test_odometry_oracle_chains_and_accumulates(po, frame_pair):
    """OdometryNode (odometry.cpp): first cloud unfiltered, X0 <- X, X_homo = prod X_homo_i.  Small sub-sampled
    bundled scans keep this fast."""
    from oracle import nodes_oracle as no
    s1, s2 = frame_pair
    a, b = np.ascontiguousarray(s1[:, ::4].T), np.ascontiguousarray(s2[:, ::4].T)
    o = no.OdometryOracle(runlen=3, min_d=2.0)
    assert o.callback(a) is None and o.prev.shape == a.shape          # unfiltered
    r1 = o.callback(b)
    assert r1["n_points"] == int((np.linalg.norm(b, axis=1) > 2.0).sum()) < b.shape[0]
    np.testing.assert_array_equal(o.X0, r1["X"])
    np.testing.assert_allclose(r1["X_homo"], no.homogeneous(r1["X"]), atol=1e-7)
    r2 = o.callback(a)                                                # seeded with r1's solution
    ref = po.run(r1["cur"], no.min_range_filter(a, 2.0), runlen=3, X0=r1["X"], dumps=None)
    np.testing.assert_array_equal(r2["X"], np.asarray(ref.X, np.float32))
    np.testing.assert_allclose(r2["X_homo"], no.homogeneous(r1["X"]) @ no.homogeneous(r2["X"]), atol=1e-6)
    np.testing.assert_allclose(r2["twist"], 10 * r2["X"], rtol=1e-6)
    m = no.MapMakerOracle(max_size=1000, runlen=2, trans_thresh=1e-9)
    m.callback(a, None)
    g = m.callback(b, lambda rows: np.arange(min(rows, 300)))
    assert g["guarded"] and not g["X"].any() and np.array_equal(m.odo.X0, np.zeros(6, np.float32))
    np.testing.assert_array_equal(m.q.get_queue(), g["cur"][:300])    # identity re-expression


def test_cpp_nodes_compile_and_fail_loudly_without_gpu(tmp_path):
    """include/icet_nodes.h + icet_b200/host/nodes.cpp build against the test-only Eigen stub; without a GPU the
    node constructors throw instead of falling back to a CPU path."""
    import icet_b200
    icet_b200.build()
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present (the GPU test runs the nodes)")
    f = tmp_path / "seq.f32"
    np.zeros((2, 3, 64), np.float32).tofile(f)
    for mode in ("odometry", "map"):
        r = subprocess.run([os.path.join(ROOT, "examples", "_build", "nodes_headless"), mode, "64", str(f)],
                           capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_scan_matcher_oracle(po, frame_pair):
    """ScanMatcherNode (scanMatcher.cpp:30-112): empty cloud and first cloud produce nothing; then X0 = 0 registration
    of unfiltered clouds, scan 2 re-expressed as (cloud * R^-1) - t, snail trail one row longer per registration."""
    from oracle import nodes_oracle as no
    s1, s2 = frame_pair
    a, b = np.ascontiguousarray(s1[:, ::4].T), np.ascontiguousarray(s2[:, ::4].T)
    o = no.ScanMatcherOracle(runlen=3)
    assert o.callback(np.zeros((0, 3), np.float32)) is None
    assert o.callback(a) is None
    r = o.callback(b)
    ref = po.run(a, b, runlen=3, dumps=None)
    np.testing.assert_array_equal(r["X"], np.asarray(ref.X, np.float32))
    R = no.rot_R(*r["X"][3:]).astype(np.float64)
    np.testing.assert_allclose(r["scan2_in_scan1_frame"], b.astype(np.float64) @ np.linalg.inv(R) - r["X"][:3], atol=2e-5)
    assert r["snailTrail"].shape == (2, 3) and not r["snailTrail"][-1].any()
    np.testing.assert_allclose(r["snailTrail"][0], -r["X"][:3], atol=1e-7)
    r2 = o.callback(a)
    assert r2["snailTrail"].shape == (3, 3)
