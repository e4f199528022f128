"""GPU parity tests of the callers either side of the path: the device-resident odometry and
map-maker callbacks, through the C ABI, against the numpy restatement of the reference's ROS callbacks
(oracle/nodes_oracle.py) on identical inputs.

Tolerances: row selection of the min-range filter and all counts exact; registration results as in
test_gpu_parity.py (1e-4 m, 1e-5 rad); accumulated pose within the sum of the per-step tolerances; map points
within 1e-4 m of the oracle's (they inherit the registration tolerance through the re-expression)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_M, TOL_RAD = 1e-4, 1e-5


@pytest.fixture(scope="module")
def scans():
    """seven consecutive synthetic 64-channel scans as N x 3 arrays (rows = points, like Eigen::MatrixXf)"""
    from tools import synth_host
    s = synth_host.scans(7, first_scan=3, seed=20240, rings=64, azim=2048)
    return [np.ascontiguousarray(a.T) for a in s]


def test_min_range_filter_is_exact_and_stable(ctx, scans):
    """odometry.cpp:57-70: the filtered cloud is the same rows in the same order (bitwise), for radii that cut into
    the scan; dropped returns (norm 0) always go."""
    from icet_b200 import Node, api
    from oracle import nodes_oracle as no
    for min_d in (2.0, 6.5, 0.0):
        nd = Node(ctx, api.make_params(runlen=0), api.OdometryParams(min_d, 0, 10.0, 0.0, 0.0), 131072)
        assert nd.push(scans[0]) is None
        res, pose = nd.push(scans[1])
        ref = no.min_range_filter(scans[1], min_d)
        assert int(pose["n_points"]) == ref.shape[0] < scans[1].shape[0]
        # read the device-resident filtered cloud back through a map with the identity re-expression (X = 0)
        import torch
        from icet_b200 import PointMap
        p, q, ld = nd.current_scan()
        zero = torch.zeros(6, dtype=torch.float32, device="cuda")
        pm = PointMap(ctx, 140000)
        pm.add_scan_device(p, ld, ld, zero.data_ptr(), n_dev_ptr=q, count=ld)
        got = pm.get()
        pm.close()
        assert got.shape == ref.shape
        assert got.tobytes() == ref.tobytes()
        nd.close()


def test_odometry_node_matches_oracle(ctx, scans):
    """OdometryNode::pointcloudCallback (odometry.cpp:38-168): chained registrations, accumulated pose, quaternion,
    covariance diagonal, twist."""
    from icet_b200 import OdometryNode
    from oracle import nodes_oracle as no
    node = OdometryNode(ctx)
    orc = no.OdometryOracle()
    for k, s in enumerate(scans[:5]):
        g, o = node.callback(s), orc.callback(s)
        if k == 0:
            assert g is None and o is None
            continue
        assert g["n_points"] == o["n_points"]
        assert np.abs(g["X"][:3] - o["X"][:3]).max() < TOL_M * k      # the seed carries the previous difference
        assert np.abs(g["X"][3:] - o["X"][3:]).max() < TOL_RAD * k
        assert np.abs(g["X_homo"] - o["X_homo"]).max() < 2 * TOL_M * k
        assert np.abs(g["position"] - o["position"]).max() < 2 * TOL_M * k
        q, qo = g["orientation"], o["orientation"]
        assert abs(np.linalg.norm(q) - 1) < 1e-5 and min(np.abs(q - qo).max(), np.abs(q + qo).max()) < 1e-5 * k
        np.testing.assert_allclose(g["twist"], 10.0 * g["X"], rtol=1e-6)
        np.testing.assert_allclose(g["covariance_diag"], g["pred_stds"])
        # (fp32 oracle without its double twin here: its small rotational entries carry ~1e-3 of fp32 COD noise, see
        # check_final in test_gpu_parity.py for the rule with the twin)
        np.testing.assert_allclose(g["pred_stds"], o["pred_stds"], rtol=5e-3)
    # chaining is what the node does: the last registration started from the previous solution
    assert np.abs(node.X0 - orc.X0).max() < 1e-3


def test_batched_callbacks_equal_streamed_callbacks(ctx, scans):
    """icet_b200_node_push_device with all scans at once (ONE persistent kernel in chain order) == one blocking
    callback per scan, byte for byte; also for the unchained (map-maker) flavour, which batches independent pairs."""
    import torch
    from icet_b200 import Node, api
    planes = np.stack([np.ascontiguousarray(s.T) for s in scans])
    dev = torch.from_numpy(planes).cuda()
    ns, n = planes.shape[0], planes.shape[2]
    for chain, runlen in ((1, 7), (0, 4)):
        op = api.OdometryParams(2.0, chain, 10.0, 0.3, 0.3)
        a = Node(ctx, api.make_params(runlen=runlen), op, n)
        streamed = [a.push(s) for s in scans]
        assert streamed[0] is None
        b = Node(ctx, api.make_params(runlen=runlen), op, n)
        res = torch.zeros((ns, 56), dtype=torch.float32, device="cuda")
        pose = torch.zeros((ns, 45), dtype=torch.float32, device="cuda")
        # in two calls: the node state (prev scan, seed, pose) carries over
        k1 = b.push_device(dev.data_ptr(), 3, n, res.data_ptr(), pose.data_ptr())
        k2 = b.push_device(dev[3:].data_ptr(), ns - 3, n, res[k1:].data_ptr(), pose[k1:].data_ptr())
        ctx.synchronize()
        assert (k1, k2) == (2, ns - 3)
        r = res.cpu().numpy().view(api.RESULT_DTYPE).reshape(-1)
        p = pose.cpu().numpy().view(api.POSE_DTYPE).reshape(-1)
        for k in range(1, ns):
            sr, sp = streamed[k]
            assert r[k - 1]["X"].tobytes() == sr["X"].tobytes(), (chain, k)
            assert r[k - 1]["Q"].tobytes() == sr["Q"].tobytes()
            assert p[k - 1]["X_homo"].tobytes() == sp["X_homo"].tobytes()
            assert p[k - 1]["orientation"].tobytes() == sp["orientation"].tobytes()
            assert int(p[k - 1]["n_points"]) == int(sp["n_points"]) and int(p[k - 1]["frame"]) == k
        a.close()
        b.close()


def test_map_maker_node_matches_oracle(ctx, scans):
    """MapMakerNode::pointcloudCallback + EigenQueue (simpleMapMaker.cpp:18-58, :78-240) with a small ring so that
    the FIFO wraps: same sample rows, same queue order, points within the registration tolerance."""
    from icet_b200 import MapMakerNode
    from oracle import nodes_oracle as no
    cap, ds = 5000, 2000
    node = MapMakerNode(ctx, capacity=cap, downsample=ds, runlen=5)
    orc = no.MapMakerOracle(max_size=cap, runlen=5)
    for k, s in enumerate(scans[:5]):
        g = node.callback(s)
        o = orc.callback(s, (lambda rows: g["sample"]) if g is not None else None)
        if k == 0:
            assert g is None and o is None
            continue
        assert g["n_points"] == o["n_points"] and g["guarded"] == o["guarded"]
        assert np.abs(g["X"][:3] - o["X"][:3]).max() < TOL_M and np.abs(g["X"][3:] - o["X"][3:]).max() < TOL_RAD
        m, mo = node.map_points(), orc.q.get_queue()
        assert m.shape == mo.shape == (min(cap, ds * k), 3)
        # points up to ~100 m away, re-expressed k times with transforms that agree to 1e-4 m / 1e-5 rad
        assert np.abs(m - mo).max() < k * (TOL_M + 120.0 * TOL_RAD)
    assert orc.q.filled


def test_divergence_guard(ctx, scans):
    """simpleMapMaker.cpp:128-137: a solution beyond the thresholds is replaced by zero for the pose and the map."""
    from icet_b200 import MapMakerNode
    node = MapMakerNode(ctx, capacity=4000, downsample=1000, runlen=5, trans_thresh=0.05)
    node.callback(scans[0])
    g = node.callback(scans[1])           # the synthetic vehicle moves 0.2 .. 0.8 m per scan
    assert g["guarded"] and not g["X"].any()
    np.testing.assert_array_equal(g["X_homo"], np.eye(4, dtype=np.float32))
    m = node.map_points()
    cur = scans[1][np.linalg.norm(scans[1].astype(np.float32), axis=1) > 0.2]
    np.testing.assert_array_equal(m, cur[g["sample"]])   # X = 0: the re-expression is the identity


def test_cpp_node_classes(ctx, scans, tmp_path):
    """The C++ mirrors (include/icet_nodes.h, examples/nodes_headless.cpp) replay the same scans: same library, so
    the same bits as the Python mirrors; the map holds min(capacity, k * downsample) rows."""
    import json
    import os
    import subprocess
    from conftest import ROOT
    from icet_b200 import MapMakerNode, OdometryNode
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "all"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exe = os.path.join(ROOT, "examples", "_build", "nodes_headless")
    f = tmp_path / "seq.f32"
    use = scans[:4]
    np.stack([np.ascontiguousarray(s.T) for s in use]).tofile(f)
    n = use[0].shape[0]

    def poses(out):
        rows = []
        for l in out.splitlines():
            if l.startswith("POSE"):
                x = json.loads(l[l.index("X [") + 2: l.index("] pos") + 1])
                p = json.loads(l[l.index("pos [") + 4: l.index("] quat") + 1])
                q = json.loads(l[l.index("quat [") + 5: l.index("] points") + 1])
                rows.append((np.float32(x), np.float32(p), np.float32(q), int(l.split("points ")[1].split()[0])))
        return rows

    out = subprocess.run([exe, "odometry", str(n), str(f)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    node = OdometryNode(ctx)
    py = [node.callback(s) for s in use][1:]
    cpp = poses(out.stdout)
    assert len(cpp) == len(py) == 3
    for (x, p, q, npts), g in zip(cpp, py):
        np.testing.assert_array_equal(x, g["X"])
        np.testing.assert_array_equal(p, g["position"])
        np.testing.assert_array_equal(q, g["orientation"])
        assert npts == g["n_points"]

    out = subprocess.run([exe, "map", str(n), str(f), "5000", "2000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    mm = MapMakerNode(ctx, capacity=5000, downsample=2000)
    py = [mm.callback(s) for s in use][1:]
    cpp = poses(out.stdout)
    for (x, p, q, npts), g in zip(cpp, py):
        np.testing.assert_array_equal(x, g["X"])
        np.testing.assert_array_equal(p, g["X_homo"][:3, 3])
    ml = [l for l in out.stdout.splitlines() if l.startswith("MAP")][0]
    assert "rows 5000" in ml

    from icet_b200 import ScanMatcherNode
    out = subprocess.run([exe, "match", str(n), str(f)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    sm = ScanMatcherNode(ctx)
    py = [sm.callback(s) for s in use][1:]
    cpp = poses(out.stdout)
    assert len(cpp) == len(py) == 3
    for (x, p, q, npts), g in zip(cpp, py):
        np.testing.assert_array_equal(x, g["X"])
    al = [l for l in out.stdout.splitlines() if l.startswith("ALIGNED")][0]
    first = np.float32(json.loads(al[al.index("first [") + 6: al.index("] trail") + 1]))
    np.testing.assert_array_equal(first, py[-1]["scan2_in_scan1_frame"][0])
    assert "rows %d" % n in al and al.endswith("trail 4")


def test_ingest_native_layouts(ctx, frame_pair):
    """clouds in the layouts the reference's callers hold them in -- float64 .npy arrays in C and
    Fortran order (src/sample_data, python/point_clouds), PointCloud2 records (pcl::PointXYZ: 16-byte step),
    integer millimetres (Ouster CSV, src/utils.cpp:42-52) -- converted on the device: bit-identical to converting on
    the host the way the reference does (cast<float>(), `/ 1000`) and registering the planes."""
    import torch
    from icet_b200 import Node, api
    s1, s2 = frame_pair                      # float32 [3, N] planes
    ref = ctx.register(s1, s2)
    a64, b64 = np.ascontiguousarray(s1.T.astype(np.float64)), np.ascontiguousarray(s2.T.astype(np.float64))
    for a, b in ((a64, b64), (np.asfortranarray(a64), np.asfortranarray(b64)),
                 (np.ascontiguousarray(s1.T), np.ascontiguousarray(s2.T))):
        r = ctx.register_clouds(a, b)
        assert r["X"].tobytes() == ref["X"].tobytes() and r["Q"].tobytes() == ref["Q"].tobytes()
    # PointCloud2-style records: x, y, z, padding (pcl::PointXYZ), plus an intensity-first variant with offsets 4, 8, 12
    n = s1.shape[1]
    rec = np.zeros((n, 4), np.float32); rec[:, :3] = s1.T; rec[:, 3] = 7.0
    rec2 = np.zeros((n, 5), np.float32); rec2[:, 1:4] = s2.T; rec2[:, 0] = 3.0
    r = ctx.register_clouds(api.cloud_desc(rec.tobytes(), point_step=16, offsets=(0, 4, 8)),
                            api.cloud_desc(rec2.tobytes(), point_step=20, offsets=(4, 8, 12)))
    assert r["X"].tobytes() == ref["X"].tobytes()
    # integer millimetres / 1000 (Ouster CSV path): the host reference conversion is float(int) / 1000.0f
    mm1, mm2 = np.round(s1.T * 1000).astype(np.int32), np.round(s2.T * 1000).astype(np.int32)
    h1 = (mm1.astype(np.float32) / np.float32(1000)).T.copy()
    h2 = (mm2.astype(np.float32) / np.float32(1000)).T.copy()
    rh = ctx.register(h1, h2)
    r = ctx.register_clouds(mm1, mm2, divide=1000.0)
    assert r["X"].tobytes() == rh["X"].tobytes() and r["pred_stds"].tobytes() == rh["pred_stds"].tobytes()
    # plain conversion check, including double -> float rounding and a leading dimension
    out = torch.zeros((3, n + 8), dtype=torch.float32, device="cuda")
    v = (a64 * (1 + 1e-9)).copy()
    ctx.ingest(v, out.data_ptr(), n + 8)
    ctx.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy()[:, :n], v.astype(np.float32).T)
    # the node callback straight from records
    op = api.OdometryParams(2.0, 1, 10.0, 0.0, 0.0)
    na, nb = Node(ctx, api.make_params(), op, n), Node(ctx, api.make_params(), op, n)
    assert na.push(s1) is None and nb.push_cloud(api.cloud_desc(rec.tobytes(), point_step=16, offsets=(0, 4, 8))) is None
    ra, rb = na.push(s2), nb.push_cloud(api.cloud_desc(rec2.tobytes(), point_step=20, offsets=(4, 8, 12)))
    assert ra[0]["X"].tobytes() == rb[0]["X"].tobytes() and ra[1]["X_homo"].tobytes() == rb[1]["X_homo"].tobytes()
    with pytest.raises(Exception):
        ctx.register_clouds(api.cloud_desc(rec.tobytes(), point_step=16, offsets=(0, 4, 14)), a64)


def test_scan_matcher_node_matches_oracle(ctx, scans):
    """ScanMatcherNode (src/scanMatcher.cpp:30-112): unfiltered consecutive pairs, X0 = 0, scan 2 re-expressed in the
    frame of scan 1 on the device, snail trail."""
    from icet_b200 import ScanMatcherNode
    from oracle import nodes_oracle as no
    node, orc = ScanMatcherNode(ctx), no.ScanMatcherOracle()
    assert node.callback(np.zeros((0, 3), np.float32)) is None and orc.callback(np.zeros((0, 3), np.float32)) is None
    for k, s in enumerate(scans[:3]):
        g, o = node.callback(s), orc.callback(s)
        if k == 0:
            assert g is None and o is None
            continue
        assert np.abs(g["X"][:3] - o["X"][:3]).max() < TOL_M and np.abs(g["X"][3:] - o["X"][3:]).max() < TOL_RAD
        a, b = g["scan2_in_scan1_frame"], o["scan2_in_scan1_frame"]
        assert a.shape == b.shape == s.shape
        assert np.abs(a - b).max() < TOL_M + 120.0 * TOL_RAD
        # the re-expression itself, from the GPU's own X: exact to rounding
        Rinv = np.linalg.inv(no.rot_R(*g["X"][3:]).astype(np.float64))
        ref = s.astype(np.float64) @ Rinv - g["X"][:3].astype(np.float64)
        assert np.abs(a - ref).max() < 6e-5   # fp32 evaluation on coordinates up to 120 m (ulp 7.6e-6)
        assert g["snailTrail"].shape == o["snailTrail"].shape == (k + 1, 3)
        assert np.abs(g["snailTrail"] - o["snailTrail"]).max() < k * (TOL_M + 5.0 * TOL_RAD)


def test_callers_edge_cases(ctx, scans):
    """Empty / tiny / NaN clouds, zero iterations, odd map capacities: defined behaviour, no device faults."""
    import torch
    from icet_b200 import Node, PointMap, api
    op = api.OdometryParams(2.0, 1, 10.0, 0.0, 0.0)
    # empty and tiny clouds through a chained node: X stays the seed, pose stays the identity
    nd = Node(ctx, api.make_params(), op, 1024)
    assert nd.push(np.zeros((0, 3), np.float32)) is None
    for cloud in (np.zeros((0, 3), np.float32), np.ones((5, 3), np.float32) * 3, np.full((40, 3), np.nan, np.float32)):
        res, pose = nd.push(cloud)
        assert res["status"] == 0 and not res["X"].any() and np.isfinite(pose["X_homo"]).all()
        np.testing.assert_array_equal(pose["X_homo"], np.eye(4, dtype=np.float32))
    assert int(pose["n_points"]) == 0            # NaN rows never pass `distance > minD`
    nd.close()
    # zero iterations: the chain hands the seed through unchanged
    seed = np.array([0.1, 0, 0, 0, 0, 0.01], np.float32)
    nd = Node(ctx, api.make_params(runlen=0), op, scans[0].shape[0], X0=seed)
    nd.push(scans[0])
    for s in scans[1:3]:
        res, pose = nd.push(s)
        np.testing.assert_array_equal(res["X"], seed)
    nd.close()
    # a clouds larger than the node's capacity is refused, the node stays usable
    nd = Node(ctx, api.make_params(), op, 1000)
    with pytest.raises(Exception):
        nd.push(np.zeros((2000, 3), np.float32))
    assert nd.push(np.zeros((10, 3), np.float32)) is None
    nd.close()
    # map: capacity that is not a multiple of 4, insertions of 0 rows and of more rows than the ring holds
    pm = PointMap(ctx, 1003)
    X = torch.zeros(6, dtype=torch.float32, device="cuda")
    pts = torch.arange(3 * 2500, dtype=torch.float32, device="cuda").reshape(3, 2500).contiguous()
    pm.add_scan_device(pts.data_ptr(), 2500, 2500, X.data_ptr(), count=0)
    assert pm.get().shape == (0, 3)
    pm.add_scan_device(pts.data_ptr(), 2500, 2500, X.data_ptr(), count=2500)
    m = pm.get()
    assert m.shape == (1003, 3)
    np.testing.assert_array_equal(m[:, 0], np.arange(2500 - 1003, 2500, dtype=np.float32))   # the newest 1003 rows, oldest first
    pm.close()
    # ingest of an empty cloud, registration of empty clouds
    r = ctx.register_clouds(np.zeros((0, 3), np.float64), np.zeros((0, 3), np.float32))
    assert r["status"] == 0 and not r["X"].any()
    # shipped-order mode with NaN rows and dropped returns in scan 1
    c = scans[0].copy()
    c[::97] = np.nan
    r = ctx.register(c, scans[1], params=api.make_params(flags=api.FLAG_SHIPPED_ORDER))
    assert np.isfinite(r["X"]).all()
    ctx.synchronize()
