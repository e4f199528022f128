#!/usr/bin/env python
"""Device-resident p50 latency of BASELINE.json configs[1], [3] and [4] (GPU box):
   [1] one synthetic 64-channel pair, 75 x 24, 7 iterations
   [3] one synthetic 128-channel pair (262 144 points), 150 x 48, 10 iterations
   [4] scan-to-submap: scan 1 = 2 000 000-point map (1000 earlier scans x 2000 random points, expressed in the frame
       of the newest of them by the generator's known poses -- here simply 1000 scans of the sequence sub-sampled, all
       in their own sensor frame shifted by the accumulated mean motion), scan 2 = one 64-channel scan, 75 x 24, 7 it."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icet_b200
from icet_b200 import api

ctx = icet_b200.Context(0)
st = torch.cuda.current_stream()
ctx.set_stream(st.cuda_stream)


def p50(s1, n1, s2, n2, p, reps=100):
    res = torch.zeros((1, 56), dtype=torch.float32, device="cuda")
    lat = []
    for i in range(reps + 20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        ctx.register_batch_ptrs([s1.data_ptr()], [n1], [s2.data_ptr()], [n2], res.data_ptr(), params=p, device=True)
        b.record(st)
        torch.cuda.synchronize()
        if i >= 20:
            lat.append(a.elapsed_time(b))
    r = res.cpu().numpy().view(api.RESULT_DTYPE).reshape(-1)[0]
    return float(np.median(lat)), float(np.percentile(lat, 95)), r


n64 = 64 * 2048
a = torch.empty((2, 3, n64), dtype=torch.float32, device="cuda")
ctx.synth_scans_device(a.data_ptr(), 2, first_scan=10)
m, q, r = p50(a[0], n64, a[1], n64, api.make_params())
print("configs[1] 64-ch pair 75x24 7 it:        p50 %.3f ms  p95 %.3f ms  (voxels used %d)" % (m, q, r["n_used"]))
n128 = 128 * 2048
b = torch.empty((2, 3, n128), dtype=torch.float32, device="cuda")
ctx.synth_scans_device(b.data_ptr(), 2, first_scan=10, rings=128)
m, q, r = p50(b[0], n128, b[1], n128, api.make_params(runlen=10, bins_phi=48, bins_theta=150))
print("configs[3] 128-ch pair 150x48 10 it:     p50 %.3f ms  p95 %.3f ms  (voxels used %d)" % (m, q, r["n_used"]))
# map: 2000 random points of each of 1000 consecutive scans; consecutive scans differ by ~0.5 m of forward motion, which
# is removed with the generator's nominal step so that the map is roughly consistent (its exact consistency does not
# matter for timing: the cost depends on the point count and on the occupancy of the voxels)
ns = 1000
g = torch.Generator(device="cuda").manual_seed(1)
parts = []
for c0 in range(0, ns, 100):
    blk = torch.empty((100, 3, n64), dtype=torch.float32, device="cuda")
    ctx.synth_scans_device(blk.data_ptr(), 100, first_scan=20 + c0)
    torch.cuda.synchronize()
    for k in range(100):
        idx = torch.randperm(n64, device="cuda", generator=g)[:2000]
        pts = blk[k][:, idx].clone()
        nz = (pts != 0).any(0)
        pts[0, nz] -= 0.5 * (ns - 1 - (c0 + k))   # express in the newest frame (nominal forward motion only)
        parts.append(pts)
mp = torch.cat(parts, dim=1).contiguous()
cur = torch.empty((1, 3, n64), dtype=torch.float32, device="cuda")
ctx.synth_scans_device(cur.data_ptr(), 1, first_scan=20 + ns)
m, q, r = p50(mp, mp.shape[1], cur[0], n64, api.make_params(), reps=50)
print("configs[4] %d-point map vs 64-ch scan:  p50 %.3f ms  p95 %.3f ms  (gaussians %d, voxels used %d)" % (mp.shape[1], m, q, r["n_gauss1"], r["n_used"]))
ctx.set_profile(True)
res = torch.zeros((1, 56), dtype=torch.float32, device="cuda")
for _ in range(3):
    ctx.register_batch_ptrs([mp.data_ptr()], [mp.shape[1]], [cur[0].data_ptr()], [n64], res.data_ptr(), params=api.make_params(), device=True)
prof = ctx.get_profile()
ctx.set_profile(False)
print("configs[4] per kernel (ms):", {k: round(v[0] / 3, 3) for k, v in prof.items() if v[1]})
cnt = None
