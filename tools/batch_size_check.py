#!/usr/bin/env python
"""Is a bigger batch faster per pair because of pipelining, or because the synthetic sequence gets easier?
Registers the SAME 512 pairs once, twice, four and eight times in one call (GPU box)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icet_b200
from icet_b200 import api
P, NPTS = 512, 131072
ctx = icet_b200.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
scans = torch.empty((P + 1, 3, NPTS), dtype=torch.float32, device="cuda")
ctx.synth_scans_device(scans.data_ptr(), P + 1, first_scan=first, seed=20240, rings=64, azim=2048)
torch.cuda.synchronize()
nz = (scans[:, 0] != 0) | (scans[:, 1] != 0) | (scans[:, 2] != 0)
print("first scan %d: non-zero points per scan: first %d, middle %d, last %d" % (first, nz[0].sum(), nz[P // 2].sum(), nz[P].sum()))
ptr0, stride = scans.data_ptr(), 3 * NPTS * 4
params = api.make_params()
for rep in (1, 2, 4, 8):
    n = P * rep
    p1 = [ptr0 + (i % P) * stride for i in range(n)]
    p2 = [ptr0 + ((i % P) + 1) * stride for i in range(n)]
    nn = np.full(n, NPTS, np.int32)
    res = torch.zeros((n, 56), dtype=torch.float32, device="cuda")
    for _ in range(2):
        ctx.register_batch_ptrs(p1, nn, p2, nn, res.data_ptr(), params=params, device=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        ctx.register_batch_ptrs(p1, nn, p2, nn, res.data_ptr(), params=params, device=True)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%5d pairs per call: %.2f ms = %.0f pairs/s" % (n, ms, n / ms * 1e3))
