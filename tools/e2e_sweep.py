#!/usr/bin/env python
"""Throughput of the batch entry points for several chunk sizes / lane counts (GPU box)."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icet_b200
from icet_b200 import api

P, NPTS = 512, 131072
ctx = icet_b200.Context(0)
stream = torch.cuda.current_stream()
ctx.set_stream(stream.cuda_stream)
scans = torch.empty((P + 1, 3, NPTS), dtype=torch.float32, device="cuda")
ctx.synth_scans_device(scans.data_ptr(), P + 1, first_scan=0, seed=20240, rings=64, azim=2048)
host = torch.empty((P + 1, 3, NPTS), dtype=torch.float32, pin_memory=True)
host.copy_(scans)
torch.cuda.synchronize()
out = torch.zeros((P, 56), dtype=torch.float32, pin_memory=True)
ptr0, stride = host.data_ptr(), 3 * NPTS * 4
p1 = [ptr0 + i * stride for i in range(P)]
p2 = [ptr0 + (i + 1) * stride for i in range(P)]
nn = np.full(P, NPTS, np.int32)
dst = torch.empty_like(scans)
for _ in range(2):
    dst.copy_(host, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    dst.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print("raw H2D of %d MB: %.2f ms = %.1f GB/s -> ceiling %.0f pairs/s" % (host.numel() * 4 / 1e6, dt * 1e3, host.numel() * 4 / dt / 1e9, P / dt))
del dst
params = api.make_params(7, 24, 75, 25, 0.1, 0.1)
res = torch.zeros((P, 56), dtype=torch.float32, device="cuda")
for lanes in (1, 2):
    ctx.set_lanes(lanes)
    for hc in (32, 64, 128, 256):
        ctx.set_host_chunk(hc)
        for _ in range(2):
            ctx.register_batch_ptrs(p1, nn, p2, nn, out.data_ptr(), params=params, device=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            ctx.register_batch_ptrs(p1, nn, p2, nn, out.data_ptr(), params=params, device=False)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        print("lanes=%d host_chunk=%3d: %.2f ms per %d pairs = %.0f pairs/s (host buffers)" % (lanes, hc, dt * 1e3, P, P / dt))
    for ch in (32, 64, 128, 256):
        ctx.set_chunk(ch)
        for _ in range(2):
            ctx.register_sequence_device(scans.data_ptr(), P + 1, NPTS, res.data_ptr(), params)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            ctx.register_sequence_device(scans.data_ptr(), P + 1, NPTS, res.data_ptr(), params)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print("lanes=%d device chunk=%3d: %.2f ms per %d pairs = %.0f pairs/s (HBM resident)" % (lanes, ch, ms, P, P / ms * 1e3))
    ctx.set_chunk(0)
