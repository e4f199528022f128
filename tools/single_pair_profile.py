#!/usr/bin/env python
"""Per-kernel device time of ONE pair (latency shape): events around every launch (GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import icet_b200
from icet_b200 import api
ctx = icet_b200.Context(0)
st = torch.cuda.current_stream()
ctx.set_stream(st.cuda_stream)
scans = torch.empty((2, 3, 131072), dtype=torch.float32, device="cuda")
ctx.synth_scans_device(scans.data_ptr(), 2)
res = torch.zeros((1, 56), dtype=torch.float32, device="cuda")
p = api.make_params()
for _ in range(20):
    ctx.register_sequence_device(scans.data_ptr(), 2, 131072, res.data_ptr(), p)
torch.cuda.synchronize()
ctx.set_profile(True)
R = 50
for _ in range(R):
    ctx.register_sequence_device(scans.data_ptr(), 2, 131072, res.data_ptr(), p)
prof = ctx.get_profile()
ctx.set_profile(False)
tot = 0
for k, (ms, n) in prof.items():
    if n:
        print("%-16s %8.2f us  (%d launches)" % (k, ms * 1e3 / R, n // R))
        tot += ms * 1e3 / R
print("sum %.1f us" % tot)
lat = []
for i in range(100):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); ctx.register_sequence_device(scans.data_ptr(), 2, 131072, res.data_ptr(), p); b.record(st)
    torch.cuda.synchronize()
    lat.append(a.elapsed_time(b) * 1e3)
print("end-to-end device latency p50 %.1f us" % np.median(lat))
