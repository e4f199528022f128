#!/usr/bin/env python
"""Persistent loop kernel vs the split loop for several batch sizes, each in its own process with a timeout
(a hang must not eat the GPU lease).  GPU box."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
import icet_b200
from icet_b200 import api
P = int(sys.argv[1]); chunk = int(sys.argv[2])
ctx = icet_b200.Context(0)
ctx.set_chunk(chunk)
scans = torch.empty((P + 1, 3, 131072), dtype=torch.float32, device="cuda")
ctx.synth_scans_device(scans.data_ptr(), P + 1, first_scan=0, seed=20240, rings=64, azim=2048)
out = {}
for name, flags in (("split", api.FLAG_UNFUSED_LOOP), ("loop", api.FLAG_PERSISTENT_LOOP)):
    res = torch.zeros((P, 56), dtype=torch.float32, device="cuda")
    p = api.make_params(7, 24, 75, 25, 0.1, 0.1, flags=flags)
    ctx.register_sequence_device(scans.data_ptr(), P + 1, 131072, res.data_ptr(), p); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.register_sequence_device(scans.data_ptr(), P + 1, 131072, res.data_ptr(), p)
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / 3
    out[name] = res.cpu().numpy().view(api.RESULT_DTYPE).reshape(-1)
    print("P=%%d chunk=%%d %%s: %%.1f us/pair" %% (P, chunk, name, dt * 1e6 / P), flush=True)
a, b = out["split"], out["loop"]
print("   max|dX| = %%.2e, n_used equal: %%s" %% (np.abs(a["X"] - b["X"]).max(), bool((a["n_used"] == b["n_used"]).all())), flush=True)
''' % ROOT
for P, chunk in ((1, 256), (2, 256), (3, 256), (8, 256), (16, 16), (64, 64), (256, 256), (300, 256)):
    try:
        r = subprocess.run([sys.executable, "-c", CHILD, str(P), str(chunk)], capture_output=True, text=True, timeout=40)
        print(r.stdout.strip() or r.stderr.strip()[-500:], flush=True)
    except subprocess.TimeoutExpired as e:
        print("P=%d chunk=%d: TIMEOUT (hang) -- partial output: %s" % (P, chunk, (e.stdout or b"").decode()[-300:]), flush=True)
        break
