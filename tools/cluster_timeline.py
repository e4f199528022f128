#!/usr/bin/env python
"""Per-iteration phase stamps of the cluster-form loop (k_loop_cluster, %globaltimer of CTA 0 / thread 0) for one
synthetic 64-channel pair, plus the per-kernel device time of the whole single-pair call.

    python tools/cluster_timeline.py [first_scan] [x0 | seed]     (GPU box)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icet_b200  # noqa: E402
from icet_b200 import api  # noqa: E402
from tools import synth_host  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seeded = len(sys.argv) > 2
sc = synth_host.scans(2, first_scan=first)
ctx = icet_b200.Context(0)
x0 = None
if seeded:
    x0 = ctx.register(sc[0], sc[1])["X"]   # a converged seed: the chained-odometry situation
p = api.make_params()
for _ in range(3):
    ctx.register(sc[0], sc[1], X0=x0, params=p, dump=True)
tl = ctx.debug_timeline(p.runlen).astype(np.int64)
names = ["tiles", "sync1", "vox", "sync2", "solve", "sync3"]
print("iteration   " + "  ".join("%7s" % n for n in names) + "    total (us)")
for it in range(p.runlen):
    d = np.diff(tl[it, :7]) / 1e3
    print("%9d   " % it + "  ".join("%7.2f" % v for v in d) + "   %7.2f" % ((tl[it, 6] - tl[it, 0]) / 1e3))
print("loop total %.1f us" % ((tl[p.runlen - 1, 6] - tl[0, 0]) / 1e3))
print("solve phase: partial sums gathered (DSMEM) / solved and stored / next iteration announced, us after the phase began")
for it in range(p.runlen):
    print("%9d   %7.2f %7.2f %7.2f" % (it, (tl[it, 11] - tl[it, 4]) / 1e3, (tl[it, 12] - tl[it, 4]) / 1e3, (tl[it, 5] - tl[it, 4]) / 1e3))
print("voxel phase of CTA 0: last voxel done / last warp reduced, us after the phase began")
for it in range(p.runlen):
    print("%9d   %7.2f %7.2f   (phase %.2f)" % (it, (tl[it, 8] - tl[it, 2]) / 1e3, (tl[it, 9] - tl[it, 2]) / 1e3, (tl[it, 3] - tl[it, 2]) / 1e3))
ctx.set_profile(True)
for _ in range(20):
    ctx.register(sc[0], sc[1], X0=x0, params=p)
prof = ctx.get_profile()
ctx.set_profile(False)
print({k: round(1e3 * v[0] / max(1, v[1]), 1) for k, v in prof.items() if v[1]}, "us per launch")
