#!/usr/bin/env python
"""Where one iteration of the persistent loop kernel spends its time (single pair; GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icet_b200
from tools import synth_host
ctx = icet_b200.Context(0)
sc = synth_host.scans(2, first_scan=0, seed=20240, rings=64, azim=2048)
for rep in range(3):
    r, g = ctx.register(sc[0], sc[1], dump=True)
tl = ctx.debug_timeline(7).astype(np.int64)
names = ["last vox seen", "vox begin", "vox end", "partials sum", "solve done", "published", "tile0 begin", "tile0 end"]
t0 = tl[0, 6]
print("iteration:   " + "  ".join("%12s" % n for n in names), " (us since tile 0 of iteration 0 began)")
for it in range(7):
    print("it %d         " % it + "  ".join("%12.2f" % ((tl[it, k] - t0) / 1e3) for k in range(8)))
print("per iteration: tiles+wait %.1f us (tile0 begin -> latest vox begin), vox algebra %.1f, last-arrival %.1f, partial sums %.1f, "
      "solve %.1f, publish %.1f, observe (published -> next tile0 begin) %.1f; tile0 itself %.1f" % tuple(np.mean(x) / 1e3 for x in (
          tl[:, 1] - tl[:, 6], tl[:, 2] - tl[:, 1], tl[:, 0] - tl[:, 2], tl[:, 3] - tl[:, 0], tl[:, 4] - tl[:, 3],
          tl[:, 5] - tl[:, 4], tl[1:, 6] - tl[:-1, 5], tl[:, 7] - tl[:, 6])))
print("latest vox task: gate+algebra %.1f us, warp reduction + stores %.1f, fence + mask %.1f" % tuple(np.mean(x) / 1e3 for x in (
    tl[:, 8] - tl[:, 1], tl[:, 9] - tl[:, 8], tl[:, 2] - tl[:, 9])))
print("X =", r["X"], "n_used", r["n_used"])
if tl.shape[0] > 3:
    ts = ctx.tile_stamps.astype(np.int64)
    ok = ts[:, 0] > 0
    pub = tl[2, 5]  # iteration 3 starts when iteration 2 is published
    b, e = (ts[ok, 0] - pub) / 1e3, (ts[ok, 1] - pub) / 1e3
    pc = lambda a: "min %.1f p50 %.1f p90 %.1f p99 %.1f max %.1f" % (a.min(), np.percentile(a, 50), np.percentile(a, 90), np.percentile(a, 99), a.max())
    print("iteration 3, %d tiles, us after publication: begin %s | end %s | duration %s" % (ok.sum(), pc(b), pc(e), pc(e - b)))
    mid = (ctx.tile_mid.astype(np.int64)[ok] - pub) / 1e3
    print("phase A %s | phase B + flush %s" % (pc(mid - b), pc(e - mid)))
    print("latest vox begin of iteration 3: %.1f us after publication" % ((tl[3, 1] - pub) / 1e3))
