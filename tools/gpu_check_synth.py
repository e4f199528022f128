#!/usr/bin/env python
"""Developer script: GPU vs oracle on synthetic pairs (host generator); saves GPU dumps to gpurun_out/."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import icet_b200
from oracle import pyoracle as po
from tools import synth_host
np.set_printoptions(precision=6, suppress=True, linewidth=220)
NP = int(sys.argv[1]) if len(sys.argv) > 1 else 12
S = synth_host.scans(NP + 1)
ctx = icet_b200.Context()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "synth_scans_first3.npy"), S[:3])
for k in range(NP):
    r, g = ctx.register(S[k], S[k + 1], dump=True)
    o = po.run(S[k], S[k + 1], dumps="small")
    m = (o.has1 > 0) & (g["has1"] > 0)
    evd = np.abs(g["evec1"][m] - o.evec1[m]).reshape(-1, 9).max(1)
    print("pair %2d dX %.2e m %.2e rad | has1 eq %s lmask mism %d evec mism %d | cnt1 eq %s bounds neq %d | used %s vs %s" % (
        k, np.abs(r["X"][:3] - o.X[:3]).max(), np.abs(r["X"][3:] - o.X[3:]).max(), np.array_equal(g["has1"], o.has1),
        int((g["lmask"][m] != o.lmask[m]).any(1).sum()), int((evd > 1e-3).sum()), np.array_equal(g["cnt1"], o.cnt1),
        int((np.abs(g["bounds"] - o.bounds).max(1) > 0).sum()), g["used2"].sum(1), o.used2.sum(1)))
    print("        X gpu", r["X"], " ora", o.X)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "sdump_%02d.npz" % k), res=r, **g)
