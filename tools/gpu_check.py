#!/usr/bin/env python
"""Developer script: GPU vs oracle, stage by stage, on the bundled fixture pairs (prints a report)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import icet_b200
from icet_b200 import api
from oracle import pyoracle as po
np.set_printoptions(precision=6, suppress=True, linewidth=220)

def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

def main():
    ctx = icet_b200.Context()
    for name in ("frame", "sample_pc"):
        d = np.load(os.path.join(ROOT, "tests", "golden", "inputs_%s.npz" % name))
        s1, s2 = d["scan1"], d["scan2"]
        print("=== %s  n1=%d n2=%d" % (name, s1.shape[1], s2.shape[1]))
        o = po.run(s1, s2, dumps="all")
        oh = po.run(s1, s2, dumps="small", precise=True)
        sph, cell = ctx.spherical_bins(s1)
        print("sph1 bit-exact:", [(sph[k].view(np.int32) == o.sph1[k].view(np.int32)).mean() for k in range(3)],
              "max ulp diff th/ph:", [int(np.abs(sph[k].view(np.int32).astype(np.int64) - o.sph1[k].view(np.int32)).max()) for k in range(3)])
        print("cell1 mismatches:", int((cell != o.cell1).sum()))
        t = time.time()
        r, g = ctx.register(s1, s2, dump=True)
        print("gpu register (with dump) %.1f ms" % ((time.time() - t) * 1e3))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", "dump_%s.npz" % name), res=r, sph1=sph, cell1=cell, **g)
        print("cnt1 equal:", np.array_equal(g["cnt1"], o.cnt1), " bounds maxabs diff:", float(np.abs(g["bounds"] - o.bounds).max()),
              " nonequal rows:", int((np.abs(g["bounds"] - o.bounds).max(1) > 0).sum()))
        print("has1 equal:", np.array_equal(g["has1"], o.has1), int(g["has1"].sum()), int(o.has1.sum()),
              " nin1 mism:", int((g["nin1"][o.has1 > 0] != o.nin1[o.has1 > 0]).sum()))
        m = (o.has1 > 0) & (g["has1"] > 0)
        mu_err = np.abs(g["mu1"][m] - o.mu1[m]).max(1) / np.abs(o.mu1[m]).max(1)
        sg_err = np.abs(g["sigma1"][m] - o.sigma1[m]).reshape(-1, 9).max(1) / np.abs(o.sigma1[m]).reshape(-1, 9).max(1)
        print("mu1 rel err max %.2e  sigma1 rel err max %.2e (vs fp32 oracle); vs double twin: mu %.2e sigma %.2e" % (
            mu_err.max(), sg_err.max(),
            (np.abs(g["mu1"][m] - oh.mu1[m]).max(1) / np.abs(oh.mu1[m]).max(1)).max(),
            (np.abs(g["sigma1"][m] - oh.sigma1[m]).reshape(-1, 9).max(1) / np.abs(oh.sigma1[m]).reshape(-1, 9).max(1)).max()))
        print("lmask mismatching voxels:", int((g["lmask"][m] != o.lmask[m]).any(1).sum()),
              " evec sign-exact voxels: %d / %d" % (int((np.abs(g["evec1"][m] - o.evec1[m]).reshape(-1, 9).max(1) < 1e-3).sum()), int(m.sum())))
        for it in range(o.Xit.shape[0]):
            mm = (g["cnt2"][it] >= 0)
            print(" it %d  X gpu %s\n       X ora %s   used %d/%d cnt2 mism %d nin2 mism %d" % (
                it, g["Xit"][it], o.Xit[it], int(g["used2"][it].sum()), int(o.used2[it].sum()),
                int((g["cnt2"][it][mm] != o.cnt2[it][mm]).sum()),
                int((g["nin2"][it][mm] != o.nin2[it][mm]).sum())))
            if it == 0:
                print("       HTWH rel %.2e  HTWdz rel %.2e" % (rel(g["HTWH"][0], o.HTWH[0]), rel(g["HTWdz"][0], o.HTWdz[0])))
        print("X   gpu", r["X"], "\nX   ora", o.X, "\nX   o64", oh.X)
        print("dX m", np.abs(r["X"][:3] - o.X[:3]).max(), " rad", np.abs(r["X"][3:] - o.X[3:]).max(),
              "| vs o64: m", np.abs(r["X"][:3] - oh.X[:3]).max(), " rad", np.abs(r["X"][3:] - oh.X[3:]).max())
        print("stds gpu", r["pred_stds"], "\nstds ora", o.pred_stds)
        print("Q rel (fro) vs fp32 oracle: %.2e   vs double twin: %.2e" % (
            np.linalg.norm(r["Q"] - o.Q) / np.linalg.norm(o.Q), np.linalg.norm(r["Q"] - oh.Q) / np.linalg.norm(oh.Q)))
        print("n_gauss1", r["n_gauss1"], "n_used", r["n_used"], "cond", r["cond"], "status", r["status"])
        # timing of the single-pair host API
        ts = []
        for _ in range(20):
            t = time.time(); ctx.register(s1, s2); ts.append(time.time() - t)
        print("host-API single pair: p50 %.3f ms  min %.3f ms" % (np.median(ts) * 1e3, np.min(ts) * 1e3))
    print("kernel launches:", ctx.kernel_launches)

if __name__ == "__main__":
    main()
