#!/usr/bin/env python
"""Regenerate tests/golden/*.npz (run in the build container, where /root/reference exists).

inputs_*.npz   the reference's bundled sensor scans (data files, not source code) cast to the
               float32 the reference itself works in (Eigen::MatrixXf), stored as [3, N] planes
golden_*.npz   outputs of the CPU oracle (oracle/icet_oracle.cpp) on those inputs

The reference publishes no expected outputs for this path, so the golden
vectors pin the ORACLE (against drift and host libm differences); they are not reference outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
PAIRS = {
    "frame": ("src/sample_data/frame_804.npy", "src/sample_data/frame_805.npy"),
    "sample_pc": ("python/point_clouds/sample_pc_1.npy", "python/point_clouds/sample_pc_2.npy"),
}
KEEP = ["cnt1", "bounds", "nin1", "has1", "mu1", "sigma1", "eval1", "evec1", "lmask", "cnt2", "nin2",
        "used2", "mu2", "sigma2", "HTWH", "HTWdz", "dx", "Xit", "Qit", "stds_it", "cond_it", "trunc_it"]


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (a, b) in PAIRS.items():
        s1 = po.as_planes(np.load(os.path.join(REF, a)))
        s2 = po.as_planes(np.load(os.path.join(REF, b)))
        np.savez_compressed(os.path.join(OUT, "inputs_%s.npz" % name), scan1=s1, scan2=s2)
        for tag, kw in {"sorted": dict(order_mode=po.ORDER_SORTED),
                        "shipped": dict(order_mode=po.ORDER_REF_SHIPPED),
                        "sorted_x0demo": dict(order_mode=po.ORDER_SORTED, X0=[1, 0, 0, 0, 0, 0])}.items():
            if tag == "sorted_x0demo" and name != "frame":
                continue
            r = po.run(s1, s2, runlen=7, bins_phi=24, bins_theta=75, n=25, thresh=0.1, buff=0.1,
                       dumps="small", **kw)
            d = {k: r.dumps[k] for k in KEEP}
            # per-voxel scan-2 statistics only for the first and last iteration (size)
            for k in ("mu2", "sigma2"):
                d[k] = d[k][[0, -1]]
            np.savez_compressed(os.path.join(OUT, "golden_%s_%s.npz" % (name, tag)), X=r.X,
                                pred_stds=r.pred_stds, Q=r.Q, status=r.status, **d)
            print(name, tag, r.X, int(r.has1.sum()), int(r.used2[-1].sum()))


def submap_golden():
    """BASELINE.json configs[4] at FULL size: the 2 M-point accumulated map of tools/synth_host.submap() against one
    64-channel scan.  Only outputs are stored (the map is regenerated bit for bit from the seed; its hash is kept)."""
    import hashlib
    from tools import synth_host
    mp, cur = synth_host.submap()
    r = po.run(mp, cur, dumps="small")
    np.savez_compressed(os.path.join(OUT, "golden_submap_2M.npz"), X=r.X, pred_stds=r.pred_stds, Q=r.Q,
                        cnt1=r.cnt1, bounds=r.bounds, has1=r.has1, nin1=r.nin1, used2_last=r.used2[-1],
                        Xit=r.Xit, n_map=np.int64(mp.shape[1]),
                        map_sha256=hashlib.sha256(mp.tobytes()).hexdigest(),
                        scan_sha256=hashlib.sha256(cur.tobytes()).hexdigest())
    print("submap", mp.shape, r.X, int(r.has1.sum()), int(r.used2[-1].sum()), "largest cell", int(r.cnt1.max()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "submap":
        submap_golden()
        sys.exit(0)
    main()
    submap_golden()
