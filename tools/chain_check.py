"""GPU check: chained batch (ICET_B200_FLAG_CHAIN_X0) == sequential single-pair calls seeded with the previous X."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icet_b200
from icet_b200 import api
from tools import synth_host

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 9
scans = synth_host.scans(ns, first_scan=0, seed=20240, rings=64, azim=2048)
ctx = icet_b200.Context(0)
p1 = api.make_params()
seq = []
x = np.zeros(6, np.float32)
for k in range(ns - 1):
    r = ctx.register(scans[k], scans[k + 1], x, p1)
    x = r["X"].copy()
    seq.append(r.copy())
pc = api.make_params(flags=api.FLAG_CHAIN_X0)
ctx.set_chunk(3)  # several chunks: the seed crosses chunk boundaries
s1 = [scans[k] for k in range(ns - 1)]
s2 = [scans[k + 1] for k in range(ns - 1)]
t0 = time.perf_counter()
out = ctx.register_batch(s1, s2, None, pc)
dt = time.perf_counter() - t0
for k in range(ns - 1):
    same = out[k]["X"].tobytes() == seq[k]["X"].tobytes() and out[k]["Q"].tobytes() == seq[k]["Q"].tobytes()
    print(k, out[k]["X"], "bitwise equal" if same else "DIFF %s" % (out[k]["X"] - seq[k]["X"]))
print("chained batch of %d pairs: %.2f ms" % (ns - 1, dt * 1e3))
ctx.set_chunk(0)
out2 = ctx.register_batch(s1, s2, None, pc)
print("one chunk == three-pair chunks:", out2["X"].tobytes() == out["X"].tobytes())
