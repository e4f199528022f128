#!/usr/bin/env python
"""Stress of the persistent loop kernel: many small registrations; a protocol failure shows up as an IcetError
(watchdog) instead of a hang.  GPU box."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icet_b200
from icet_b200 import api
from tools import synth_host
ctx = icet_b200.Context(0)
sc = synth_host.scans(6, first_scan=0, seed=20240, rings=64, azim=2048)
z = np.zeros((3, 4096), np.float32)
x0 = np.array([0.1, 0, 0, 0, 0, 0.01], np.float32)
cases = [(z, z), (np.zeros((3, 0), np.float32), z), (z, np.zeros((3, 0), np.float32)),
         (np.ones((3, 10), np.float32), np.ones((3, 17), np.float32)), (sc[0], sc[1]), (sc[0][:, :5000].copy(), sc[1][:, :7000].copy())]
p = api.make_params(flags=api.FLAG_PERSISTENT_LOOP)
ref = {}
t0 = time.time()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
bad = 0
for rep in range(n):
    for ci, (a, b) in enumerate(cases):
        try:
            r = ctx.register(a, b, X0=x0, params=p)
        except icet_b200.IcetError as e:
            print("rep %d case %d: %s" % (rep, ci, e), flush=True)
            bad += 1
            continue
        if ci not in ref:
            ref[ci] = r.tobytes()
        elif ref[ci] != r.tobytes():
            print("rep %d case %d: result differs from the first run" % (rep, ci), flush=True)
            bad += 1
    if rep % 50 == 0:
        # multi-pair batches through the persistent kernel
        for P in (2, 3, 5):
            try:
                ctx.register_batch([sc[k] for k in range(P)], [sc[k + 1] for k in range(P)], params=p)
            except icet_b200.IcetError as e:
                print("rep %d batch %d: %s" % (rep, P, e), flush=True)
                bad += 1
    if bad > 5 or time.time() - t0 > 150:
        break
print("loop stress: %d reps, %d failures, %.1f s" % (rep + 1, bad, time.time() - t0))
