#!/usr/bin/env python
"""Runs each degenerate-input case in its own process with a short timeout (GPU box)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import icet_b200
from icet_b200 import api
case, flags = int(sys.argv[1]), int(sys.argv[2])
z = np.zeros((3, 4096), np.float32)
cases = [(z, z), (np.zeros((3, 0), np.float32), z), (z, np.zeros((3, 0), np.float32)),
         (np.ones((3, 10), np.float32), np.ones((3, 17), np.float32))]
a, b = cases[case]
ctx = icet_b200.Context(0)
x0 = np.array([0.1, 0, 0, 0, 0, 0.01], np.float32)
r = ctx.register(a, b, X0=x0, params=api.make_params(flags=flags))
print("case %%d flags %%d ok X=%%s" %% (case, flags, r["X"]), flush=True)
''' % ROOT
for flags in (2, 0):
    for case in range(4):
        try:
            r = subprocess.run([sys.executable, "-c", CHILD, str(case), str(flags)], capture_output=True, text=True, timeout=25)
            print(r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
        except subprocess.TimeoutExpired:
            print("case %d flags %d: TIMEOUT" % (case, flags), flush=True)
