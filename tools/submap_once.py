#!/usr/bin/env python
"""One registration of BASELINE.json configs[4] (2 M-point submap vs one 64-channel scan), for `ncu` launch lists:
   ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file out.csv python tools/submap_once.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import icet_b200  # noqa: E402
from icet_b200 import api  # noqa: E402
from tools import synth_host  # noqa: E402

mp_h, cur_h = synth_host.submap()
ctx = icet_b200.Context(0)
mp, cur = torch.from_numpy(mp_h).cuda(), torch.from_numpy(cur_h).cuda()
res = torch.zeros((1, 56), dtype=torch.float32, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ctx.register_batch_ptrs([mp.data_ptr()], [mp.shape[1]], [cur.data_ptr()], [cur.shape[1]], res.data_ptr(),
                            params=api.make_params(), device=True)
ctx.synchronize()
print(res.cpu().numpy().view(api.RESULT_DTYPE)[0]["X"])
