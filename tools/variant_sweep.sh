#!/bin/bash
# EXPERIMENT: pass-kernel shapes (ICET_B200_PASS_VARIANT), device-resident bench, one lane
mkdir -p gpurun_out/var
for v in 0 1 2 3 4 5; do
  ICET_B200_PASS_VARIANT=$v timeout 120 python bench.py --steps 4 --warmup 3 --lanes 1 --no-cpu-baseline --no-latency --no-e2e 2>/dev/null > gpurun_out/var/v$v.json
  python - <<EOF
import json
d=json.loads(open("gpurun_out/var/v$v.json").read().strip().splitlines()[-1])
k=d["kernel_ms_per_step"]
print("variant $v: value %.0f pairs/s  k_pass<scan2> %.3f ms  k_pass<scan1> %.3f ms (per 512 pairs)" % (d["value"], k["k_pass<scan2>"], k["k_pass<scan1>"]))
EOF
done
