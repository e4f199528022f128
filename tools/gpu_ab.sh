#!/bin/bash
# quick GPU visit: parity tests + bench + timeline
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -m gpu -x -q --timeout 60 --timeout-method=thread 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== bench" ; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> $OUT/bench.err > $OUT/bench.json; python - <<EOF
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "lat", d["latency"])
print(d["kernel_ms_per_step"]); print(d["roofline"])
EOF
tail -3 $OUT/bench.err
echo "== timeline"; timeout 100 python tools/timeline.py 2>&1 | tail -4 | tee $OUT/timeline.txt
