#!/bin/bash
# quick GPU visit: parity tests + bench
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 300 python -m pytest tests -m gpu -x -q -s --timeout 60 --timeout-method=thread 2>&1 | grep -v "^$" | tail -25 | cut -c1-220 | tee $OUT/pytest_gpu.txt
echo "== bench (1 lane)" ; timeout 300 python bench.py --steps 5 --warmup 3 --lanes 1 --no-cpu-baseline --no-e2e --no-latency 2> $OUT/bench1.err > $OUT/bench1.json; python - <<EOF
import json
d=json.loads(open("$OUT/bench1.json").read().strip().splitlines()[-1])
print("1 lane value", d["value"]); print(d["kernel_ms_per_step"])
EOF
echo "== bench" ; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> $OUT/bench.err > $OUT/bench.json; python - <<EOF
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "lat", d["latency"])
print(d["roofline"])
EOF
tail -3 $OUT/bench.err
