#!/bin/bash
# resident bench over batch size x compute lanes x chunk size
TAG=${1:-sweepb}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --pairs-per-gpu $1 --lanes $2 --chunk $3 --no-cpu-baseline --no-e2e --no-latency --no-callers --no-configs 2> $OUT/b.err > $OUT/b.json
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/b.json").read().strip().splitlines()[-1])
    print("pairs $1 lanes $2 chunk $3: value", round(d["value"]), "ms/step", round(d["ms_per_step"],2))
except Exception as e:
    print("bench failed", e); print(open("$OUT/b.err").read()[-800:])
PY
}
for spec in "512 4 512" "512 8 512" "512 2 512" "1024 4 512" "1024 8 512" "2048 4 512" "2048 8 512" "128 4 512" "128 8 512" "128 2 512"; do run $spec; done
