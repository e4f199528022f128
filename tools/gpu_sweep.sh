#!/bin/bash
# resident bench over compute lanes x chunk size (no tests)
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  timeout 300 python bench.py --steps 10 --warmup 3 --lanes $1 --chunk $2 --no-cpu-baseline --no-e2e --no-latency --no-callers --no-configs 2> $OUT/bench_$1_$2.err > $OUT/bench_$1_$2.json
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$1_$2.json").read().strip().splitlines()[-1])
    print("lanes $1 chunk $2: value", round(d["value"]), "ms/step", round(d["ms_per_step"],2))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$1_$2.err").read()[-800:])
PY
}
for spec in "4 256" "4 512" "8 512" "4 256" "4 512" "8 512"; do run $spec; done
