#!/bin/bash
# resident bench over compute lanes x chunk size (no tests)
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  timeout 300 python bench.py --steps 5 --warmup 3 --lanes $1 --chunk $2 --no-cpu-baseline --no-e2e --no-latency --no-callers --no-configs 2> $OUT/bench_$1_$2.err > $OUT/bench_$1_$2.json
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$1_$2.json").read().strip().splitlines()[-1])
    print("lanes $1 chunk $2: value", round(d["value"]), "ms/step", round(d["ms_per_step"],2))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$1_$2.err").read()[-800:])
PY
}
for l in 4 6 8; do for c in 128 256; do run $l $c; done; done
run 2 256
run 8 64
