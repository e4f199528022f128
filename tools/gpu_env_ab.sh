#!/bin/bash
# resident bench under several environment settings:  bash tools/gpu_env_ab.sh <tag> "NAME VAR=val VAR=val" "NAME2 ..." ...
# (TESTS=1 runs the GPU parity tests first)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ "${TESTS:-0}" = "1" ]; then
  echo "== pytest -m gpu" ; timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method=thread 2>&1 | grep -v "^$" | tail -15 | cut -c1-220 | tee $OUT/pytest_gpu.txt
fi
for spec in "$@"; do
  set -- $spec
  name=$1; shift
  echo "== bench $name: $@"
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-latency --no-callers --no-configs ${BENCH_ARGS:-} 2> $OUT/bench_$name.err > $OUT/bench_$name.json
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items() if v})
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$name.err").read()[-1500:])
PY
done
