#!/bin/bash
# quick GPU visit: parity tests + resident bench at several rebuild bounds of the incremental loop
TAG=${1:-ab2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method=thread 2>&1 | grep -v "^$" | tail -25 | cut -c1-220 | tee $OUT/pytest_gpu.txt
run() {
  echo "== bench $1 SA=$2 SB=$3"
  ICET_B200_INC_SA=$2 ICET_B200_INC_SB=$3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-latency --no-callers --no-configs 2> $OUT/bench_$1.err > $OUT/bench_$1.json
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$1.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()})
    print(" roofline frac", round(d["roofline"]["frac"],3), "parity", d.get("parity_vs_oracle"))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$1.err").read()[-1500:])
PY
}
run a 1e-3 0.05
run b 2e-3 0.08
run c 4e-3 0.12
