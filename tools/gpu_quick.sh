#!/bin/bash
# quick GPU visit: parity tests + resident bench with latency / callers / configs (no CPU arm, no e2e)
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method=thread 2>&1 | grep -v "^$" | tail -15 | cut -c1-220 | tee $OUT/pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2> $OUT/bench.err > $OUT/bench.json
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items() if v})
    print("latency", {k:(round(v,4) if isinstance(v,float) else v) for k,v in d["latency"].items()})
    for k,v in d["configs"].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("workload",)})
    print("callers", json.dumps(d["callers"])[:600])
    print("roofline", d["roofline"])
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench.err").read()[-1500:])
PY
