#!/bin/bash
# latency / callers of the bench under several environment settings:  bash tools/gpu_lat_ab.sh <tag> "name VAR=val ..." ...
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in "$@"; do
  set -- $spec
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 2 --warmup 3 --pairs-per-gpu 512 --no-cpu-baseline --no-e2e 2> $OUT/bench_$name.err > $OUT/bench_$name.json
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    print("$name: p50", round(d["latency"]["device_resident_p50_ms"],4), "pinned host", round(d["latency"]["host_api_pinned_p50_ms"],4),
          "128ch", round(d["configs"]["ouster_128ch"]["p50_ms"],4), "submap", round(d["configs"]["submap_2M"]["p50_ms"],4),
          "chained", round(d["callers"]["odometry_chained"]["pairs_per_s"]), "mapmaker", round(d["callers"]["mapmaker_batched"]["pairs_per_s"]))
except Exception as e:
    print("bench failed", e); print(open("$OUT/bench_$name.err").read()[-1500:])
PY
done
