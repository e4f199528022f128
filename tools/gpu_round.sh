#!/bin/bash
# One GPU visit: parity tests, smoke, bench (ours + reference arm), ncu launch list and one full capture
# of the dominant kernel.  Run under gpurun from the repo root; everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench" ; timeout 900 python bench.py --steps 5 --warmup 3 2> $OUT/bench.err | tee $OUT/bench.json
tail -3 $OUT/bench.err
if [ "${SKIP_REF:-0}" != "1" ]; then
  echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> $OUT/bench_ref.err | tee $OUT/bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --pairs-per-gpu 256 --lanes 1 --no-cpu-baseline --no-latency --no-e2e --no-callers --no-configs > $OUT/ncu_bench.log 2>&1
  echo "== ncu full capture: the 7 scan-2 pass launches of one 256-pair chunk; the first launch of each scan-1 kernel"
  NARGS="--steps 1 --warmup 1 --pairs-per-gpu 256 --lanes 1 --no-cpu-baseline --no-latency --no-e2e --no-callers --no-configs"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass2 -s 7 -c 7 -f -o $OUT/prof \
    python bench.py $NARGS > $OUT/ncu_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_scan1_bin|k_scatter|k_cluster|k_pass<' -s 0 -c 4 -f -o $OUT/prof_setup \
    python bench.py $NARGS > $OUT/ncu_setup.log 2>&1
  ls -la $OUT
fi
