#!/bin/bash
# ncu launch list + full capture of the scan-2 pass kernels of one 256-pair chunk (7 iterations).
#   gpurun --timeout 900 -- 'bash tools/gpu_prof.sh <tag> [kernel regex] [launches]'
TAG=${1:-r02p}; KRE=${2:-k_pass2}; NL=${3:-7}
OUT=gpurun_out/$TAG
mkdir -p $OUT
ARGS="--steps 1 --warmup 1 --pairs-per-gpu 256 --lanes 1 --no-cpu-baseline --no-latency --no-e2e --no-callers --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py $ARGS > $OUT/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $NL -c $NL -f -o $OUT/prof \
  python bench.py $ARGS > $OUT/ncu_full.log 2>&1
ls -la $OUT
