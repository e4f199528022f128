#!/bin/bash
# full ncu captures: (a) the 7 scan-2 pass launches of one 256-pair chunk, (b) the first launch of each set-up kernel
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
ARGS="--steps 1 --warmup 1 --pairs-per-gpu 256 --lanes 1 --no-cpu-baseline --no-latency --no-e2e --no-callers --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv \
  python bench.py $ARGS > $OUT/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass2 -s 7 -c 7 -f -o $OUT/prof \
  python bench.py $ARGS > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_scan1_bin|k_scatter|k_cluster|k_pass<|k_vox2|k_solve6' -s 6 -c 6 -f -o $OUT/prof_setup \
  python bench.py $ARGS > $OUT/ncu_setup.log 2>&1
ls -la $OUT
