#!/bin/bash
# full ncu capture of the first launches of the set-up kernels (scan 1 side) of one 256-pair chunk
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
ARGS="--steps 1 --warmup 1 --pairs-per-gpu 256 --lanes 1 --no-cpu-baseline --no-latency --no-e2e --no-callers --no-configs"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_scan1_bin|k_scatter|k_cluster|k_pass<' -s 0 -c 4 -f -o $OUT/prof_setup \
  python bench.py $ARGS > $OUT/ncu_setup.log 2>&1
ls -la $OUT
