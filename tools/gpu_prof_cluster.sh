#!/bin/bash
# full ncu capture of one k_loop_cluster launch (single synthetic pair, latency shape)
TAG=${1:-r02c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_loop_cluster -s 5 -c 1 -f -o $OUT/prof_cluster \
  python tools/single_pair_profile.py > $OUT/ncu_cluster.log 2>&1
ls -la $OUT
