#!/bin/bash
# compute-sanitizer over the GPU parity tests; final logs land in gpurun_out/sanitizer_r02/ (copied to profiles/sanitizer_r02/).
#   gpurun --timeout 2400 -- 'bash tools/gpu_sanitize.sh'
OUT=gpurun_out/sanitizer_r02
mkdir -p $OUT
export ICET_B200_LOOP_TIMEOUT_MS=600000   # the device-side watchdog of k_loop must not fire under a 50x slowdown
CS="compute-sanitizer --error-exitcode 1 --launch-timeout 0"
INC_SEL="scan2_classes_vs_oracle_listed and frame"
run() {  # tool, log name, pytest args...
  tool=$1; log=$2; shift 2
  echo "== $tool: $*"
  timeout 1500 $CS --tool $tool python -m pytest "$@" -q -x -p no:cacheprovider > $OUT/$log 2>&1
  echo "exit $?" >> $OUT/$log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit " $OUT/$log | tail -4
}
# memcheck: the whole GPU suite except the 4096-pair and the 1200-call stress tests
run memcheck memcheck_parity.txt tests/test_gpu_parity.py -m gpu -k "not full_size and not stress"
run memcheck memcheck_incremental_nodes.txt tests/test_gpu_incremental.py tests/test_gpu_nodes.py -m gpu
run racecheck racecheck.txt tests/test_gpu_parity.py tests/test_gpu_incremental.py -m gpu -k "(stage_parity_fixture and frame-None) or big_cells or ($INC_SEL) or submap_config"
run synccheck synccheck.txt tests/test_gpu_parity.py tests/test_gpu_incremental.py -m gpu -k "(stage_parity_fixture and frame-None) or big_cells or ($INC_SEL) or chained_batch"
run initcheck initcheck.txt tests/test_gpu_parity.py tests/test_gpu_incremental.py tests/test_gpu_nodes.py -m gpu -k "(stage_parity_fixture and frame-None) or chained_batch or big_cells or shipped_order or ($INC_SEL) or node or map"
ls -la $OUT
