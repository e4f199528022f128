#!/bin/bash
# quick GPU visit: parity tests + sweeps
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== sweep"; timeout 300 python tools/e2e_sweep.py 2>&1 | tee $OUT/e2e_sweep.txt
