OUT=gpurun_out/r02v; mkdir -p $OUT
ARGS="--steps 1 --warmup 1 --pairs-per-gpu 256 --lanes 1 --no-cpu-baseline --no-latency --no-e2e --no-callers --no-configs"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_vox2|k_solve6' -s 4 -c 4 -f -o $OUT/prof_vox python bench.py $ARGS > $OUT/ncu.log 2>&1
