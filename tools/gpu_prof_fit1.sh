OUT=gpurun_out/r02f1; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fit1 -s 5 -c 1 -f -o $OUT/prof_fit1 python tools/single_pair_profile.py > $OUT/ncu.log 2>&1
