#!/bin/bash
# round 2, run 2: ncu --set full of the paired x pass and of the four strided passes (1024^3)
mkdir -p gpurun_out/r2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xpass_pair -s 4 -c 1 -o gpurun_out/r2/prof_xpair_v0 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/run2_ncu_x.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_strided -s 4 -c 4 -o gpurun_out/r2/prof_strided_r1final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/run2_ncu_s.log 2>&1
tail -3 gpurun_out/r2/run2_ncu_x.log gpurun_out/r2/run2_ncu_s.log
ls -la gpurun_out/r2/
