#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:xpass_fused|fft_strided" -s 6 -c 5 -o gpurun_out/prof_r1_v1_1024 python bench.py --n 1024 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out/
