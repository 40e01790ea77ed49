#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:xpass_fused" -s 4 -c 1 -o gpurun_out/prof_r1_v8_x1024 python bench.py --size 1024 --steps 1 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full3.log 2>&1
tail -2 gpurun_out/ncu_full3.log
