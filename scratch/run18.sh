#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fft_|xpass|rk_stage|rot_kernel" -s 30 -c 104 --csv --log-file gpurun_out/launches_r1_1024.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log | cut -c1-200
