#!/bin/bash
mkdir -p gpurun_out
python bench.py --n 1024 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench1024_v5.json 2> gpurun_out/bench1024_v5.err; python -c "
import json;d=json.load(open('gpurun_out/bench1024_v5.json'));print(d['ms_per_step'], d['step_roofline']['frac']);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
ncu --set full --clock-control none --import-source on -k "regex:xpass_fused" -s 0 -c 1 -o gpurun_out/prof_r1_v5_x512 python bench.py --n 512 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
