#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_default_r1.json 2> gpurun_out/bench_default_r1.err
tail -4 gpurun_out/bench_default_r1.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_default_r1.json') if l.startswith('{')][-1]);print(d['ms_per_step'], d['value'], d['step_roofline']['frac'], d['step_traffic_as_built']);print(d['roofline']);print(d['e2e']);print(d['cpu_baseline']);print(d['clocks'], d['gpu_launches'])"
( time timeout 600 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/bench_reference_r1.json 2> gpurun_out/bench_reference_r1.err
tail -4 gpurun_out/bench_reference_r1.err; cat gpurun_out/bench_reference_r1.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/launches_r1_1024.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log
