#!/bin/bash
mkdir -p gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_default_r1.json 2> gpurun_out/bench_default_r1.err
tail -4 gpurun_out/bench_default_r1.err
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_default_r1.json') if l.startswith('{')][-1]);print(d['ms_per_step'], d['value'], d['step_roofline']['frac'], d['step_traffic_as_built']['frac']);print(d['roofline']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches']);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_reference_r1.json 2> gpurun_out/bench_reference_r1.err; cut -c1-200 gpurun_out/bench_reference_r1.json
