#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for n in 512 1024; do
timeout 300 python bench.py --size $n --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench${n}_v7.json 2> gpurun_out/bench${n}_v7.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench${n}_v7.json') if l.startswith('{')][-1]);print($n, d['ms_per_step'], d['step_roofline']['frac']);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
tail -3 gpurun_out/bench${n}_v7.err
done
