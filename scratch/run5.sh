#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench512_v2.json 2> gpurun_out/bench512_v2.err; python -c "
import json;d=json.load(open('gpurun_out/bench512_v2.json'));print(d['ms_per_step'], d['step_roofline']['frac']);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
python bench.py --n 1024 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench1024_v2.json 2> gpurun_out/bench1024_v2.err; python -c "
import json;d=json.load(open('gpurun_out/bench1024_v2.json'));print(d['ms_per_step'], d['step_roofline']['frac']);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
tail -3 gpurun_out/bench1024_v2.err
