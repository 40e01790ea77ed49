#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for n in 256 512 1024; do
python bench.py --n $n --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench${n}_v6.json 2> gpurun_out/bench${n}_v6.err; python -c "
import json;d=json.load(open('gpurun_out/bench${n}_v6.json'));print($n, d['ms_per_step'], d['step_roofline']['frac']);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
tail -3 gpurun_out/bench${n}_v6.err
done
