#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
for v in 0 1; do B2_SVAR=$v python scratch/micro_strided.py 1024 3 2>&1 | grep SVAR; done
for v in 0 1 2; do B2_SVAR=$v python scratch/micro_strided.py 512 6 2>&1 | grep SVAR; done
python scratch/micro_strided.py 256 6 2>&1 | grep SVAR
for n in 512 1024; do
python bench.py --n $n --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench${n}_v3.json 2> gpurun_out/bench${n}_v3.err; python -c "
import json;d=json.load(open('gpurun_out/bench${n}_v3.json'));print(d['ms_per_step'], d['step_roofline']['frac']);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
tail -3 gpurun_out/bench${n}_v3.err
done
