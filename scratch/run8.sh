#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for v in 0 1 2; do B2_SVAR=$v python scratch/micro_strided.py 1024 3 2>&1 | grep SVAR; done
B2_L2GRAN=128 B2_SVAR=0 python scratch/micro_strided.py 1024 3 2>&1 | grep SVAR | sed 's/^/L2GRAN128 /'
B2_L2GRAN=32 B2_SVAR=0 python scratch/micro_strided.py 1024 3 2>&1 | grep SVAR | sed 's/^/L2GRAN32 /'
for v in 0 1 2 3; do B2_SVAR=$v python scratch/micro_strided.py 512 6 2>&1 | grep SVAR; done
B2_L2GRAN=128 B2_SVAR=1 python scratch/micro_strided.py 512 6 2>&1 | grep SVAR | sed 's/^/L2GRAN128 /'
python scratch/micro_strided.py 256 6 2>&1 | grep SVAR
for n in 512 1024; do
python bench.py --n $n --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench${n}_v4.json 2> gpurun_out/bench${n}_v4.err; python -c "
import json;d=json.load(open('gpurun_out/bench${n}_v4.json'));print(d['ms_per_step'], d['step_roofline']['frac']);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
tail -3 gpurun_out/bench${n}_v4.err
done
