#!/bin/bash
# round 2, run 1: paired x pass -- parity + variants
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2/run1_smi.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2/run1_tests.txt
cat gpurun_out/r2/run1_tests.txt
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
for v in 0 1 2; do
  B2_XVAR=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run1_v$v.err | tee gpurun_out/r2/run1_v$v.json | summ "XVAR=$v"
done
B2_XPASS_OLD=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run1_old.err | tee gpurun_out/r2/run1_old.json | summ "OLD"
B2_XPPG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run1_ppg1.err | tee gpurun_out/r2/run1_ppg1.json | summ "PPG1"
B2_XPPG=16 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run1_ppg16.err | tee gpurun_out/r2/run1_ppg16.json | summ "PPG16"
timeout 300 python bench.py --size 512 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run1_512.err | tee gpurun_out/r2/run1_512.json | summ "512"
