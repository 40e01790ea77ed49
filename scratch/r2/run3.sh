#!/bin/bash
# round 2, run 3: bulk-async staged x pass -- parity + timing
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2/run3_tests.txt
cat gpurun_out/r2/run3_tests.txt
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run3_pf.err | tee gpurun_out/r2/run3_pf.json | summ "PF"
B2_XPPG=8 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run3_pf8.err | tee gpurun_out/r2/run3_pf8.json | summ "PF ppg8"
B2_XPPG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run3_pf1.err | tee gpurun_out/r2/run3_pf1.json | summ "PF ppg1"
B2_XVAR=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run3_pfv1.err | tee gpurun_out/r2/run3_pfv1.json | summ "PF var1"
B2_XNOPF=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run3_nopf.err | tee gpurun_out/r2/run3_nopf.json | summ "NOPF"
timeout 300 python bench.py --size 512 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run3_512.err | tee gpurun_out/r2/run3_512.json | summ "512"
timeout 300 python bench.py --solver ns3d.strat --size 512 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run3_s512.err | tee gpurun_out/r2/run3_s512.json | summ "strat512"
tail -2 gpurun_out/r2/*.err | grep -v "^$" | head -20
