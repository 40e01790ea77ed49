#!/bin/bash
# round 2, run 4: TMA-pipelined strided passes -- parity + timing
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2/run4_tests.txt
cat gpurun_out/r2/run4_tests.txt
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run4_tma.err | tee gpurun_out/r2/run4_tma.json | summ "TMA l2=128"
B2_STMA_L2=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run4_tma0.err | tee gpurun_out/r2/run4_tma0.json | summ "TMA l2=none"
B2_STMA_L2=2 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run4_tma2.err | tee gpurun_out/r2/run4_tma2.json | summ "TMA l2=256"
B2_STMA=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run4_ldg.err | tee gpurun_out/r2/run4_ldg.json | summ "LDG"
timeout 300 python bench.py --size 512 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run4_512.err | tee gpurun_out/r2/run4_512.json | summ "512 TMA"
B2_STMA=0 timeout 300 python bench.py --size 512 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>gpurun_out/r2/run4_512l.err | tee gpurun_out/r2/run4_512l.json | summ "512 LDG"
for f in gpurun_out/r2/run4_*.err; do echo "== $f"; tail -n 3 $f; done 2>&1 | head -40
