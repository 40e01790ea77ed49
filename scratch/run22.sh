#!/bin/bash
run() { timeout 300 python bench.py --size $1 --steps 4 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$2', $1, 'ms/step', round(d['ms_per_step'],2), 'frac', round(d['step_roofline']['frac'],3), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
run 1024 base
B2_SMINB=1 run 1024 minb3
run 512 base
B2_SMINB=1 run 512 minb3
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "512_properties" 2>&1 | tail -2
