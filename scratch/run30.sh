#!/bin/bash
export B2_XTWC=1
timeout 120 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "golden or pruned_steps or config2 or 2048 or 512_prop" 2>&1 | tail -2
timeout 120 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('XTWC 1024 ms/step', round(d['ms_per_step'],2), 'frac', round(d['step_roofline']['frac'],3), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"
