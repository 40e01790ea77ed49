#!/bin/bash
timeout 400 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "cfl or golden or pruned or 2048 or nan" 2>&1 | tail -3
timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('1024 ms/step', round(d['ms_per_step'],2), 'frac', round(d['step_roofline']['frac'],3), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"
