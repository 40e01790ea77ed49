#!/bin/bash
timeout 200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "fft or golden or pruned_steps or c2r or 2048 or config1" 2>&1 | tail -3
timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('1024 ms/step', round(d['ms_per_step'],2), 'frac', round(d['step_roofline']['frac'],3), 'e2e', round(d['e2e']['ms_per_step'],1), d['config'].get('size_note'), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"
