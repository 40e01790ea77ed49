#!/bin/bash
timeout 200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "pruning_is_not_used or pruned_steps or golden or cfl" 2>&1 | tail -3
timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('1024 ms/step', round(d['ms_per_step'],2), 'e2e', d['e2e'])"
