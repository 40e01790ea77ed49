#!/bin/bash
# round 2, run 11 (1 GPU): aliased buffers, parity_check in bench, strat 1024^3
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2/run11_tests.txt
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), 'parity', d.get('parity_check',{}).get('max_rel_err'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), ' '.join(k[:6]+':'+str(round(v['ms_per_step'],2)) for k,v in d['kernel_classes'].items()))"; }
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2/run11_g1.err | tee gpurun_out/r2/run11_g1.json | summ "ns3d 1024 x1"
tail -n 3 gpurun_out/r2/run11_g1.err
timeout 600 python bench.py --solver ns3d.strat --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2/run11_s1.err | tee gpurun_out/r2/run11_s1.json | summ "strat 1024 x1"
tail -n 3 gpurun_out/r2/run11_s1.err
