#!/bin/bash
# round 2, run 20 (2 GPUs): shift-based mappers -- slab parity (2-GPU cases) + bench
mkdir -p gpurun_out/r2
timeout 400 python -m pytest tests/test_slab_gpu.py -x -q -m gpu -k "test_slab_matches_reference_golden and (2-1 or 2-2)" 2>&1 | tail -4 | tee gpurun_out/r2/run20_tests.txt
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), 'parity', d.get('parity_check',{}).get('max_rel_err'), ' '.join(k[:6]+':'+str(round(v['ms_per_step'],2)) for k,v in d['kernel_classes'].items()))"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2/run20_g2.err | tee gpurun_out/r2/run20_g2.json | summ "ns3d 1024 x2 shifts"
grep -E "Error|error" gpurun_out/r2/run20_g2.err | tail -n 3
