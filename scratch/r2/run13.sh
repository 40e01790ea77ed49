#!/bin/bash
# round 2, run 13 (2 GPUs): slab parity (native NCCL + torch.distributed orchestration, lean buffers), 2-GPU benches
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_slab_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r2/run13_tests.txt
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), 'parity', d.get('parity_check',{}).get('max_rel_err'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), ' '.join(k[:6]+':'+str(round(v['ms_per_step'],2)) for k,v in d['kernel_classes'].items()))"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2/run13_g2.err | tee gpurun_out/r2/run13_g2.json | summ "ns3d 1024 x2 native"
grep -E "Error|error" gpurun_out/r2/run13_g2.err | tail -n 3
B2_SLAB_NATIVE=0 timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2/run13_g2t.err | tee gpurun_out/r2/run13_g2t.json | summ "ns3d 1024 x2 torch a2a"
grep -E "Error|error" gpurun_out/r2/run13_g2t.err | tail -n 3
timeout 600 $TR bench.py --gpus 2 --solver ns3d.strat --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2/run13_s2.err | tee gpurun_out/r2/run13_s2.json | summ "strat 1024 x2"
grep -E "Error|error" gpurun_out/r2/run13_s2.err | tail -n 3
