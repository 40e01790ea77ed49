#!/bin/bash
# round 2, run 17 (8 GPUs): slab parity on 4 / 8 ranks, ns3d 2048^3 x8 (config 5 size), ns3d / strat 1024^3 x8 and x4
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2/run17_smi.txt
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), 'parity', d.get('parity_check',{}).get('max_rel_err'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), ' '.join(k[:6]+':'+str(round(v['ms_per_step'],2)) for k,v in d['kernel_classes'].items()))"; }
run() { tag=$1; n=$2; shift; shift; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu-baseline "$@" 2>gpurun_out/r2/run17_$tag.err | tee gpurun_out/r2/run17_$tag.json | summ "$tag"; grep -E "Error|error|OutOfMemory" gpurun_out/r2/run17_$tag.err | tail -n 3; }
run ns3d2048_g8 8 --size 2048 --steps 3 --warmup 2 --no-e2e
run ns3d1024_g8 8 --steps 5 --warmup 3
run strat1024_g8 8 --solver ns3d.strat --steps 5 --warmup 3 --no-e2e
run strat1024_g4 4 --solver ns3d.strat --steps 5 --warmup 3 --no-e2e
run ns3d1024_g4 4 --steps 5 --warmup 3 --no-e2e
timeout 400 python -m pytest tests/test_slab_gpu.py -x -q -m gpu -k "4-2 or 8-1 or 8-2" 2>&1 | tail -5 | tee gpurun_out/r2/run17_tests.txt
