#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_slab_gpu.py -x -q -k "ns3d_16x16x16_rk4 or strat_16x8x32_rk2" 2>&1 | tail -3
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench1024_g2_auto.json 2> gpurun_out/bench1024_g2_auto.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench1024_g2_auto.json') if l.startswith('{')][-1]);tot=sum(v['avg_ms']*v['launches_per_step'] for v in d['kernel_classes'].values());print('gpus', d['n_gpus'], round(d['ms_per_step'],2), '%.3e'%d['value'], 'kernels', round(tot,1), 'exposed', round(d['ms_per_step']-tot,1), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"
