#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_slab_gpu.py -x -q -k "(ns3d_16x16x16_rk4 or strat_16x8x32_rk2) and (4-2-cyclic or 8-1-cyclic or 8-2-block)" 2>&1 | tail -4
bench() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2954$1 bench.py --gpus $1 --steps 5 --warmup 3 > gpurun_out/bench1024_g$1_$2.json 2> gpurun_out/bench1024_g$1_$2.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench1024_g$1_$2.json') if l.startswith('{')][-1]);tot=sum(v['avg_ms']*v['launches_per_step'] for v in d['kernel_classes'].values());print('$2 gpus', d['n_gpus'], round(d['ms_per_step'],2), '%.3e'%d['value'], 'kernels', round(tot,1), 'exposed', round(d['ms_per_step']-tot,1), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; grep -i "error" gpurun_out/bench1024_g$1_$2.err | head -2; }
bench 8 auto
B2_SLAB_KY=block bench 8 block
bench 4 auto
