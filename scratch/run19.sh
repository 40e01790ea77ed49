#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 420 python -m pytest tests/test_slab_gpu.py -x -q  2>&1 | tail -6
for g in 4 8; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2953$g bench.py --gpus $g --steps 5 --warmup 3 > gpurun_out/bench1024_g$g.json 2> gpurun_out/bench1024_g$g.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench1024_g$g.json') if l.startswith('{')][-1]);print('gpus', d['n_gpus'], d['ms_per_step'], d['value'], d['step_roofline']['frac']);tot=sum(v['avg_ms']*v['launches_per_step'] for v in d['kernel_classes'].values());print('kernels', tot, 'comm+gaps', d['ms_per_step']-tot);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
grep -i "error" gpurun_out/bench1024_g$g.err | head -3
done
