#!/bin/bash
for nc in 4 8; do
B2_SLAB_NCHUNK=$nc timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$nc bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);tot=sum(v['avg_ms']*v['launches_per_step'] for v in d['kernel_classes'].values());print('nchunk $nc gpus', d['n_gpus'], round(d['ms_per_step'],2), 'kernels', round(tot,1), 'exposed', round(d['ms_per_step']-tot,1))"
done
