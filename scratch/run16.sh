#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_gpu.py -x -q 2>&1 | tail -8
for n in 1024; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --size $n --steps 5 --warmup 3 > gpurun_out/bench${n}_g2.json 2> gpurun_out/bench${n}_g2.err; python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench${n}_g2.json') if l.startswith('{')][-1]);print($n, 'gpus', d['n_gpus'], d['ms_per_step'], d['value'], d['step_roofline']['frac']);tot=sum(v['avg_ms']*v['launches_per_step'] for v in d['kernel_classes'].values());print('kernels', tot, 'comm+gaps', d['ms_per_step']-tot);[print(k, round(v['avg_ms'],3), round(v['frac'],3)) for k,v in d['kernel_classes'].items()]"
tail -3 gpurun_out/bench${n}_g2.err
done
