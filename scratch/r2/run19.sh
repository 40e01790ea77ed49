#!/bin/bash
# round 2, run 19 (2 GPUs): why are the slab y/z passes slower per byte than on one GPU?  NCCL knobs.
mkdir -p gpurun_out/r2
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['ms_per_step'],2)) for k,v in d['kernel_classes'].items()))"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { tag=$1; shift; env "$@" timeout 200 $TR bench.py --gpus 2 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e --no-parity 2>gpurun_out/r2/run19_$tag.err | tee gpurun_out/r2/run19_$tag.json | summ "$tag"; grep -E "Error|error" gpurun_out/r2/run19_$tag.err | tail -n 2; }
run nch4 NCCL_MAX_NCHANNELS=4
run nch2 NCCL_MAX_NCHANNELS=2
run ce NCCL_P2P_USE_CUDA_MEMCPY=1
run chunk4 B2_SLAB_NCHUNK=4
run chunk1 B2_SLAB_NCHUNK=1
