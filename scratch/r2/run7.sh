#!/bin/bash
# round 2, run 7: L2 prefetch-ahead in the LDG strided kernels, distance sweep
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -x -q -m gpu -k "512 or 2048 or golden or config2" 2>&1 | tail -3
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $SIZE 2>gpurun_out/r2/run7_$tag.err | tee gpurun_out/r2/run7_$tag.json | summ "$tag"; tail -n 2 gpurun_out/r2/run7_$tag.err; }
SIZE=""
for d in 0 150 300 600 1200 2400 4800; do run pf$d B2_L2PF=$d; done
SIZE="--size 512"
for d in 0 600 2400; do run 512pf$d B2_L2PF=$d; done
