#!/bin/bash
# round 2, run 6: cp.async-prefetch persistent strided passes (B2_STMA=2) vs LDG (0)
mkdir -p gpurun_out/r2
B2_STMA=2 timeout 600 python -m pytest tests -x -q -m gpu -k "512 or 2048 or golden" 2>&1 | tail -5
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $SIZE 2>gpurun_out/r2/run6_$tag.err | tee gpurun_out/r2/run6_$tag.json | summ "$tag"; tail -n 2 gpurun_out/r2/run6_$tag.err; }
SIZE=""
run cpasync B2_STMA=2
run ldg B2_STMA=0
SIZE="--size 512"
run 512cpasync B2_STMA=2
