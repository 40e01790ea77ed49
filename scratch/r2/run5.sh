#!/bin/bash
# round 2, run 5: TMA-pipelined strided passes, 128-byte aligned boxes, TK = 4 / 8
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests -x -q -m gpu -k "512 or 2048 or golden" 2>&1 | tail -5
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $SIZE 2>gpurun_out/r2/run5_$tag.err | tee gpurun_out/r2/run5_$tag.json | summ "$tag"; tail -n 2 gpurun_out/r2/run5_$tag.err; }
SIZE=""
run tma4 B2_X=0
run tma8 B2_STMA_TK=8
run tma4_l2none B2_STMA_L2=0
SIZE="--size 512"
run 512tma4 B2_X=0
run 512tma8 B2_STMA_TK=8
run 512ldg B2_STMA=0
