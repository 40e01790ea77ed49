#!/bin/bash
# round 2, run 8: __grid_constant__ parameters (no local-memory copies of the pointer tables), L2 prefetch 150
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $SIZE 2>gpurun_out/r2/run8_$tag.err | tee gpurun_out/r2/run8_$tag.json | summ "$tag"; tail -n 2 gpurun_out/r2/run8_$tag.err; }
SIZE=""
run pf150 B2_L2PF=150
run pf0 B2_L2PF=0
run pf75 B2_L2PF=75
SIZE="--size 512"
run 512 B2_X=0
SIZE="--solver ns3d.strat"
run strat1024 B2_X=0
