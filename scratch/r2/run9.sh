#!/bin/bash
# round 2, run 9: z-pass tile variants with L2 prefetch; ncu of the staged x pass
mkdir -p gpurun_out/r2
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['avg_ms'],2)) for k,v in d['kernel_classes'].items()))"; }
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline $SIZE 2>gpurun_out/r2/run9_$tag.err | tee gpurun_out/r2/run9_$tag.json | summ "$tag"; tail -n 2 gpurun_out/r2/run9_$tag.err; }
SIZE=""
run base B2_X=0
run svar2 B2_SVAR=2
run svar1 B2_SVAR=1
run pf300z B2_L2PF=300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xpass_pair -s 4 -c 1 -o gpurun_out/r2/prof_xpair_v1 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/run9_ncu_x.log 2>&1
ls -la gpurun_out/r2/*.ncu-rep
