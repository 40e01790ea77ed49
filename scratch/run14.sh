#!/bin/bash
run() { timeout 300 python bench.py --size $1 --steps 4 --warmup 2 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$2', $1, round(d['ms_per_step'],2), 'x:', round(d['kernel_classes']['x_fused_c2r_cross_r2c']['avg_ms'],3), 'yinv:', round(d['kernel_classes']['y_inverse']['avg_ms'],3))"; }
run 1024 base
B2_XMINB=1 run 1024 minb2
B2_XCARVE=1 run 1024 carve
B2_XCARVE=1 B2_XMINB=1 run 1024 carve+minb2
run 512 base
B2_XMINB=1 run 512 minb2
B2_XCARVE=1 run 512 carve
B2_XCARVE=1 B2_XMINB=1 run 512 carve+minb2
