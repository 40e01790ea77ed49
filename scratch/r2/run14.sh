#!/bin/bash
# round 2, run 14 (1 GPU): evidence run -- default bench + reference arm, slab code path on one rank
# (class times + ncu), ncu launch list, ncu --set full of the x pass / strided passes / epilogue,
# compute-sanitizer racecheck + memcheck of the smoke
mkdir -p gpurun_out/r2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/run14_bench.json 2> gpurun_out/r2/run14_bench.err
tail -c 300 gpurun_out/r2/run14_bench.json; echo
summ() { python -c "
import json,sys;d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]);print('$1', 'ms/step', round(d['ms_per_step'],2), ' '.join(k[:6]+':'+str(round(v['ms_per_step'],2)) for k,v in d['kernel_classes'].items()))"; }
B2_BENCH_FORCE_SLAB=1 timeout 400 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-parity 2>gpurun_out/r2/run14_slab1.err | tee gpurun_out/r2/run14_slab1.json | summ "slab path on 1 rank"
grep -E "Error|error" gpurun_out/r2/run14_slab1.err | tail -n 3
B2_BENCH_FORCE_SLAB=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_strided -s 42 -c 12 -o gpurun_out/r2/prof_r2_slab_strided -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_slab.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fft_|xpass|rk_stage|rot_kernel|forcing' -s 30 -c 104 --csv --log-file gpurun_out/r2/launches_r2_1024.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xpass_pair -s 4 -c 1 -o gpurun_out/r2/prof_r2_xpair_final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_x.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_strided -s 4 -c 4 -o gpurun_out/r2/prof_r2_strided_final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_s.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'rk_stage' -s 1 -c 1 -o gpurun_out/r2/prof_r2_rk_final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_r.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/sanitizer_racecheck_smoke.log 2>&1
tail -n 3 gpurun_out/r2/sanitizer_racecheck_smoke.log
timeout 400 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/sanitizer_memcheck_smoke.log 2>&1
tail -n 3 gpurun_out/r2/sanitizer_memcheck_smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/run14_ref.json 2> gpurun_out/r2/run14_ref.err
tail -c 300 gpurun_out/r2/run14_ref.json; echo
ls -la gpurun_out/r2/ | grep -E "prof_r2|sanitizer|launches_r2"
