#!/bin/bash
# round 2, run 22 (1 GPU, ~6 GPU-minutes left): full GPU suite incl. the new Euler / phase-shift /
# ns2d.strat / ns2d.bouss cases, final headline bench (with the pipelined e2e), ncu launch list of the
# bench command, ncu sections of the final strided kernels (reports stay on the box, CSV pages come back).
mkdir -p gpurun_out/r2
T0=$(date +%s)
mark() { echo "[$(( $(date +%s) - T0 )) s] $1" | tee -a gpurun_out/r2/run22_marks.txt; }
mark start
timeout 170 python -m pytest tests -q -m gpu -n 6 -p no:cacheprovider 2>&1 | tail -n 25 > gpurun_out/r2/run22_tests.txt
tail -n 6 gpurun_out/r2/run22_tests.txt
mark tests
timeout 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2/run22_bench.err > gpurun_out/r2/run22_bench.json
python - <<'EOF'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2/run22_bench.json") if l.startswith("{")][-1])
    print("bench ms/step", round(d["ms_per_step"], 2), "e2e", d.get("e2e"), "parity", (d.get("parity_check") or {}).get("max_rel_err"))
    print(" ".join(k[:6] + ":" + str(round(v["ms_per_step"], 2)) for k, v in d["kernel_classes"].items()))
except Exception as e:
    print("bench line unreadable:", e)
EOF
tail -n 3 gpurun_out/r2/run22_bench.err
mark bench
timeout 95 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fft_|xpass|xpair|rk_stage|rot_kernel|forcing|cfl' -c 400 --csv --log-file gpurun_out/r2/launches_r2_1024.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run22_ncu_l.log 2>&1
wc -l gpurun_out/r2/launches_r2_1024.csv
mark launchlist
timeout 70 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats --section LaunchStats --section SchedulerStats --clock-control none -k regex:fft_strided -s 4 -c 4 -o /tmp/prof_r2_strided_final -f python bench.py --size 512 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run22_ncu_s.log 2>&1
ncu -i /tmp/prof_r2_strided_final.ncu-rep --page raw --csv > gpurun_out/r2/r2_ncu_strided_final_512_raw.csv 2>/dev/null
wc -c gpurun_out/r2/r2_ncu_strided_final_512_raw.csv
mark sections
