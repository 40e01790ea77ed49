#!/bin/bash
# round 2, run 14 (1 GPU): evidence run -- default bench + reference arm, ncu launch list, ncu --set full of
# the x pass and the four strided passes (final kernels), compute-sanitizer racecheck/memcheck of the smoke
mkdir -p gpurun_out/r2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2/run14_bench.json 2> gpurun_out/r2/run14_bench.err
tail -c 600 gpurun_out/r2/run14_bench.json; echo
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2/run14_ref.json 2> gpurun_out/r2/run14_ref.err
tail -c 400 gpurun_out/r2/run14_ref.json; echo
timeout 600 python bench.py --solver ns3d.strat --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2/run14_strat.json 2> gpurun_out/r2/run14_strat.err
timeout 600 python bench.py --size 512 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2/run14_512.json 2> gpurun_out/r2/run14_512.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fft_|xpass|rk_stage|rot_kernel|forcing' -s 30 -c 104 --csv --log-file gpurun_out/r2/launches_r2_1024.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xpass_pair -s 4 -c 1 -o gpurun_out/r2/prof_r2_xpair_final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_x.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_strided -s 4 -c 4 -o gpurun_out/r2/prof_r2_strided_final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_s.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'rk_stage|observables' -s 1 -c 2 -o gpurun_out/r2/prof_r2_rk_final -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2/run14_ncu_r.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/sanitizer_racecheck_smoke.log 2>&1
tail -n 4 gpurun_out/r2/sanitizer_racecheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/sanitizer_memcheck_smoke.log 2>&1
tail -n 4 gpurun_out/r2/sanitizer_memcheck_smoke.log
ls -la gpurun_out/r2/ | tail -n 12
