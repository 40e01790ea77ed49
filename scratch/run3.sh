#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench512.json 2> gpurun_out/bench512.err; tail -c 3000 gpurun_out/bench512.json; tail -5 gpurun_out/bench512.err
python bench.py --n 1024 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench1024.json 2> gpurun_out/bench1024.err; tail -c 3000 gpurun_out/bench1024.json; tail -5 gpurun_out/bench1024.err
python bench.py --n 256 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench256.json 2> gpurun_out/bench256.err; tail -c 1500 gpurun_out/bench256.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_512.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
