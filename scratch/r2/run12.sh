#!/bin/bash
# round 2, run 12 (1 GPU): full GPU suite (forcing, observables, config 2 at 100 steps)
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25 | tee gpurun_out/r2/run12_tests.txt
