#!/bin/bash
# round 2, run 23 (1 GPU, last ~2.5 GPU-minutes): the two RK2_phaseshift_exact strat cases that failed in
# run 22 (the reference aliases output and input there) + the new strat RK2_phaseshift_random golden
mkdir -p gpurun_out/r2
timeout 120 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "strat_16x16x16_rk2_phaseshift" 2>&1 | tail -n 15 | tee gpurun_out/r2/run23_tests.txt
