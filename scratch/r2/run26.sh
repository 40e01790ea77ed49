#!/bin/bash
# round 2, run 26 (1 GPU, last GPU seconds): ns2d.strat CFL rule on the GPU path vs the oracle
mkdir -p gpurun_out/r2
timeout 30 python -m pytest tests/test_zz_ns2d_strat_cfl_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 25 | tee gpurun_out/r2/run26_tests.txt
