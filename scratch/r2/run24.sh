#!/bin/bash
# round 2, run 24 (1 GPU, last GPU seconds): checkpoint / restart GPU test
mkdir -p gpurun_out/r2
timeout 70 python -m pytest tests/test_zz_checkpoint_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 25 | tee gpurun_out/r2/run24_tests.txt
