#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_slab_gpu.py -x -q -m gpu 2>&1 | tail -30 | tee gpurun_out/r2/run16_tests.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "golden or forced or observables or plugin" 2>&1 | tail -5
