#!/bin/bash
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_slab_gpu.py -x -q -m gpu -k "2-1-block-False-True-ns3d_16" 2>&1 | tail -60 | tee gpurun_out/r2/run15_tests.txt
