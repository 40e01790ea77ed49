#!/bin/bash
# round 2, run 25 (1 GPU, last GPU seconds): Euler / trapezoid / phase-shift goldens regenerated with
# coef_dealiasing = 0.9 (sensitive to the phase shifts) after the pair-renewal fix
mkdir -p gpurun_out/r2
timeout 45 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "phaseshift or euler or trapezoid" 2>&1 | tail -n 25 | tee gpurun_out/r2/run25_tests.txt
