#!/bin/bash
for v in 0 1 2 3 4; do B2_SVAR=$v python scratch/micro_strided.py 1024 3 2>&1 | grep SVAR; done
for v in 0 1 2 3 4; do B2_SVAR=$v python scratch/micro_strided.py 512 6 2>&1 | grep SVAR; done
