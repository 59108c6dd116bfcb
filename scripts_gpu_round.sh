#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== pytest (TMA merge)"
DBEEL_MERGE=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
echo "=== tune"
timeout 900 python tools/tune.py "" DBEEL_MERGE=3 DBEEL_EXTRACT_TUNE=16 DBEEL_EXTRACT_TUNE=18 DBEEL_EXTRACT_TUNE=25 2>&1 | tail -6
