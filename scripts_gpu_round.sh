#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== tune"
timeout 1200 python tools/tune.py "" DBEEL_GATHER_TUNE=4 DBEEL_GATHER_TUNE=6 DBEEL_GATHER_TUNE=14 DBEEL_GATHER_TUNE=15 DBEEL_GATHER_TUNE=16 \
   DBEEL_GATHER=0 DBEEL_MERGE=2 DBEEL_MERGE=1 2>&1 | tail -12
echo "=== pytest -m gpu (quick parity on default)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4
