#!/bin/bash
# GPU parity suite -> gpurun_out/pytest_gpu.txt  (usage: tools/gpu_tests.sh [pytest args])
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -q -x "$@" 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
