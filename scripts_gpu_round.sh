#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== pytest -m gpu (full, durations)"
timeout 900 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14
echo "=== tune"
timeout 600 python tools/tune.py "" DBEEL_MERGE=0 DBEEL_NARROW=0 2>&1 | tail -4
echo "=== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | cut -c1-900
