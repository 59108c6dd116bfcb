#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== flush tests"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_host_tree.py -m gpu -q -x -k "flush or cfg5" 2>&1 | tail -25 | cut -c1-300
echo "=== flush bench"
timeout 600 python tools/flush_bench.py 400000 2>&1 | tail -7
