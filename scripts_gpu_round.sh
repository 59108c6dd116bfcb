#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== tune default"
timeout 300 python tools/tune.py "" 2>&1 | tail -1
echo "=== tune resolve 128 threads"
DBEEL_LIB=$PWD/dbeel_b200/libdbeel_compact_r128.so timeout 300 python tools/tune.py "" 2>&1 | tail -1
echo "=== parity resolve 128"
DBEEL_LIB=$PWD/dbeel_b200/libdbeel_compact_r128.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
