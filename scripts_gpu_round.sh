#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu (full, 128-thread gather with 32-byte blocks)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "=== tune 128"
timeout 300 python tools/tune.py "" 2>&1 | tail -1
echo "=== tune 64"
DBEEL_LIB=$PWD/dbeel_b200/libdbeel_compact_g64.so timeout 300 python tools/tune.py "" 2>&1 | tail -1
echo "=== tune 256"
DBEEL_LIB=$PWD/dbeel_b200/libdbeel_compact_g256.so timeout 300 python tools/tune.py "" 2>&1 | tail -1
echo "=== parity 64"
DBEEL_LIB=$PWD/dbeel_b200/libdbeel_compact_g64.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
