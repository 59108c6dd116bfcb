#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== tune"
timeout 900 python tools/tune.py "" DBEEL_NARROW=1 2>&1 | tail -4
echo "=== pytest -m gpu (narrow loads)"
DBEEL_NARROW=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
echo "=== ncu narrow"
DBEEL_NARROW=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:k_extract|k_resolve' -s 6 -c 3 --csv --log-file gpurun_out/narrow.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
cat gpurun_out/narrow.csv | grep -E "k_extract|k_resolve" | cut -d, -f5,13- | head -12
