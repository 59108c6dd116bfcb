#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== compute-sanitizer memcheck (subset of the parity tests)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg1 or adversarial or tie or tomb or empty or truncated or window or bloom_small or large_and_tiny or flush_matches or pipelined" > gpurun_out/memcheck.log 2>&1
echo "exit=$?"; tail -8 gpurun_out/memcheck.log | cut -c1-300; grep -c "Invalid\|out of bounds" gpurun_out/memcheck.log
echo "=== compute-sanitizer racecheck (small subset)"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg1 or tomb or flush_matches" > gpurun_out/racecheck.log 2>&1
echo "exit=$?"; tail -5 gpurun_out/racecheck.log | cut -c1-300
echo "=== bench cfg3"
timeout 600 python bench.py --workload cfg3 --no-cpu --steps 20 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'], d['pipeline_roofline'], d['e2e']['value'])"
echo "=== flush bench"
timeout 600 python tools/flush_bench.py 400000 2>&1 | tail -4
