#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "=== bench (default)"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
for v in 32 64 256 512; do
  echo "=== bench DBEEL_PARTITION_MB=$v"
  DBEEL_PARTITION_MB=$v timeout 600 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['e2e'])"
done
echo "=== bench DBEEL_PIPELINE=0"
DBEEL_PIPELINE=0 timeout 600 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['e2e'])"
echo "=== ncu full"
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:k_gather|k_merge|k_resolve|k_extract|k_emit' -s 36 -c 12 \
    -f -o gpurun_out/prof_full python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/
