#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== pytest -m gpu (default variants)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "=== bench (default)"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
for v in "0 0" "0 1" "3 0" "3 1" "2 1"; do
  set -- $v
  echo "=== bench DBEEL_GATHER=$1 DBEEL_MERGE=$2"
  DBEEL_GATHER=$1 DBEEL_MERGE=$2 timeout 600 python bench.py --no-cpu --steps 30 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stage_ms'], d['roofline']['frac'])"
done
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches.csv
echo "=== ncu full"
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:k_gather|k_merge|k_resolve|k_extract|k_emit' -s 36 -c 12 \
    -f -o gpurun_out/prof_full python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/
