#!/bin/bash
# One gpurun call that validates a build end to end on a B200:
#   /usr/local/graft/bin/gpurun --timeout 3000 -- './tools/gpu_round.sh'
# parity tests, smoke, the bench line, the reference arm, the ncu launch list and one --set full capture
# of the hot kernels (summarise here with tools/ncu_summary.py / tools/sass_hotspots.py, copy into profiles/).
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== pytest -m gpu (full)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default)"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "=== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-600 gpurun_out/bench_ref.json
echo "=== lookups (N2), flushes / WAL replay (N1, N4), many compactions (N1), cfg5 pipeline"
timeout 600 python tools/lookup_bench.py 2>/dev/null | tee gpurun_out/lookup_bench.txt | tail -9
timeout 600 python tools/flush_bench.py 400000 2>/dev/null | tee gpurun_out/flush_bench.txt | tail -8
timeout 600 python tools/many_bench.py 32 2>/dev/null | tee gpurun_out/many_bench.txt | tail -5
timeout 600 python tools/cfg5_bench.py 1500000 2>/dev/null | tee gpurun_out/cfg5_bench.txt | tail -5
echo "=== cfg3 (BASELINE configs[2]) with the oracle comparison"
timeout 900 python bench.py --workload cfg3 --steps 20 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; tail -1 gpurun_out/bench_cfg3.err; cut -c1-400 gpurun_out/bench_cfg3.json
echo "=== randomised parity sweep"
timeout 600 python tools/parity_sweep.py 1400 40000 2>&1 | tail -2
echo "=== ncu launch list (DBEEL_PIPELINE=0 keeps the host-path part of bench.py to one job per call)"
DBEEL_PIPELINE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches.csv
echo "=== ncu full"
DBEEL_PIPELINE=0 timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:k_gather|k_merge_tma|k_resolve|k_extract|k_emit|k_merge_partition' -s 30 -c 10 \
    -f -o gpurun_out/prof_full python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
