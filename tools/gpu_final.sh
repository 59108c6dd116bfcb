#!/bin/bash
# Round-end validation of the default build on one B200: GPU suite, smoke, bench line, ncu launch list + full capture of one cfg2 job, memcheck.
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -1
echo "=== pytest -m gpu (full)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default)"
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
echo "=== ncu full (one cfg2 job, every kernel of the pipeline)"
timeout 900 ncu --set full --clock-control none --import-source on -s 57 -c 19 -f -o gpurun_out/prof_full python tools/tune.py "" > gpurun_out/prof_full.log 2>&1; tail -2 gpurun_out/prof_full.log
echo "=== ncu launch list (DBEEL_PIPELINE=0 keeps the host-path part of bench.py to one job per call)"
DBEEL_PIPELINE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
grep -c . gpurun_out/launches.csv
echo "=== memcheck: scaled cfg2 / cfg3 jobs through the default path"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import numpy as np
from dbeel_b200 import capi, workloads as W
import oracle
seed = bytes(range(32))
for cfg in (W.scaled(W.CFG2, 30000), W.scaled(W.CFG3, 4000)):
    runs = W.make_merge_runs(cfg)
    exp = oracle.compact(runs, cfg.keep_tombstones, seed=seed, bloom_min_size=100000)
    e = capi.Engine(0); g = e.compact(runs, cfg.keep_tombstones, seed=seed, bloom_min_size=100000); e.close()
    print(cfg.name, all(np.array_equal(a, b) for a, b in zip(g[:3], exp[:3])))
" 2>&1 | tail -5 | tee gpurun_out/sanitizer.txt
ls -la gpurun_out/ | tail -8
