#!/bin/bash
# A/B of the lean per-entry passes of k_gather32 (DBEEL_GATHER=9..12) and of k_resolve's chunked timestamp loads, then the
# parity suite and memcheck with the variants selected
mkdir -p gpurun_out
nvidia-smi -L | head -1
timeout 600 python tools/tune.py "" "DBEEL_GATHER=13" "DBEEL_GATHER=14" "DBEEL_GATHER=15" "DBEEL_GATHER=1" "" 2>&1 | grep -v "^\[" | tee gpurun_out/ab_lean.txt
for v in 15; do
echo "=== parity suite with DBEEL_GATHER=$v"
DBEEL_GATHER=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_reader.py tests/test_host_tree.py tests/test_gpu_cfg5.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_lean_$v.txt
done
echo "=== memcheck"
for v in 15; do
DBEEL_GATHER=$v timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import numpy as np
from dbeel_b200 import capi, workloads as W
import oracle
seed = bytes(range(32))
for cfg in (W.scaled(W.CFG2, 30000), W.scaled(W.CFG3, 4000)):
    runs = W.make_merge_runs(cfg)
    exp = oracle.compact(runs, cfg.keep_tombstones, seed=seed, bloom_min_size=100000)
    e = capi.Engine(0); g = e.compact(runs, cfg.keep_tombstones, seed=seed, bloom_min_size=100000); e.close()
    print(cfg.name, all(np.array_equal(a, b) for a, b in zip(g[:3], exp[:3])))
" 2>&1 | tail -4 | tee gpurun_out/sanitizer_lean_$v.txt
done
