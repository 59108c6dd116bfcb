#!/bin/bash
# A/B sweep of kernel variants on the cfg2 job (tools/tune.py) -> gpurun_out/ab.txt, then the GPU parity suite and a short bench
mkdir -p gpurun_out
nvidia-smi -L
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python tools/tune.py "$@" 2>&1 | grep -v "^\[" | tee gpurun_out/ab.txt
if [ -z "$NOTESTS" ]; then timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt; fi
if [ -n "$BENCH" ]; then
  timeout 1200 python bench.py --steps 5 --warmup 3 $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -12 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
