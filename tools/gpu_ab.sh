#!/bin/bash
# A/B sweep of kernel variants on the cfg2 job (tools/tune.py) -> gpurun_out/ab.txt, then the GPU parity suite and a short bench
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python tools/tune.py "$@" 2>&1 | grep -v "^\[" | tee gpurun_out/ab.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
if [ -n "$BENCH" ]; then
  timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -12 gpurun_out/bench.err; cat gpurun_out/bench.json
fi
