#!/bin/bash
# Last call of the round: the full GPU suite + smoke on the final build, then configs[4] on this one GPU.
mkdir -p gpurun_out
nvidia-smi -L | head -1
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_last.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/pytest_gpu_last.txt
timeout 170 python bench.py --workload cfg5 --steps 2 --warmup 1 > gpurun_out/bench_cfg5_last.json 2> gpurun_out/bench_cfg5_last.err; tail -2 gpurun_out/bench_cfg5_last.err; cut -c1-1800 gpurun_out/bench_cfg5_last.json
