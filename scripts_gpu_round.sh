#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "=== pytest (flush/cfg5/async)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_host_tree.py -m gpu -q -k "flush or cfg5 or async" 2>&1 | tail -4
echo "=== bench --gpus 2 (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -4 gpurun_out/bench_n2.err | cut -c1-300; cat gpurun_out/bench_n2.json | cut -c1-3000
echo "=== reference arm under torchrun (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-400
