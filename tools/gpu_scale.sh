#!/bin/bash
# Multi-GPU bench lines the way the driver launches them: tools/gpu_scale.sh <N> [bench args]  ->  gpurun_out/bench_n<N>*.json
N=$1; shift
mkdir -p gpurun_out
nvidia-smi -L | head -8
tag=$(echo "$*" | tr -c 'a-zA-Z0-9' '_' | cut -c1-40)
if [ "$N" = "1" ]; then
  timeout 1700 python bench.py --gpus 1 "$@" > gpurun_out/bench_n1_$tag.json 2> gpurun_out/bench_n1_$tag.err
else
  timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N "$@" > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
fi
grep "^\[" gpurun_out/bench_n${N}_$tag.err | tail -20
tail -3 gpurun_out/bench_n${N}_$tag.err | cut -c1-400
cat gpurun_out/bench_n${N}_$tag.json | cut -c1-6000
