#!/bin/bash
# Row N3 on a GPU box: the streaming tests, the file-fed job at a few thread settings (mmap'ed outputs vs pwrite), then the bench.
mkdir -p gpurun_out
nvidia-smi -L | head -2; nproc; uname -r
echo "=== pytest -m gpu (host tree + stream)"
timeout 600 python -m pytest tests/test_host_tree.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
for spec in "16 3 1" "8 3 1" "32 4 1" "16 3 0"; do
  set -- $spec
  echo "=== file-fed job: $1 threads, ring $2, mmap outputs $3"
  DBEEL_TRACE=1 DBEEL_IO_THREADS=$1 DBEEL_STREAM_RING=$2 DBEEL_STREAM_MMAP=$3 timeout 300 python tools/file_fed.py > gpurun_out/file_fed_t$1_r$2_m$3.json 2> gpurun_out/file_fed_t$1_r$2_m$3.err
  tail -1 gpurun_out/file_fed_t$1_r$2_m$3.err | cut -c1-300; cat gpurun_out/file_fed_t$1_r$2_m$3.json
done
if [ -n "$BENCH" ]; then
  echo "=== bench (default)"
  timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -6 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
fi
