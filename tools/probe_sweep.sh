#!/bin/bash
# sweep of tools/bin/tma_probe (payload gather design probe); output -> gpurun_out/tma_probe.txt
mkdir -p gpurun_out
P=tools/bin/tma_probe
{
nvidia-smi -L
# selftest <layout 0 flat1d | 1 rows256 | 2 overlapped rows> <box0> <box1> <byte offset> <which 0 param | 1 param array | 2 global> <op 0 load | 1 store>
for t in "0 256 1 0 0 0" "0 256 1 3 0 0" "0 48 1 1001 0 0" "1 256 2 1280 0 0" "2 256 1 0 0 0" "2 256 1 3 0 0" "2 256 4 2381 0 0" "2 48 1 2381 0 0" \
         "2 256 1 3 1 0" "2 256 1 3 2 0" "0 256 1 3 1 0" "0 256 1 3 2 0" \
         "0 256 1 0 0 1" "0 256 1 3 0 1" "0 48 1 1001 0 1" "2 256 1 3 0 1" "2 256 4 2381 0 1" "2 256 4 2381 1 1"; do
  timeout 60 $P selftest $t 2>&1 | tail -1
done
for cfg in "305 4000000 4 4 1" "305 4000000 4 4 2" "305 4000000 8 4 1" "305 4000000 2 4 4" "305 4000000 4 6 1" "305 4000000 8 3 2" "305 4000000 4 4 1 0" \
           "113 8000000 4 4 2" "561 3000000 4 4 1" "1073 2000000 4 4 1" "1073 2000000 2 4 1" "4200 500000 2 3 1"; do
  echo "--- $cfg"
  timeout 120 $P $cfg 2>&1 | tail -3
done
} > gpurun_out/tma_probe.txt 2>&1
cat gpurun_out/tma_probe.txt
