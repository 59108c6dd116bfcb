#!/bin/bash
# ncu --set full of selected kernels on the cfg2 job (tools/tune.py as the app): tools/gpu_prof.sh '<kernel regex>' <skip> <count> [tune specs...]
mkdir -p gpurun_out
K=$1; S=$2; C=$3; shift 3
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$K" -s $S -c $C -f -o gpurun_out/prof_sel python tools/tune.py "$@" > gpurun_out/prof_sel.log 2>&1
tail -3 gpurun_out/prof_sel.log
ls -la gpurun_out/prof_sel.ncu-rep
