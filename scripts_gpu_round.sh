#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu (full)"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "=== tune"
timeout 600 python tools/tune.py "" 2>&1 | tail -2
