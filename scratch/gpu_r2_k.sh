#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_ur.py -x -q -k "urc or plan" 2>&1 | tail -8
timeout 100 python scratch/conv_table.py 1000 60000 > gpurun_out/r2k_table.txt 2>&1; head -1 gpurun_out/r2k_table.txt; grep " 1  1 |" gpurun_out/r2k_table.txt; tail -1 gpurun_out/r2k_table.txt
timeout 100 python scratch/ur_hang_hunt.py 200 | tail -3
