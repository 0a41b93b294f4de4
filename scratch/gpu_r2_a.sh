#!/bin/bash
# round 2, first GPU call: unique-row kernel tests + per-convolution A/B table
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_ur.py -x -q > gpurun_out/r2a_pytest_ur.log 2>&1
echo "pytest ur rc=$?"; tail -n 15 gpurun_out/r2a_pytest_ur.log | cut -c1-300
timeout 60 python scratch/conv_table.py 20000 60000 > gpurun_out/r2a_table_ur.txt 2>&1
echo "table ur rc=$?"; cat gpurun_out/r2a_table_ur.txt | tail -60
timeout 60 python scratch/conv_table.py 1000000000 60000 > gpurun_out/r2a_table_old.txt 2>&1
echo "table old rc=$?"; head -1 gpurun_out/r2a_table_old.txt; tail -1 gpurun_out/r2a_table_old.txt
