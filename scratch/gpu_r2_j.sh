#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_grid.py tests/test_gpu_ur.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2w_pytest.log | cut -c1-300
for U in 0 100000 0 100000; do
timeout 300 python bench.py --no-cpu-baseline --ur12-min-rows $U --steps 20 > gpurun_out/r2w_bench_u$U.json 2> gpurun_out/r2w_bench_u$U.err
python -c "
import json; d=json.load(open('gpurun_out/r2w_bench_u$U.json')); print('ur12-min-rows $U ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
done
python scratch/phases.py tc32 10 > gpurun_out/r2w_phases.txt 2>&1; grep -E "whole|SUM|L3|encoder" gpurun_out/r2w_phases.txt
