#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_grid.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2u_pytest.log | cut -c1-300
for O in -1 0 20000 200000 -1 0; do
timeout 300 python bench.py --no-cpu-baseline --overlap-max-rows $O --steps 20 > gpurun_out/r2u_bench_ov$O.json 2> gpurun_out/r2u_bench_ov$O.err
python -c "
import json; d=json.load(open('gpurun_out/r2u_bench_ov$O.json')); print('overlap-max-rows $O ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
done
python scratch/phases.py tc32 10 > gpurun_out/r2u_phases.txt 2>&1; grep -E "whole call|SUM|L0|L1 " gpurun_out/r2u_phases.txt
