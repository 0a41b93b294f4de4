#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_tc32.py tests/test_gpu_scene.py -m gpu -x -q > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2x_pytest.log | cut -c1-300
python scratch/fmt_cost.py
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2x_bench.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
