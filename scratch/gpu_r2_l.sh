#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r2v2_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2v2_pytest.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2v2_bench.json 2> gpurun_out/r2v2_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2v2_bench.json')); r=d['roofline']; print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step']); print([p for p in r['phases_us_one_pass'] if 'dense' in p[0]])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dense_ --csv --log-file gpurun_out/r2v2_dense.csv python scratch/one_pass.py tc32 1 > /dev/null 2>&1
grep -E "dense_" gpurun_out/r2v2_dense.csv | awk -F'","' '{print $5, $NF}' | cut -c1-90 | tail -10
