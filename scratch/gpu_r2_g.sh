#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2s_pytest.log | cut -c1-300
for O in 0 100000 300000 0; do
timeout 300 python bench.py --no-cpu-baseline --overlap-min-rows $O --steps 20 > gpurun_out/r2s_bench_ov$O.json 2> gpurun_out/r2s_bench_ov$O.err
python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_ov$O.json')); print('overlap-min-rows $O ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
done
