#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp pytest-subset
timeout 500 python -m pytest tests/test_gpu_grid.py tests/test_gpu_conv.py tests/test_gpu_model.py -q -x > gpurun_out/pytest_subset2.log 2>&1
RC=$?
echo "pytest rc=$RC"; tail -n 12 gpurun_out/pytest_subset2.log | cut -c1-300
for IMPL in 0 30; do
  stamp bench-impl-$IMPL
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --conv-impl $IMPL > gpurun_out/bench_dct_impl$IMPL.json 2> gpurun_out/bench_dct_impl$IMPL.err
  echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_dct_impl$IMPL.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'conv share', d['roofline']['share_of_step'])"
done
stamp done
