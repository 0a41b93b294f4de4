#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_tc32.py -q -x -k "presplit" > gpurun_out/pytest_pm.log 2>&1
RC=$?
echo "pytest rc=$RC"; tail -n 12 gpurun_out/pytest_pm.log | cut -c1-300
if [ "$RC" != "0" ]; then
  SGNN_DIAG_IMPL=27 timeout 120 python scratch/tc32_diag.py > gpurun_out/tc32_diag_pm.log 2>&1; head -30 gpurun_out/tc32_diag_pm.log
fi
for IMPL in 27 0; do
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --conv-impl $IMPL --ledger gpurun_out/ledger_pm_impl$IMPL.json > gpurun_out/bench_pm_impl$IMPL.json 2> gpurun_out/bench_pm_impl$IMPL.err
  echo "impl $IMPL rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_pm_impl$IMPL.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
t=json.load(open('gpurun_out/ledger_pm_impl$IMPL.json'))['conv_times_last_profiled_step']
print([ (r['n_out'], r['cin'], r['K'], r['us']) for r in t if r['tc'] and not r['child']])"
done
