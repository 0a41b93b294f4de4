#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp pytest-tc32-tm
timeout 400 python -m pytest tests/test_gpu_tc32.py -q -x -k "tmem-operand" > gpurun_out/pytest_tc32_tm.log 2>&1
RC=$?
echo "pytest tc32 tm rc=$RC"; tail -n 15 gpurun_out/pytest_tc32_tm.log | cut -c1-300
if [ "$RC" != "0" ]; then
  stamp diag
  SGNN_DIAG_IMPL=24 timeout 200 python scratch/tc32_diag.py > gpurun_out/tc32_diag_tm.log 2>&1
  tail -n 60 gpurun_out/tc32_diag_tm.log
else
for MR in 60000 0; do
  stamp bench-tc32-impl-24-mr$MR
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-mode tc32 --conv-impl 24 --tc32-min-rows $MR --ledger gpurun_out/ledger_tc32_impl24_mr$MR.json \
    > gpurun_out/bench_tc32_impl24_mr$MR.json 2> gpurun_out/bench_tc32_impl24_mr$MR.err
  echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tc32_impl24_mr$MR.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'conv share', d['roofline']['share_of_step'])"
done
fi
stamp done
