#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp pytest-tc32
timeout 400 python -m pytest tests/test_gpu_tc32.py -q > gpurun_out/pytest_tc32.log 2>&1
echo "pytest tc32 rc=$?"; tail -n 6 gpurun_out/pytest_tc32.log | cut -c1-300
stamp bias
timeout 400 python scratch/tc32_bias.py > gpurun_out/tc32_bias2.log 2>&1
echo "bias rc=$?"; grep -v "^fused" gpurun_out/tc32_bias2.log | tail -n 40
for MR in 60000 0 150000; do
  stamp bench-tc32-minrows-$MR
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-mode tc32 --tc32-min-rows $MR --ledger gpurun_out/ledger_tc32_mr$MR.json \
    > gpurun_out/bench_tc32_mr$MR.json 2> gpurun_out/bench_tc32_mr$MR.err
  echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_tc32_mr$MR.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'conv share', d['roofline']['share_of_step'])"
done
stamp bench-tc32-kg1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-mode tc32 --conv-impl 21 --ledger gpurun_out/ledger_tc32_kg1.json \
  > gpurun_out/bench_tc32_kg1.json 2> gpurun_out/bench_tc32_kg1.err
python -c "
import json; d=json.load(open('gpurun_out/bench_tc32_kg1.json')); print('kg1 ms/step', d['ms_per_step'])"
stamp done
