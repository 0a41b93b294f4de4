#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp pytest-subset
timeout 500 python -m pytest tests/test_gpu_model.py tests/test_gpu_tc32.py -q -x -k "not warp-specialised and not tmem-operand and not child-single-role" > gpurun_out/pytest_subset.log 2>&1
RC=$?
echo "pytest rc=$RC"; tail -n 12 gpurun_out/pytest_subset.log | cut -c1-300
for IMPL in 0 26; do
  stamp bench-impl-$IMPL
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --conv-impl $IMPL --ledger gpurun_out/ledger_kg2_impl$IMPL.json \
    > gpurun_out/bench_kg2_impl$IMPL.json 2> gpurun_out/bench_kg2_impl$IMPL.err
  echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_kg2_impl$IMPL.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'conv share', d['roofline']['share_of_step'])
t=json.load(open('gpurun_out/ledger_kg2_impl$IMPL.json'))['conv_times_last_profiled_step']
print([ (r['n_out'], r['cin'], r['us']) for r in t if r['tc'] and r['cin'] in (26,30)])"
done
stamp done
