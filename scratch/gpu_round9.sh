#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp pytest-tc32
timeout 500 python -m pytest tests/test_gpu_tc32.py -q -x > gpurun_out/pytest_tc32_elect.log 2>&1
RC=$?
echo "pytest rc=$RC"; tail -n 15 gpurun_out/pytest_tc32_elect.log | cut -c1-300
if [ "$RC" = "0" ]; then
for IMPL in 0 23 24; do
  stamp bench-impl-$IMPL
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-impl $IMPL --ledger gpurun_out/ledger_el_impl$IMPL.json \
    > gpurun_out/bench_el_impl$IMPL.json 2> gpurun_out/bench_el_impl$IMPL.err
  echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_el_impl$IMPL.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'conv share', d['roofline']['share_of_step'])
t=json.load(open('gpurun_out/ledger_el_impl$IMPL.json'))['conv_times_last_profiled_step']
print([ (r['n_out'], r['cin'], r['us']) for r in t if r['tc'] and r['n_out']>150000])"
done
fi
stamp done
