#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp bias
timeout 400 python scratch/tc32_bias.py > gpurun_out/tc32_bias.log 2>&1
echo "bias rc=$?"; tail -n 70 gpurun_out/tc32_bias.log
stamp bench-tc32
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-mode tc32 --ledger gpurun_out/ledger_tc32.json \
  > gpurun_out/bench_tc32.json 2> gpurun_out/bench_tc32.err
echo "bench tc32 rc=$?"; cut -c1-200 gpurun_out/bench_tc32.json; tail -3 gpurun_out/bench_tc32.err
stamp bench-tc32-kg1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-mode tc32 --conv-impl 21 --ledger gpurun_out/ledger_tc32_kg1.json \
  > gpurun_out/bench_tc32_kg1.json 2> gpurun_out/bench_tc32_kg1.err
echo "bench tc32 kg1 rc=$?"; cut -c1-200 gpurun_out/bench_tc32_kg1.json
stamp ncu-launch-list
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_tc32.csv \
  python scratch/one_pass.py tc32 4 > gpurun_out/ncu_list_tc32.log 2>&1
echo "ncu list rc=$?"
stamp done
