#!/bin/bash
# Round-end measurement pass: full GPU test suite, smoke, bench (both arms + exact mode), ncu launch list, DRAM traffic of
# every convolution launch of one pass, full capture of the largest child-mode convolution.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp pytest-all
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1
echo "pytest all rc=$?"; tail -n 8 gpurun_out/pytest_all.log | cut -c1-250
stamp smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
stamp bench-n1
timeout 400 python bench.py --ledger gpurun_out/ledger_final.json > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_final_n1.json; tail -2 gpurun_out/bench_final_n1.err
stamp bench-reference
timeout 400 python bench.py --impl reference > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_final_reference.json
for MR in 30000 120000; do
  stamp bench-minrows-$MR
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tc32-min-rows $MR > gpurun_out/bench_final_mr$MR.json 2> gpurun_out/bench_final_mr$MR.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_final_mr$MR.json')); print('min_rows $MR ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
done
stamp bench-exact
timeout 300 python bench.py --conv-mode exact --no-cpu-baseline > gpurun_out/bench_final_exact.json 2> gpurun_out/bench_final_exact.err
echo "exact rc=$?"; cut -c1-200 gpurun_out/bench_final_exact.json
stamp ncu-launch-list
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_final.csv \
  python scratch/one_pass.py tc32 4 > gpurun_out/ncu_list_final.log 2>&1
echo "ncu list rc=$?"
stamp ncu-dram
timeout 420 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:conv_" --launch-skip 153 -c 51 \
  --csv --log-file gpurun_out/conv_dram_final.csv python scratch/one_pass.py tc32 4 > gpurun_out/ncu_dram_final.log 2>&1
echo "ncu dram rc=$?"
stamp ncu-child-full
timeout 420 ncu --set full --clock-control none --import-source on -k "regex:conv_tc32_child" --launch-skip 8 -c 1 \
  -f -o gpurun_out/tc32_child_ws_full python scratch/one_pass.py tc32 3 > gpurun_out/ncu_child_full.log 2>&1
echo "ncu child rc=$?"
timeout 120 ncu -i gpurun_out/tc32_child_ws_full.ncu-rep --page raw --csv > gpurun_out/tc32_child_ws_full_raw.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/tc32_child_ws_full.ncu-rep --page source --csv > gpurun_out/tc32_child_ws_full_source.csv 2>/dev/null
stamp ncu-regular-full
# the two 16->16 convolutions of the surface head's finest level (622 k rows): 17th and 18th tensor-core launch of a pass
timeout 420 ncu --set full --clock-control none --import-source on -k "regex:conv_tc32_kernel" --launch-skip 31 -c 2 \
  -f -o gpurun_out/tc32_regular_full python scratch/one_pass.py tc32 3 > gpurun_out/ncu_regular_full.log 2>&1
echo "ncu regular rc=$?"
stamp done
