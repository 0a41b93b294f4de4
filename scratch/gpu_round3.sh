#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp ncu-full-tc32
SKIP=$((3 * 43 + 33))
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:conv_tc32" --launch-skip $SKIP -c 3 \
  -f -o gpurun_out/tc32_full python scratch/one_pass.py tc32 4 > gpurun_out/ncu_full_tc32.log 2>&1
echo "ncu full rc=$?"; tail -5 gpurun_out/ncu_full_tc32.log
timeout 120 ncu -i gpurun_out/tc32_full.ncu-rep --page raw --csv > gpurun_out/tc32_full_raw.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/tc32_full.ncu-rep --page source --csv > gpurun_out/tc32_full_source.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/tc32_full.ncu-rep --page details > gpurun_out/tc32_full_details.txt 2>/dev/null
ls -la gpurun_out/tc32_full*
stamp done
