#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
SGNN_CONV_IMPL=27 timeout 200 ncu --set full --clock-control none --import-source on -k "regex:conv_tc32_pm" --launch-skip 31 -c 1 \
  -f -o gpurun_out/tc32_pm_full python scratch/one_pass.py tc32 3 > gpurun_out/ncu_pm_full.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_pm_full.log
timeout 60 ncu -i gpurun_out/tc32_pm_full.ncu-rep --page raw --csv > gpurun_out/tc32_pm_full_raw.csv 2>/dev/null
