#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_urc_kernel" --launch-skip 5 -c 1 -f -o gpurun_out/r2_urc_full python scratch/one_pass.py tc32 2 > gpurun_out/r2l_ncu.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv_ur_kernel" --launch-skip 57 -c 3 -f -o gpurun_out/r2_ur_full python scratch/one_pass.py tc32 2 > gpurun_out/r2l_ncu2.log 2>&1; echo rc=$?
