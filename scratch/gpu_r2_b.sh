#!/bin/bash
# ncu: full capture of the unique-row convolutions of one pass (17 launches)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_ur_kernel" --launch-skip 17 -c 17 \
  -f -o gpurun_out/r2_ur_full python scratch/one_pass.py tc32 2 > gpurun_out/r2b_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r2b_ncu_full.log
