#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 200 python scratch/ur_hang_hunt.py 300 2>&1 | tail -3
bash scratch/gpu_r2_a.sh
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:conv_ur_kernel" --launch-skip 21 -c 3 \
  -f -o gpurun_out/r2_ur_full python scratch/one_pass.py tc32 2 > gpurun_out/r2b_ncu_full.log 2>&1
echo "ncu full rc=$?"
