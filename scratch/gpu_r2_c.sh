#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
for D in 3 1 2 0; do
  SGNN_UR_DBG=$D timeout 45 python scratch/conv_table.py 20000 60000 > gpurun_out/r2c_table_dbg$D.txt 2>&1
  echo "dbg $D rc=$?"; head -1 gpurun_out/r2c_table_dbg$D.txt; tail -1 gpurun_out/r2c_table_dbg$D.txt
done
