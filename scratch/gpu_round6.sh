#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp ncu-list
# count matching launches per pass first (cheap): names + durations of one tc32 pass
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:conv_tc32" -c 400 --csv --log-file gpurun_out/tc32_only_list.csv \
  python scratch/one_pass.py tc32 2 > gpurun_out/ncu_l.log 2>&1
N=$(grep -c "conv_tc32" gpurun_out/tc32_only_list.csv)
echo "matching launches in 2 passes: $N"
PER=$((N / 2))
stamp ncu-full
# last 10 tensor-core convolutions of the 3rd pass: last refinement level's child conv + surface head
SKIP=$((2 * PER + PER - 10))
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:conv_tc32" --launch-skip $SKIP -c 4 \
  -f -o gpurun_out/tc32v2_full python scratch/one_pass.py tc32 3 > gpurun_out/ncu_full_tc32v2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_tc32v2.log
timeout 120 ncu -i gpurun_out/tc32v2_full.ncu-rep --page raw --csv > gpurun_out/tc32v2_full_raw.csv 2>/dev/null
timeout 120 ncu -i gpurun_out/tc32v2_full.ncu-rep --page source --csv > gpurun_out/tc32v2_full_source.csv 2>/dev/null
ls -la gpurun_out/tc32v2*
stamp done
