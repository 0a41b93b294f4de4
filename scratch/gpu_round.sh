#!/bin/bash
# One gpurun call: tc32 probes + tests, benches (exact vs tc32), ncu launch list + full capture.  Everything lands in
# gpurun_out/.  Every step is under its own timeout so a hung kernel cannot eat the box.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }

stamp diag
timeout 300 python scratch/tc32_diag.py > gpurun_out/tc32_diag.log 2>&1
echo "diag rc=$?"
tail -n 60 gpurun_out/tc32_diag.log

stamp pytest-tc32
timeout 600 python -m pytest tests/test_gpu_tc32.py -q -x > gpurun_out/pytest_tc32.log 2>&1
TC_RC=$?
echo "pytest tc32 rc=$TC_RC"
tail -n 25 gpurun_out/pytest_tc32.log

stamp bench-exact
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-mode exact --ledger gpurun_out/ledger_exact.json \
  > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err
echo "bench exact rc=$?"; cat gpurun_out/bench_exact.json | cut -c1-400

MODE=exact
KREGEX='regex:conv_ro_kernel|conv_child'
if [ "$TC_RC" = "0" ] || [ "$FORCE_TC" = "1" ]; then
  MODE=tc32
  KREGEX='regex:conv_tc32'
  stamp bench-tc32
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-mode tc32 --ledger gpurun_out/ledger_tc32.json \
    > gpurun_out/bench_tc32.json 2> gpurun_out/bench_tc32.err
  echo "bench tc32 rc=$?"; cat gpurun_out/bench_tc32.json | cut -c1-400
  stamp bench-tc32-kg1
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --conv-mode tc32 --conv-impl 21 \
    > gpurun_out/bench_tc32_kg1.json 2> gpurun_out/bench_tc32_kg1.err
  echo "bench tc32 kg1 rc=$?"; cat gpurun_out/bench_tc32_kg1.json | cut -c1-300
fi

stamp ncu-launch-list
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$MODE.csv \
  python scratch/one_pass.py $MODE 4 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"

stamp ncu-full
# the big convolutions of the 4th pass (last refinement level + surface head); 3 warm-up passes are skipped
if [ "$MODE" = "tc32" ]; then SKIP=$((3 * 43 + 24)); CNT=19; else SKIP=$((3 * 51 + 30)); CNT=21; fi
timeout 600 ncu --set full --clock-control none --import-source on -k "$KREGEX" --launch-skip $SKIP -c $CNT \
  -f -o gpurun_out/conv_full_$MODE python scratch/one_pass.py $MODE 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
timeout 120 ncu -i gpurun_out/conv_full_$MODE.ncu-rep --page raw --csv > gpurun_out/conv_full_${MODE}_raw.csv 2>/dev/null
ls -la gpurun_out | head -40

stamp pytest-all
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_all.log 2>&1
echo "pytest all rc=$?"
tail -n 8 gpurun_out/pytest_all.log
stamp done
