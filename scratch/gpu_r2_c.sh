#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 300 python scratch/sp_bench.py > gpurun_out/r2o_sp_bench.txt 2>&1; cat gpurun_out/r2o_sp_bench.txt | tail -20
timeout 1500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_grid.py tests/test_gpu_model.py tests/test_gpu_scene.py tests/test_gpu_tc32.py -m gpu -x -q -s > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "parity|run_scene|passed|failed|Error" gpurun_out/r2o_pytest.log | tail -n 30 | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --ledger gpurun_out/r2o_ledger.json > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/r2o_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'conv avg us', d['roofline']['avg_launch_us'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_launches.csv python scratch/one_pass.py tc32 4 > gpurun_out/r2o_ncu_list.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/r2o_ncu_list.log | cut -c1-300
