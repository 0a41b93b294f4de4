#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2v_pytest.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --ledger gpurun_out/r2v_ledger.json > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2v_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'conv avg us', d['roofline']['avg_launch_us'])
PY
python scratch/phases.py tc32 10 > gpurun_out/r2v_phases.txt 2>&1; cat gpurun_out/r2v_phases.txt
