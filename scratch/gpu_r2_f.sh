#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_scene.py tests/test_gpu_tc32.py -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/r2r_pytest.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --ledger gpurun_out/r2r_ledger.json > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/r2r_bench.json'))
print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'conv avg us', d['roofline']['avg_launch_us'])
PY
timeout 300 python bench.py --no-cpu-baseline --no-overlap --steps 20 > gpurun_out/r2r_bench_noov.json 2> gpurun_out/r2r_bench_noov.err
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_noov.json')); print('no-overlap ms/step', d['ms_per_step'])"
timeout 300 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/r2r_bench2.json 2> gpurun_out/r2r_bench2.err
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench2.json')); print('again ms/step', d['ms_per_step'])"
