#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_conv.py tests/test_gpu_grid.py tests/test_gpu_model.py tests/test_gpu_scene.py tests/test_gpu_tc32.py -m gpu -x -q -s > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "parity|run_scene|passed|failed|Error" gpurun_out/r2n_pytest.log | tail -n 30 | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --ledger gpurun_out/r2n_ledger.json > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
for f in ['gpurun_out/r2n_bench.json']:
    try:
        d = json.load(open(f))
        print('ms/step', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'conv avg us', d['roofline']['avg_launch_us'])
    except Exception as e:
        print('parse failed', e)
PY
timeout 300 python bench.py --no-cpu-baseline --dense-rules --steps 10 > gpurun_out/r2n_bench_dense.json 2> gpurun_out/r2n_bench_dense.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_dense.json')); print('dense-rules ms/step', d['ms_per_step'])"
for U in 4000 16000 40000; do
timeout 300 python bench.py --no-cpu-baseline --ur-min-rows $U --steps 10 > gpurun_out/r2n_bench_ur$U.json 2> gpurun_out/r2n_bench_ur$U.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_ur$U.json')); print('ur-min-rows $U ms/step', d['ms_per_step'])"
done
