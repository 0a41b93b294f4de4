#!/bin/bash
# round 2 validation pass: full GPU test suite, smoke, bench (both arms, all configs)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp pytest-all
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_all.log 2>&1
echo "pytest all rc=$?"; tail -n 12 gpurun_out/r2_pytest_all.log | cut -c1-300
stamp smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/r2_smoke.log
stamp bench-n1
timeout 500 python bench.py --ledger gpurun_out/r2_ledger.json > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2_bench_n1.json'))
    print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e value', d['e2e']['value'], 'h2d', d['e2e']['h2d_bytes_per_step'], 'd2h', d['e2e']['d2h_bytes_per_step'])
    print('roofline frac', d['roofline']['frac'], 'achieved', d['roofline']['achieved'], 'share', d['roofline']['share_of_step'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline'])
except Exception as e:
    print('parse failed', e)
PY
for C in 2 4; do
  stamp bench-config-$C
  timeout 300 python bench.py --config $C --steps 20 --warmup 3 > gpurun_out/r2_bench_config$C.json 2> gpurun_out/r2_bench_config$C.err
  echo "config $C rc=$?"; cut -c1-600 gpurun_out/r2_bench_config$C.json
done
stamp done
