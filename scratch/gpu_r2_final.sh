#!/bin/bash
# round-2 final measurement pass (one GPU): full suite, smoke, bench lines, ncu launch list, conv DRAM traffic, full captures
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
T=${1:-r2z}
t0=$(date +%s); stamp() { echo "[$(( $(date +%s) - t0 ))s] $*"; }
stamp pytest
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/${T}_pytest.log | cut -c1-200
stamp smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${T}_smoke.log
stamp bench
timeout 600 python bench.py --ledger gpurun_out/${T}_ledger.json > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python - "$T" <<'PY'
import json, sys
d = json.load(open('gpurun_out/%s_bench_n1.json' % sys.argv[1]))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
stamp bench-exact
timeout 300 python bench.py --no-cpu-baseline --conv-mode exact > gpurun_out/${T}_bench_exact.json 2> gpurun_out/${T}_bench_exact.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_exact.json')); print('exact ms/step', d['ms_per_step'])"
for C in 2 4; do
  stamp config-$C
  timeout 300 python bench.py --config $C --steps 20 --warmup 3 > gpurun_out/${T}_bench_config$C.json 2> gpurun_out/${T}_bench_config$C.err
  echo "config $C rc=$?"; cut -c1-300 gpurun_out/${T}_bench_config$C.json
done
stamp phases
python scratch/phases.py tc32 10 > gpurun_out/${T}_phases_tc32.txt 2>&1; tail -n 1 gpurun_out/${T}_phases_tc32.txt
python scratch/phases.py exact 10 > gpurun_out/${T}_phases_exact.txt 2>&1; tail -n 1 gpurun_out/${T}_phases_exact.txt
stamp ncu-list
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches.csv python scratch/one_pass.py tc32 4 > gpurun_out/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
stamp ncu-dram
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:^conv_ --csv --log-file gpurun_out/${T}_conv_dram.csv python scratch/one_pass.py tc32 1 > gpurun_out/${T}_ncu_dram.log 2>&1; echo "ncu dram rc=$?"
stamp ncu-full
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ur_kernel --launch-skip 20 --launch-count 3 -o gpurun_out/${T}_ur_full -f python scratch/one_pass.py tc32 1 > gpurun_out/${T}_ncu_ur.log 2>&1; echo "ncu ur rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_urc_kernel --launch-skip 2 --launch-count 1 -o gpurun_out/${T}_urc_full -f python scratch/one_pass.py tc32 1 > gpurun_out/${T}_ncu_urc.log 2>&1; echo "ncu urc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_sp_kernel|tile_plan_kernel' --launch-count 6 -o gpurun_out/${T}_sp_plan_full -f python scratch/one_pass.py tc32 1 > gpurun_out/${T}_ncu_sp.log 2>&1; echo "ncu sp/plan rc=$?"
stamp ops
timeout 600 python bench_ops.py --which mesh > gpurun_out/${T}_ops_mesh.json 2> gpurun_out/${T}_ops_mesh.err; echo "ops mesh rc=$?"; cut -c1-500 gpurun_out/${T}_ops_mesh.json; timeout 600 python bench_ops.py --which scene > gpurun_out/${T}_ops_scene.json 2> gpurun_out/${T}_ops_scene.err; echo "ops scene rc=$?"; cut -c1-500 gpurun_out/${T}_ops_scene.json
stamp done
