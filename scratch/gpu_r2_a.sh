#!/bin/bash
# new f-row tests + current launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_mesh.py tests/test_gpu_scene.py -m gpu -x -q -s > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/r2m_pytest.log | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_ncu_list.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/r2m_ncu_list.log | cut -c1-300
