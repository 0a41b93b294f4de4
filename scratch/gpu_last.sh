#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PYTHONUNBUFFERED=1
timeout 400 python bench.py --ledger gpurun_out/ledger_last.json > gpurun_out/bench_last_n1.json 2> gpurun_out/bench_last_n1.err
echo "bench rc=$?"; cut -c1-220 gpurun_out/bench_last_n1.json; tail -2 gpurun_out/bench_last_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/bench_last_reference.json 2> gpurun_out/bench_last_reference.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_last_reference.json
timeout 200 python -m pytest tests/test_gpu_tc32.py -q -x -k "default and (golden or child or epilogues)" > gpurun_out/pytest_last.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_last.log
