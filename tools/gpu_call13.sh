#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/xgpu_bench.py > gpurun_out/r02_xgpu_bench_1gpu.log 2>&1; echo "xgpu rc=$?"; tail -5 gpurun_out/r02_xgpu_bench_1gpu.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu.txt
PROFILE_WARM_RUNS=1 timeout 300 python tools/profile_run.py > gpurun_out/r02_stage_profile_d.txt 2>&1; echo "profile rc=$?"; tail -2 gpurun_out/r02_stage_profile_d.txt
