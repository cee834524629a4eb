#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/cdf_bench.py > gpurun_out/r02_cdf_bench.txt 2>&1; echo "cdf rc=$?"; tail -5 gpurun_out/r02_cdf_bench.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu.txt
PROFILE_WARM_RUNS=1 timeout 300 python tools/profile_run.py > gpurun_out/r02_stage_profile_b.txt 2>&1; echo "profile rc=$?"; tail -2 gpurun_out/r02_stage_profile_b.txt
