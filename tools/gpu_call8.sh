#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/cdf_bench.py > gpurun_out/r02_cdf_bench.txt 2>&1; echo "cdf rc=$?"; tail -5 gpurun_out/r02_cdf_bench.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_pytest_gpu.txt
PROFILE_WARM_RUNS=1 timeout 300 python tools/profile_run.py > gpurun_out/r02_stage_profile_c.txt 2>&1; echo "profile rc=$?"; tail -2 gpurun_out/r02_stage_profile_c.txt
PROFILE_WARM_RUNS=1 timeout 300 python tools/profile_run.py 32768 > gpurun_out/r02_stage_profile_n32768.txt 2>&1; echo "profile small rc=$?"; tail -2 gpurun_out/r02_stage_profile_n32768.txt
