#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "cdf or multinomial or guided" > gpurun_out/r02_pytest_cdf.txt 2>&1; echo "pytest cdf rc=$?"; tail -15 gpurun_out/r02_pytest_cdf.txt
timeout 200 python tools/cdf_bench.py > gpurun_out/r02_cdf_bench.txt 2>&1; echo "cdf rc=$?"; tail -5 gpurun_out/r02_cdf_bench.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mcmc_run_fast --launch-skip 20 --launch-count 1 -o gpurun_out/r02_mcmc_run_fast -f python tools/short_run.py 1048576 27 > gpurun_out/r02_ncu_mcmc.log 2>&1; echo "ncu mcmc rc=$?"; tail -3 gpurun_out/r02_ncu_mcmc.log
timeout 300 python tools/ab_normals.py 65536 24 > gpurun_out/r02_normals_ab.txt 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/r02_normals_ab.txt
