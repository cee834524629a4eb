#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/cdf_bench.py > gpurun_out/r02_cdf_bench.txt 2>&1; echo "cdf rc=$?"; tail -5 gpurun_out/r02_cdf_bench.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "cdf or multinomial or guided or next_beta" > gpurun_out/r02_pytest_cdf.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_cdf.txt
