#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(b['value'], b['e2e']['value'], b['roofline']['ms_per_step'], b['iteration_ms'][:5]); print({k:round(v['frac'],3) for k,v in b['roofline']['other_kernels'].items()})"
timeout 120 python tools/xgpu_bench.py > gpurun_out/r02_xgpu_bench_1gpu.log 2>&1; echo "xgpu rc=$?"; tail -5 gpurun_out/r02_xgpu_bench_1gpu.log
timeout 300 python tools/ab_normals.py 65536 24 > gpurun_out/r02_normals_ab.txt 2>&1; echo "ab rc=$?"; tail -3 gpurun_out/r02_normals_ab.txt
