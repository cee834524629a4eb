#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(b['value'], b['e2e']['value'], b['roofline']['ms_per_step'], b['roofline']['frac'], b['iteration_ms'][:5], b['gpu_launches'])"
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.txt
