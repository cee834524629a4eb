#!/bin/bash
# final single-GPU record of the round: tests, both bench arms, exchange latency, launch list, ncu of the Metropolis kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/r02_bench_1gpu.json')); print(b['value'], b['e2e']['value'], b['roofline']['ms_per_step'], b['roofline']['frac'], b['iteration_ms'][:5]); print({k:round(v['frac'],3) for k,v in b['roofline']['other_kernels'].items()})"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
timeout 100 python tools/xgpu_bench.py > gpurun_out/r02_xgpu_bench_1gpu.log 2>&1; echo "xgpu rc=$?"; grep -E "^grid" gpurun_out/r02_xgpu_bench_1gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/short_run.py 1048576 36 > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mcmc_run_fast --launch-skip 20 --launch-count 1 -o gpurun_out/r02_mcmc_run_fast_after -f python tools/short_run.py 1048576 27 > gpurun_out/r02_ncu_mcmc.log 2>&1; echo "ncu mcmc rc=$?"
