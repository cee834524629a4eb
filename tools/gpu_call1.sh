#!/bin/bash
# round-2 baseline measurement on one GPU: tests, both bench arms, stage profile, ncu launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r02_bench_1gpu.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
cut -c1-600 gpurun_out/r02_bench_reference_arm.json
PROFILE_WARM_RUNS=1 timeout 300 python tools/profile_run.py > gpurun_out/r02_stage_profile.txt 2>&1; echo "profile rc=$?"
tail -2 gpurun_out/r02_stage_profile.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/short_run.py 1048576 36 > gpurun_out/r02_ncu_launch.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02_launches.csv
