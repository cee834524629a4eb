#!/bin/bash
# round-2 two-GPU check: 2-GPU equality test, dist_check (small + 2^20), bench, stage profile
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
# (the pytest 2-GPU test runs the same dist_check 4096)
timeout 300 $TR tools/dist_check.py 4096 > gpurun_out/r02_dist_check_2gpu_n4096.txt 2>&1; echo "dist small rc=$?"; tail -5 gpurun_out/r02_dist_check_2gpu_n4096.txt
timeout 400 $TR tools/dist_check.py 1048576 --big > gpurun_out/r02_dist_check_2gpu_n2pow20.txt 2>&1; echo "dist big rc=$?"; tail -12 gpurun_out/r02_dist_check_2gpu_n2pow20.txt
timeout 400 $TR bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/r02_bench_2gpu.json
PROFILE_WARM_RUNS=1 timeout 300 $TR tools/profile_run.py > gpurun_out/r02_stage_profile_2gpu.txt 2>&1; echo "profile rc=$?"; tail -2 gpurun_out/r02_stage_profile_2gpu.txt
