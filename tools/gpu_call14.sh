#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 240 $TR tools/dist_check.py 4096 > gpurun_out/r02_dist_check_2gpu_n4096.txt 2>&1; echo "dist small rc=$?"; tail -5 gpurun_out/r02_dist_check_2gpu_n4096.txt
timeout 300 $TR bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/r02_bench_2gpu.json')); print('value', b['value'], 'e2e', b['e2e']['value'], 'logz', b['e2e']['logz'], 'mcmc ms/step', b['roofline']['ms_per_step']); print(b['iteration_ms'])"
PROFILE_WARM_RUNS=1 timeout 200 $TR tools/profile_run.py > gpurun_out/r02_stage_profile_2gpu.txt 2>&1; echo "profile rc=$?"; tail -2 gpurun_out/r02_stage_profile_2gpu.txt
