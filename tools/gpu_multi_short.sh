#!/bin/bash
# usage: bash tools/gpu_multi_short.sh G   -- bench + exchange latency + stage profile on G GPUs (short: G x box time is charged)
G=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517"
timeout 200 $TR bench.py --gpus $G > gpurun_out/r02_bench_${G}gpu.json 2> gpurun_out/r02_bench_${G}gpu.err; echo "bench rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/r02_bench_${G}gpu.json')); print('value', b['value'], 'e2e', b['e2e']['value'], 'logz', b['e2e']['logz'], 'mcmc ms/step', b['roofline']['ms_per_step']); print(b['iteration_ms'])"
timeout 100 $TR tools/xgpu_bench.py 4000 > gpurun_out/r02_xgpu_bench_${G}gpu.log 2>&1; echo "xgpu rc=$?"; grep -E "^grid" gpurun_out/r02_xgpu_bench_${G}gpu.log
PROFILE_WARM_RUNS=1 timeout 150 $TR tools/profile_run.py > gpurun_out/r02_stage_profile_${G}gpu.txt 2>&1; echo "profile rc=$?"; tail -2 gpurun_out/r02_stage_profile_${G}gpu.txt
