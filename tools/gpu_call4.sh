#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR tools/dist_timing.py > gpurun_out/r02_dist_timing_2gpu.txt 2>&1; echo "timing rc=$?"; grep -E "^rep|posterior" gpurun_out/r02_dist_timing_2gpu.txt | cut -c1-700
BENCH_TRACE=1 timeout 400 $TR bench.py --gpus 2 > gpurun_out/r02_bench_2gpu_b.json 2> gpurun_out/r02_bench_2gpu_b.err; echo "bench rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/r02_bench_2gpu_b.json')); print(b['value'], b['e2e']['value'], b['iteration_ms'])"
grep -E "iteration host|e2e" gpurun_out/r02_bench_2gpu_b.err | cut -c1-600
