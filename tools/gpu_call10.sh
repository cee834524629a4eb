#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 240 $TR tools/dist_check.py 4096 > gpurun_out/r02_dist_check_2gpu_n4096.txt 2>&1; echo "dist small rc=$?"; tail -6 gpurun_out/r02_dist_check_2gpu_n4096.txt
timeout 240 $TR tools/dist_timing.py > gpurun_out/r02_dist_timing_2gpu.txt 2>&1; echo "timing rc=$?"; grep -E "^rep|posterior" gpurun_out/r02_dist_timing_2gpu.txt | cut -c1-600
timeout 300 $TR bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/r02_bench_2gpu.json')); print(b['value'], b['e2e']['value'], b['e2e']['logz'], b['iteration_ms'][:6])"
tail -3 gpurun_out/r02_bench_2gpu.err
