#!/bin/bash
G=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517"
timeout 200 $TR bench.py --gpus $G > gpurun_out/r02_bench_${G}gpu.json 2> gpurun_out/r02_bench_${G}gpu.err; echo "bench rc=$?"; python -c "
import json; b=json.load(open('gpurun_out/r02_bench_${G}gpu.json')); print('value', b['value'], 'e2e', b['e2e']['value'], 'logz', b['e2e']['logz'], 'mcmc ms/step', b['roofline']['ms_per_step']); print(b['iteration_ms'])"
