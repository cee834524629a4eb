#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 240 $TR tools/host_profile.py > gpurun_out/r02_host_profile_2gpu.txt 2>&1; echo "rc=$?"; grep -E "^run" gpurun_out/r02_host_profile_2gpu.txt
