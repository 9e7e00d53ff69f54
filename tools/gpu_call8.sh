#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dist_probe.py > gpurun_out/r2_dist_probe.jsonl 2> gpurun_out/r2_dist_probe.err
cat gpurun_out/r2_dist_probe.jsonl; tail -5 gpurun_out/r2_dist_probe.err
