#!/bin/bash
# round 2, last multi-GPU check (2 GPUs) of the final code: peer-memory exchange tests + strong-scaling line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q -k "two_ranks and p2p" > gpurun_out/r2_pytest_distributed_n2_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_distributed_n2_final.log; tail -3 gpurun_out/r2_pytest_distributed_n2_final.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_strong_final.json 2> gpurun_out/r2_bench_n2_strong_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_n2_strong_final.json').read().strip().splitlines()[-1])
print('strong2 final', 'ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['ms_per_step'], 2), {k: v for k, v in d.items() if k.startswith('parity') or k.startswith('one_gpu') or k.startswith('speedup')})
PY
tail -c 300 gpurun_out/r2_bench_n2_strong_final.err | grep -v "^\*\|OMP"
