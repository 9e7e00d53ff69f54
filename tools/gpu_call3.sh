#!/bin/bash
# round 2, GPU call 3 (N GPUs, N = $1): NCCL / peer-memory tests of the partitioned path, strong + weak bench
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q -rs > gpurun_out/r2_pytest_distributed_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_distributed_n$N.log; tail -6 gpurun_out/r2_pytest_distributed_n$N.log
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $2 > gpurun_out/r2_bench_n${N}_$1.json 2> gpurun_out/r2_bench_n${N}_$1.err
  echo "$1 rc=$?"; tail -c 400 gpurun_out/r2_bench_n${N}_$1.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_n${N}_$1.json').read().strip().splitlines()[-1])
    print('$1', 'ms', round(d['ms_per_step'], 3), 'value', '%.3e' % d['value'], 'frac', round(d['roofline']['frac'], 3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'], 2),
          {k: d.get(k) for k in ('parity_rel_err', 'parity_bit_identical_on_every_rank', 'one_gpu_same_graph_ms_per_step', 'speedup_vs_one_gpu_same_run', 'parity_rel_err_one_column_vs_oracle')}, d['halo'], d['graph'])
except Exception as e:
    print('$1 unparsed', e)
PY
}
run strong ""
run weak "--scaling weak"
