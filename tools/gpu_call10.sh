#!/bin/bash
# round 2, GPU call 10 (4 GPUs): 4-rank tests (interior ranks have two neighbours), strong scaling of the 10M target,
# a 4-GPU slice of config 5, config 4 (SBM) at 4 GPUs
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q -rs -k "four_ranks" > gpurun_out/r2_pytest_distributed_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_distributed_n$N.log; tail -5 gpurun_out/r2_pytest_distributed_n$N.log
run() { # name, env assignments, extra args
  env $2 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $3 > gpurun_out/r2_bench_n${N}_$1.json 2> gpurun_out/r2_bench_n${N}_$1.err
  echo "== $1 rc=$?"; tail -c 300 gpurun_out/r2_bench_n${N}_$1.err | grep -v "^\*\|OMP_NUM" 
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_n${N}_$1.json').read().strip().splitlines()[-1])
    print('$1', 'ms', round(d['ms_per_step'], 3), 'value', '%.3e' % d['value'], 'frac', round(d['roofline']['frac'], 3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'], 2),
          {k: v for k, v in d.items() if (k.startswith('parity') and k != 'parity_note') or k.startswith('one_gpu') or k.startswith('speedup')}, d['halo'] and (d['halo']['rows_received_per_rank'], d['halo']['exchange'][:5]), d['graph'])
except Exception as e:
    print('$1 unparsed', e)
PY
}
run strong "GSPB200_X=0" ""
run config5slice "GSPB200_X=0" "--workload config5 --vertices $((6250000 * N))"
run config4 "GSPB200_X=0" "--workload config4 --no-e2e"
