#!/bin/bash
# round 2, GPU call 3 (N GPUs, N = $1): NCCL / peer-memory tests of the partitioned path, strong + weak bench,
# a per-GPU-sized slice of config 5, racecheck of a 2-rank run
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
timeout 1200 python -m pytest tests/test_distributed_gpu.py -m gpu -x -q -rs > gpurun_out/r2_pytest_distributed_n$N.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_distributed_n$N.log; tail -6 gpurun_out/r2_pytest_distributed_n$N.log
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $2 > gpurun_out/r2_bench_n${N}_$1.json 2> gpurun_out/r2_bench_n${N}_$1.err
  echo "$1 rc=$?"; tail -c 600 gpurun_out/r2_bench_n${N}_$1.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_bench_n${N}_$1.json').read().strip().splitlines()[-1])
    print('$1', 'ms', round(d['ms_per_step'], 3), 'value', '%.3e' % d['value'], 'frac', round(d['roofline']['frac'], 3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'], 2),
          {k: v for k, v in d.items() if k.startswith('parity') or k.startswith('one_gpu') or k.startswith('speedup')}, d['halo'], d['graph'])
except Exception as e:
    print('$1 unparsed', e)
PY
}
run strong ""
run weak "--scaling weak"
run config5slice "--workload config5 --n $((6250000 * N))"
if [ "$N" = "2" ]; then
  for r in 0 1; do
    timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tests/dist_worker.py nccl 2 $r 29533 8000 64 1 6 p2p > gpurun_out/r2_racecheck_2rank_fused_rank$r.log 2>&1 &
  done
  wait
  tail -4 gpurun_out/r2_racecheck_2rank_fused_rank0.log gpurun_out/r2_racecheck_2rank_fused_rank1.log
fi
