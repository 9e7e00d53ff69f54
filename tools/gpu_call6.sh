#!/bin/bash
# round 2, GPU call 6 (2 GPUs): where does the partitioned step lose time?  weak scaling (1e6 rows per GPU), variants
N=${1:-2}
mkdir -p gpurun_out
run() { # name, env assignments, extra args
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-e2e $3 > gpurun_out/r2_diag_n${N}_$1.json 2> gpurun_out/r2_diag_n${N}_$1.err
  echo "== $1 rc=$?"; grep "dist trace" gpurun_out/r2_diag_n${N}_$1.err | tail -2
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_diag_n${N}_$1.json').read().strip().splitlines()[-1])
    print('$1', 'ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), {k: v for k, v in d.items() if k.startswith('parity_') and k != 'parity_note'}, d['halo'] and d['halo']['exchange'][:12])
except Exception as e:
    print('$1 unparsed', e)
PY
}
STEPS=3 run trace "GSPB200_DIST_TRACE=1" "--scaling weak"
run default "GSPB200_X=0" "--scaling weak"
run forward "GSPB200_BENCH_CLENSHAW=0" "--scaling weak"
run unfused "GSPB200_FUSE_HALO=0" "--scaling weak"
run nccl "GSPB200_EXCHANGE=nccl" "--scaling weak"
run norev "GSPB200_TILE_REV=0" "--scaling weak"
run localorder "GSPB200_BENCH_LOCAL_ORDER=1" "--scaling weak"
run config5slice "GSPB200_X=0" "--workload config5 --vertices $((6250000 * N))"
tail -c 1500 gpurun_out/r2_diag_n${N}_config5slice.err
