#!/bin/bash
# round 2, GPU call 1 (one GPU): tests, default bench line, e2e staging variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_n1.err
for v in "kernel 32" "dma 16" "kernel 16" "dma 0"; do
  set -- $v
  GSPB200_STAGE=$1 GSPB200_E2E_CHUNK=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-targets --no-cpu-baseline \
     > gpurun_out/r2_bench_n1_stage_$1_chunk$2.json 2>> gpurun_out/r2_bench_n1.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_n1*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'e2e ms', d['e2e'] and round(d['e2e']['ms_per_step'], 3),
              'lanczos', d.get('estimate_lmax'), 'targets', {k: (round(v.get('ms_per_step', 0), 2), round(v.get('roofline_frac', 0), 3), v.get('parity_rel_err_one_column_vs_oracle'), v.get('error')) for k, v in (d.get('targets') or {}).items()})
    except Exception as e:
        print(f, 'unparsed', e)
PY
