#!/bin/bash
# round 2, GPU call 7 (one GPU): SpMV variants, halo-capable instantiation after the per-tile split, e2e with 2 x 32
mkdir -p gpurun_out
timeout 300 python tools/spmv_probe.py > gpurun_out/r2_spmv_probe.jsonl 2> gpurun_out/r2_spmv_probe.err; cat gpurun_out/r2_spmv_probe.jsonl; tail -3 gpurun_out/r2_spmv_probe.err
timeout 600 python tools/perf_probe.py --rounds 4 --calls 10 clenshaw "clenshaw:FORCE_HALO=1" "clenshaw:TILE_VDIR=0" forward > gpurun_out/r2_probe_halo_split.jsonl 2> gpurun_out/r2_probe_halo_split.err; cat gpurun_out/r2_probe_halo_split.jsonl; tail -3 gpurun_out/r2_probe_halo_split.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_b.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_b.log
for c in default 32; do
  if [ $c = default ]; then unset GSPB200_E2E_CHUNK; else export GSPB200_E2E_CHUNK=$c; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-targets --no-cpu-baseline > gpurun_out/r2_bench_e2e_chunk_$c.json 2>> gpurun_out/r2_probe_halo_split.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_e2e_chunk_$c.json').read().strip().splitlines()[-1]);print('chunk $c', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pipeline'][:60])"
done
