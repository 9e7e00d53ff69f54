#!/bin/bash
# round 2, GPU call 12 (one GPU): two-packet lane mapping (parity + A/B), SpMV lanes per row, e2e chunking
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane_mappings or spmv_forms" > gpurun_out/r2_pytest_p2.log 2>&1; tail -3 gpurun_out/r2_pytest_p2.log
timeout 200 python tools/perf_probe.py --rounds 3 --calls 10 clenshaw "clenshaw:TILE_P2=1" "clenshaw:TILE_P2=1,TILE_S=3" "clenshaw:TILE_P2=1,TILE_VDIR=0" forward "forward:TILE_P2=1" > gpurun_out/r2_probe_p2_nsig64.jsonl 2> gpurun_out/r2_probe_p2.err; cat gpurun_out/r2_probe_p2_nsig64.jsonl; tail -2 gpurun_out/r2_probe_p2.err
timeout 120 python tools/perf_probe.py --rounds 3 --calls 10 --nsig 32 clenshaw "clenshaw:TILE_P2=1" > gpurun_out/r2_probe_p2_nsig32.jsonl 2>> gpurun_out/r2_probe_p2.err; cat gpurun_out/r2_probe_p2_nsig32.jsonl
timeout 120 python tools/perf_probe.py --rounds 3 --calls 5 --nsig 128 clenshaw "clenshaw:TILE_P2=1" > gpurun_out/r2_probe_p2_nsig128.jsonl 2>> gpurun_out/r2_probe_p2.err; cat gpurun_out/r2_probe_p2_nsig128.jsonl
timeout 120 python tools/spmv_probe.py > gpurun_out/r2_spmv_probe_b.jsonl 2> gpurun_out/r2_spmv_probe_b.err; cat gpurun_out/r2_spmv_probe_b.jsonl; tail -2 gpurun_out/r2_spmv_probe_b.err
for c in default 32; do
  if [ $c = default ]; then unset GSPB200_E2E_CHUNK; else export GSPB200_E2E_CHUNK=$c; fi
  timeout 200 python bench.py --steps 10 --warmup 3 --no-targets --no-cpu-baseline > gpurun_out/r2_bench_e2e_chunk_$c.json 2>> gpurun_out/r2_probe_p2.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_e2e_chunk_$c.json').read().strip().splitlines()[-1]);print('chunk $c', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pipeline'][:60])"
done
unset GSPB200_E2E_CHUNK
GSPB200_TILE_P2=1 GSPB200_E2E_CHUNK=32 timeout 200 python bench.py --steps 10 --warmup 3 --no-targets --no-cpu-baseline > gpurun_out/r2_bench_e2e_chunk_32_p2.json 2>> gpurun_out/r2_probe_p2.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_e2e_chunk_32_p2.json').read().strip().splitlines()[-1]);print('chunk 32 + P2', d['ms_per_step'], d['e2e']['ms_per_step'])"
tail -3 gpurun_out/r2_probe_p2.err
