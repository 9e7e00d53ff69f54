#!/bin/bash
# round 2, GPU call 2 (one GPU): variance probe, ncu launch list + full capture of the Clenshaw step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pinned or clenshaw_is or staging or spmv" > gpurun_out/r2_pytest_new.log 2>&1; tail -3 gpurun_out/r2_pytest_new.log
timeout 600 python tools/perf_probe.py --rounds 6 --calls 10 > gpurun_out/r2_probe.jsonl 2> gpurun_out/r2_probe.err; cat gpurun_out/r2_probe.jsonl; tail -3 gpurun_out/r2_probe.err
for i in 1 2 3; do timeout 300 python bench.py --steps 10 --warmup 3 --no-targets --no-cpu-baseline --no-e2e > gpurun_out/r2_bench_repeat_$i.json 2>>gpurun_out/r2_probe.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_repeat_$i.json').read().strip().splitlines()[-1]);print('repeat $i', d['ms_per_step'], d['clocks'])"; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-targets --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_launch_run.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cheby_step_tiled -s 40 -c 2 -o gpurun_out/r2_clenshaw_step -f python bench.py --steps 2 --warmup 3 --no-targets --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_full_run.log 2>&1
ls -la gpurun_out | tail -8
# gather-bound probe: the same workload with half the neighbours (k = 5): CSR bytes -4 %, gathers -50 %
timeout 300 python bench.py --steps 10 --warmup 3 --no-targets --no-cpu-baseline --no-e2e --k 5 > gpurun_out/r2_bench_probe_k5.json 2>>gpurun_out/r2_probe.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_bench_probe_k5.json').read().strip().splitlines()[-1]);print('k5', d['ms_per_step'], d['roofline']['frac'], d['graph'])"
