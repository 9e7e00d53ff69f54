#!/bin/bash
# round 2, final GPU call (one GPU): whole GPU test-suite, driver-style bench line, reference arm, smoke, ncu evidence
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu_final.log; tail -4 gpurun_out/r2_pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_final_n1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_final_n1.json').read().strip().splitlines()[-1])
print('final', 'ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3), 'e2e ms', round(d['e2e']['ms_per_step'], 3), 'lanczos', d.get('estimate_lmax'), d['clocks'])
print({k: (round(v.get('ms_per_step', 0), 2), round(v.get('roofline_frac', 0), 3), v.get('parity_rel_err_one_column_vs_oracle'), v.get('error')) for k, v in (d.get('targets') or {}).items()})
PY
timeout 300 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-targets --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cheby_step_tiled -s 40 -c 2 -o gpurun_out/r2_final_clenshaw_step -f python bench.py --steps 2 --warmup 3 --no-targets --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_full_run.log 2>&1
ls -la gpurun_out/r2_final_clenshaw_step.ncu-rep gpurun_out/r2_launches_final.csv
