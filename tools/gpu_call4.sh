#!/bin/bash
# round 2, GPU call 4 (one GPU): PCIe staging rates, Lanczos SpMV after the unroll
mkdir -p gpurun_out
timeout 300 python tools/copy_probe.py > gpurun_out/r2_copy_probe.jsonl 2> gpurun_out/r2_copy_probe.err; cat gpurun_out/r2_copy_probe.jsonl; tail -3 gpurun_out/r2_copy_probe.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lmax or lanczos or spmv" > gpurun_out/r2_pytest_lanczos.log 2>&1; tail -3 gpurun_out/r2_pytest_lanczos.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_lanczos.csv python -c "
import pygsp_b200 as gsp
G = gsp.graphs.Sensor(1000000, k=10, seed=0, order='morton')
print(G.estimate_lmax(), G._lanczos_steps)
" > gpurun_out/r2_ncu_lanczos_run.log 2>&1
grep -c spmv_subwarp gpurun_out/r2_launches_lanczos.csv
grep spmv_subwarp gpurun_out/r2_launches_lanczos.csv | tail -3
