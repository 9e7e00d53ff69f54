#!/bin/bash
# round 2, GPU call 5 (one GPU): direct-vector mode of the tiled step, SpMV with the x window, e2e timeline
mkdir -p gpurun_out
timeout 600 python tools/perf_probe.py --rounds 4 --calls 10 clenshaw "clenshaw:FORCE_HALO=1" "clenshaw:TILE_VDIR=1" "clenshaw:TILE_VDIR=1,TILE_S=4" "clenshaw:TILE_VDIR=1,TILE_S=4,TILE_R=32" "clenshaw:TILE_VDIR=1,TILE_S=3,TILE_R=128" forward "forward:TILE_VDIR=1,TILE_S=4" > gpurun_out/r2_probe_vdir.jsonl 2> gpurun_out/r2_probe_vdir.err; cat gpurun_out/r2_probe_vdir.jsonl; tail -3 gpurun_out/r2_probe_vdir.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lmax or lanczos or spmv" > gpurun_out/r2_pytest_lanczos.log 2>&1; tail -3 gpurun_out/r2_pytest_lanczos.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_lanczos.csv python -c "
import pygsp_b200 as gsp
G = gsp.graphs.Sensor(1000000, k=10, seed=0, order='morton')
print(G.estimate_lmax(), G._lanczos_steps)
" > gpurun_out/r2_ncu_lanczos_run.log 2>&1
grep spmv_window gpurun_out/r2_launches_lanczos.csv | tail -2
GSPB200_E2E_TRACE=1 timeout 300 python - <<'PY' > gpurun_out/r2_e2e_trace.txt 2>&1
import numpy as np, torch, time, json
import pygsp_b200 as gsp
from pygsp_b200.filters import pipeline
gsp.utils.bind_to_gpu_numa(0)
G = gsp.graphs.Sensor(1000000, k=10, seed=0, order='morton'); G.estimate_lmax()
g = gsp.filters.Heat(G, scale=50)
xh = torch.randn(G.N, 64).pin_memory()
for chunk in (None, "32", "16"):
    import os
    if chunk: os.environ["GSPB200_E2E_CHUNK"] = chunk
    for i in range(4):
        t0 = time.perf_counter(); y = g.filter(xh, order=30); t1 = time.perf_counter()
    print(chunk, "wall ms", round(1e3*(t1-t0), 3)); print(json.dumps(pipeline.last_trace))
PY
cat gpurun_out/r2_e2e_trace.txt
