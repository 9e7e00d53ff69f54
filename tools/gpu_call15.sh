#!/bin/bash
# last sanity check of the final tree on one GPU (what is left of the budget)
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "readme or lane_mappings or clenshaw_is or pinned" 2>&1 | tail -2
