#!/bin/bash
# parameter sweep of the triangular-solve kernels.  usage: tri_sweep.sh N
N=${1:-128}
run() { echo "$1: $(env $1 timeout 300 python scripts/kbench.py $N 2>&1 | python -c 'import sys,json; r=json.loads(sys.stdin.readline()); print("apply_ms", r["ilu0_apply_ms"], "us/level", r["us_per_level"], "cg_it_ms", r["cg_ilu0_ms_per_it"], "bicg_it_ms", r["bicgstab_ilu0_ms_per_it"])')"; }
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tri or ilu or pc" 2>&1 | tail -3
run "LSSPG_TRI_TILED=0"
run "LSSPG_TRI_BOX_FLAGS=0"
run "LSSPG_TRI_BOX_FLAGS=1"
for t in "4,4,4" "8,8,4" "8,4,4" "16,4,4" "16,8,8" "16,16,4"; do run "LSSPG_TRI_TILE=$t"; done
