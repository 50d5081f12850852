#!/bin/bash
for d in ${DBGS:-0}; do
LSSPG_TRI_PENCIL_DBG=$d LSSPG_TRI_PROF=1 python scripts/pencil_prof.py 256 0 2>&1 | grep -E "sweep|ticket    0|per step" | sed "s/^/[dbg $d] /" | cut -c1-170
done
python scripts/pencil_prof.py 256 0
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tri or ilu" 2>&1 | tail -2
