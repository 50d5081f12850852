#!/bin/bash
# parameter sweep of the triangular-solve kernel (N=128 unless given)
N=${1:-128}
for cfg in "4 1 64 64" "4 0 64 64" "2 1 64 64" "1 1 64 64" "1 0 64 64" "8 1 64 64" "4 1 4 32" "4 1 1000000 0" "2 0 1000000 0" "1 0 1000000 0" "4 1 16 200"; do
  set -- $cfg
  echo "ctas=$1 hint=$2 spin=$3 sleep=$4: $(LSSPG_TRI_CTAS_PER_SM=$1 LSSPG_TRI_HINT=$2 LSSPG_TRI_SPIN=$3 LSSPG_TRI_SLEEP=$4 timeout 300 python scripts/kbench.py $N 2>&1 | python -c 'import sys,json; r=json.loads(sys.stdin.readline()); print(r["ilu0_apply_ms"], r["us_per_level"], r["cg_ilu0_ms_per_it"])')"
done
