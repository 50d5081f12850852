for t in "8,8,8" "16,8,4" "8,8,4" "4,8,8" "16,8,8" "8,16,8"; do
  echo "TILE=$t: $(LSSPG_TRI_TILE=$t python scripts/kbench.py 256 2>&1 | tail -1 | grep -o '"ilu0_apply_ms": [0-9.]*')"
done
for c in 4 5 6; do
  echo "CTAS=$c: $(LSSPG_TRI_TILED_CTAS_PER_SM=$c python scripts/kbench.py 256 2>&1 | tail -1 | grep -o '"ilu0_apply_ms": [0-9.]*')"
done
