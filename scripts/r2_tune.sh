#!/bin/bash
for cfg in "8 3 32" "12 3 32" "16 3 32" "8 4 32" "8 3 16" "8 3 64" "12 3 64" "16 4 64"; do
set -- $cfg
LSSPG_TRI_PENCIL_VR=$1 LSSPG_TRI_PENCIL_NB=$2 LSSPG_TRI_PENCIL_L2=$3 timeout 60 python scripts/pencil_prof.py 256 0 2>&1 | cut -c1-60 | tr '\n' ' '; echo " [VR=$1 NB=$2 L2=$3]"
done
