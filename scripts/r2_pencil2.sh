#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tri or ilu" 2>&1 | tail -3
LSSPG_TRI_PENCIL_HOLES=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "trisolve_and_ilu" 2>&1 | tail -2
python scripts/pencil_prof.py 256 0 > gpurun_out/r02_pencil_prof2.log 2>&1
LSSPG_TRI_PROF=1 python scripts/pencil_prof.py 256 0 2>&1 | grep -E "sweep|pencils|duration|ticket    0|ticket  1[0-9][0-9]:" >> gpurun_out/r02_pencil_prof2.log
cat gpurun_out/r02_pencil_prof2.log
for shape in ${SHAPES}; do
LSSPG_TRI_PENCIL=$shape python scripts/pencil_prof.py 256 0 2>&1 | sed "s/^/[$shape] /"
done
