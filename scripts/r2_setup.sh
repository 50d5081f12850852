#!/bin/bash
# round 2: device-side set-up (matops_gpu.cu, ilu_gpu.cu) and the reworked reference-order reductions on the GPU
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py -q -x -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02_setup_pytest.log; tail -12 gpurun_out/r02_setup_pytest.log
timeout 300 python -m pytest tests/test_gpu_exact_sum.py -q -x 2>&1 | tail -8 > gpurun_out/r02_exact_pytest1.log; tail -3 gpurun_out/r02_exact_pytest1.log
timeout 600 python -m pytest tests/test_gpu_baseline_sizes.py -q -x -s -k "EQUAL" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_exact_pytest3.log; tail -8 gpurun_out/r02_exact_pytest3.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "ilu" 2>&1 | tail -5
