#!/bin/bash
# round 2: the parallel reference-order reductions (exact_sum.cu) on the GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_exact_sum.py -q -x 2>&1 | tail -15 > gpurun_out/r02_exact_pytest1.log; tail -4 gpurun_out/r02_exact_pytest1.log
timeout 600 python -m pytest tests/test_gpu_all_drivers.py -q -x -k "parallel_reference" 2>&1 | tail -15 > gpurun_out/r02_exact_pytest2.log; tail -4 gpurun_out/r02_exact_pytest2.log
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -q -x -s -k "EQUAL" 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_exact_pytest3.log; tail -8 gpurun_out/r02_exact_pytest3.log
