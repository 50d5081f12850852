#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_exact_sum.py -q -x 2>&1 | tail -3
export LSSPG_SETUP_PROF=1
timeout 240 python -m pytest tests/test_gpu_setup.py -q -x -s -k "ilut or cd3d/iluk1" 2>&1 | grep -v "^$" | grep -v "setup\] \|0.000 s" | tail -40 > gpurun_out/r02_setup_pytest3.log; tail -30 gpurun_out/r02_setup_pytest3.log
