#!/bin/bash
mkdir -p gpurun_out
export LSSPG_SETUP_PROF=1
timeout 600 python -m pytest tests/test_gpu_setup.py -q -x -s -k "ilut or baseline" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02_setup_pytest2.log; tail -50 gpurun_out/r02_setup_pytest2.log
