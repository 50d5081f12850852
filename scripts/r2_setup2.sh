#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py -q -x -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02_setup_pytest.log; tail -8 gpurun_out/r02_setup_pytest.log
timeout 200 python scripts/exact_prof.py 256 3000 0,2 2>&1 | tail -4 | tee gpurun_out/r02_exact_prof.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r02_exact_launches.csv python scripts/exact_prof.py 256 40 2 > gpurun_out/ncu_exact.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/r02_exact_launches.csv 2>&1 | tail -25
