#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_exact_sum.py -q -x 2>&1 | tail -3
timeout 200 python scripts/exact_prof.py 256 3000 0,2 2>&1 | tail -4 | tee gpurun_out/r02_exact_prof2.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r02_exact_launches2.csv python scripts/exact_prof.py 256 40 2 > gpurun_out/ncu_exact2.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/r02_exact_launches2.csv "python scripts/exact_prof.py 256 40 2 (launches 600..999)" 2>&1 | tee -a gpurun_out/r02_exact_prof2.log | tail -14
