#!/bin/bash
# round 2: full GPU suite, smoke, bench, then the ncu evidence (launch list + capture of the sweep kernel)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r02_pytest_gpu.log; tail -6 gpurun_out/r02_pytest_gpu.log
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_256.json 2> gpurun_out/r02_bench_256.err; echo "bench exit $?"; cut -c1-1500 gpurun_out/r02_bench_256.json
CMD="python bench.py --grid 256 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tri_pencil -s 20 -c 2 -o gpurun_out/r02_prof_tri -f $CMD > gpurun_out/ncu_tri.log 2>&1; echo "tri capture exit $?"
