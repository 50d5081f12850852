#!/bin/bash
# round 2, final evidence: full GPU suite, smoke, bench, ncu launch list of the bench
mkdir -p gpurun_out
timeout 800 python -m pytest tests -q -m gpu -x --durations=12 2>&1 | tail -30 > gpurun_out/r02_final_pytest_gpu.log; tail -22 gpurun_out/r02_final_pytest_gpu.log
timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2 | tee gpurun_out/r02_final_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_final_bench_256.json 2> gpurun_out/r02_final_bench_256.err; echo "bench exit $?"; cut -c1-2600 gpurun_out/r02_final_bench_256.json; tail -3 gpurun_out/r02_final_bench_256.err
CMD="python bench.py --grid 256 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/r02_final_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "launch list exit $?"
python scripts/summarize_launches.py gpurun_out/r02_final_launches.csv "$CMD (launches 2000..2399)" | tee gpurun_out/r02_final_launches_summary.txt | head -12
