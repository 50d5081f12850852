#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench.  Output lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --grid ${GRID:-128} --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit: $?"
cat gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
