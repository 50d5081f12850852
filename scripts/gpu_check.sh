#!/bin/bash
# One gpurun call: GPU parity tests, smoke, kernel micro-benchmarks, a short bench.  Output -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -q -m gpu ${PYTEST_ARGS} 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
if [ -n "${KBENCH}" ]; then timeout 900 python scripts/kbench.py ${KBENCH} > gpurun_out/kbench.log 2>&1; cat gpurun_out/kbench.log; fi
if [ -n "${GRID}" ]; then
timeout 900 python bench.py --grid ${GRID} --steps 3 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_${GRID}.json 2> gpurun_out/bench_${GRID}.err; echo "bench exit: $?"
cat gpurun_out/bench_${GRID}.json; tail -5 gpurun_out/bench_${GRID}.err
fi
