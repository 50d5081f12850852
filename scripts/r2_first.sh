#!/bin/bash
# round 2, first call: run the experimental sweep variants, compare chunked vs default bench, tri phase profile
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
LSSPG_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_yexperimental.py -q 2>&1 | tail -40 > gpurun_out/r02_experimental.log
tail -15 gpurun_out/r02_experimental.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench default exit $?"
cat gpurun_out/r02_bench_default.json
for c in 2 3 4 6; do
LSSPG_TRI_CHUNKS=$c timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_chunks$c.json 2> gpurun_out/r02_bench_chunks$c.err; echo "bench chunks $c exit $?"
cat gpurun_out/r02_bench_chunks$c.json
done
LSSPG_TRI_SKEW_FORCE=1,1,1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_skewforce.json 2> gpurun_out/r02_bench_skewforce.err; echo "bench skewforce exit $?"
cat gpurun_out/r02_bench_skewforce.json
LSSPG_TRI_PROF=1 timeout 300 python scripts/tri_prof.py 256 > gpurun_out/r02_tri_prof.log 2>&1
tail -30 gpurun_out/r02_tri_prof.log
