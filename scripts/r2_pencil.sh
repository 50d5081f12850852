#!/bin/bash
# round 2: pencil sweep -- parity of the sweep tests, then bench at several pencil shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "tri or ilu or sweep or schedule" 2>&1 | tail -15 > gpurun_out/r02_pencil_pytest.log
cat gpurun_out/r02_pencil_pytest.log
for shape in ${SHAPES:-"16,16"}; do
LSSPG_TRI_PENCIL=$shape timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_pencil_${shape/,/x}.json 2> gpurun_out/r02_bench_pencil_${shape/,/x}.err; echo "bench pencil $shape exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_pencil_${shape/,/x}.json"))
    print("$shape", "it/s %.1f" % d["value"], "ms/it %.3f" % d["ms_per_iteration"], "sweep ms %.4f" % d["roofline"]["ms"], "frac %.3f" % d["roofline"]["frac"], "iters", d["config"]["iterations_per_solve"], "res", d["config"]["residual"])
except Exception as e:
    print("$shape failed", e)
PY
tail -3 gpurun_out/r02_bench_pencil_${shape/,/x}.err
done
