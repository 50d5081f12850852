#!/bin/bash
# round 2, 2 GPUs: parity under torchrun with the peer-to-peer all-reduce, then weak / strong bench lines
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/dist_check_gpu.py > gpurun_out/r02_dist_check_p2.log 2>&1; echo "dist check exit $?"; tail -12 gpurun_out/r02_dist_check_p2.log
LSSPG_P2P_ALLREDUCE=0 timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_p2_weak_nccl.json 2> gpurun_out/r02_bench_p2_weak_nccl.err; echo "weak nccl exit $?"
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_p2_weak.json 2> gpurun_out/r02_bench_p2_weak.err; echo "weak p2p exit $?"
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong > gpurun_out/r02_bench_p2_strong.json 2> gpurun_out/r02_bench_p2_strong.err; echo "strong exit $?"
for f in weak_nccl weak strong; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_p2_$f.json"))
    print("$f", "value %.1f %s" % (d["value"], d["unit"]), "ms/it %.4f" % d["ms_per_iteration"], "iters", d["config"]["iterations_per_solve"], "sweep frac %.3f" % d["roofline"]["frac"], "spmv ms %.4f" % d["roofline_spmv"]["ms"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r02_bench_p2_$f.err").read()[-1500:])
PY
done
