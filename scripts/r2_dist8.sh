#!/bin/bash
# round 2, N GPUs (default 8): weak (slab), strong, cubic 512^3, BiCGStab + ILU(0) on the Laplacian
mkdir -p gpurun_out
N=${NG:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag, args...
  tag=$1; shift
  timeout 400 $TR --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 3 --warmup 2 "$@" > gpurun_out/r02_bench_p${N}_$tag.json 2> gpurun_out/r02_bench_p${N}_$tag.err; echo "$tag exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_p${N}_$tag.json"))
    print("$tag", "value %.1f %s" % (d["value"], d["unit"]), "ms/it %.4f" % d["ms_per_iteration"], "iters", d["config"]["iterations_per_solve"], "n", d["config"]["n"], "sweep frac %.3f" % d["roofline"]["frac"], "spmv ms %.4f" % d["roofline_spmv"]["ms"], "e2e %.1f" % d["e2e"]["value"])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r02_bench_p${N}_$tag.err").read()[-800:])
PY
}
for t in ${RUNS:-weak strong cubic bicgstab}; do
case $t in
 weak) run weak_cg_ilu0 ;;
 strong) run strong_cg_ilu0 --scaling strong ;;
 cubic) run cubic_cg_ilu0 --shape cubic ;;
 bicgstab) run weak_bicgstab_ilu0_lap --workload bicgstab_ilu0 --operator lap ;;
 amg) run cubic_cg_amg --workload cg_amg --shape cubic --amg-order 2 ;;
 strongcd) run strong_bicgstab_ilu0_cd --scaling strong --workload bicgstab_ilu0 ;;
 c3) run strong_bicgstab_iluk1_cd --scaling strong --workload bicgstab_iluk1 ;;
esac
done
