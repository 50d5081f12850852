#!/bin/bash
mkdir -p gpurun_out
N=${NG:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for p2p in 1 0; do
LSSPG_GRAPHS=0 LSSPG_P2P_ALLREDUCE=$p2p timeout 200 $TR --master-port $((29700 + p2p)) tests/dist_check_gpu.py > gpurun_out/r02_dist_check_p${N}_p2p$p2p.log 2>&1; echo "P=$N p2p=$p2p exit $?"
grep -E "DIST_CHECK|ok=False|Error|error" gpurun_out/r02_dist_check_p${N}_p2p$p2p.log | head -8
done
