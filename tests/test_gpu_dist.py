"""Multi-GPU parity inside the GPU suite: tests/dist_check_gpu.py under torchrun, one rank per GPU -- sharded SpMV (halo
exchange) bit-exact against the CPU checker, every driver with block-Jacobi ILU needing exactly the iterations of a 1-GPU
run with the same blocked preconditioner, CG / BiCGStab counts equal to the reference's blocked ILU
(tests/golden/golden.json "blockjacobi").  Dot products travel through the one-shot peer-to-peer all-reduce kernel
(comm.cu) and, in a second run, through ncclAllReduce.  Skipped on a box with fewer GPUs than ranks."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("P,p2p", [(2, "1"), (2, "0"), (4, "1")])
def test_sharded_solves_match_under_torchrun(P, p2p):
    if _ngpus() < P:
        pytest.skip("needs %d GPUs" % P)
    env = dict(os.environ, LSSPG_P2P_ALLREDUCE=p2p)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(P), "--master-addr", "127.0.0.1",
           "--master-port", str(29530 + P + int(p2p)), os.path.join(ROOT, "tests", "dist_check_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0 and "DIST_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
