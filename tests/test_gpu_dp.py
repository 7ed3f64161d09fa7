"""Multi-GPU data-parallel correctness (self-skips below 2 GPUs): scripts/dp_check.py under torchrun — DP on rank shards equals
single-process full-batch training, replicas stay bit-identical, uneven and EMPTY shards included; run for both exchanges
(fused NVLink push kernel, NCCL all-reduce)."""
import os
import subprocess
import sys
import pytest
import torch

from tests.util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_dp_equals_single_process(p2p):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, ODPD_DP_P2P=p2p)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29610 + int(p2p)), os.path.join(ROOT, "scripts", "dp_check.py")], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "dp_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
