"""torchrun --nproc-per-node N scripts/dp_check.py : data-parallel NativeTrainStep == single-process full-batch training."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from opendpd_b200 import models, dp
from opendpd_b200.train import NativeTrainStep
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
net = models.CoreModel(2, 13, 1, "dgru").to(dev)
ref = copy.deepcopy(net)
g = torch.Generator().manual_seed(5)
N, T, B = 4000, 256, 24
stream = (0.25 * torch.randn(N, 2, generator=g)).to(dev)
target = 0.9 * stream
perm = dp.epoch_permutation(N - T + 1, seed=0)
tr = NativeTrainStep(net, process_group=dist.group.WORLD, world_size=world)
trr = NativeTrainStep(ref)
for step in range(3):
    idx, n_global = dp.shard_batch_indices(perm, step, B, rank, world)
    x, y = dp.gather_frames(stream, idx, T), dp.gather_frames(target, idx, T)
    loss = tr.step(x, y, global_count=2 * n_global * T)
    ia = perm[step * B:(step + 1) * B]
    lref = trr.step(dp.gather_frames(stream, ia, T), dp.gather_frames(target, ia, T))
    pa = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    pb = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    err = (pa - pb).abs().max().item()
    allp = [torch.empty_like(pa) for _ in range(world)]
    dist.all_gather(allp, pa)
    same = all(torch.equal(allp[0], q) for q in allp)
    if rank == 0:
        print(f"step {step}: dp loss {loss.item():.8f} ref loss {lref.item():.8f} max|dp-ref| params {err:.2e} replicas identical {same}")
    assert err < 2e-6 and same
dist.destroy_process_group()
if rank == 0:
    print("dp_check ok")
