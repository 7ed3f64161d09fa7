"""torchrun --nproc-per-node N scripts/dp_check.py : data-parallel NativeTrainStep == single-process full-batch training.
Covers: equal shards, a last partial batch whose shards differ in size, a partial batch smaller than the world (some ranks get an
EMPTY shard and must contribute a zero gradient), on-device framing (step_indexed), replicas bit-identical throughout."""
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
n_frames = N - T + 1
tr = NativeTrainStep(net, process_group=dist.group.WORLD, world_size=world)
trr = NativeTrainStep(ref)


def check(tag, loss, lref):
    pa = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    pb = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    err = (pa - pb).abs().max().item()
    allp = [torch.empty_like(pa) for _ in range(world)]
    dist.all_gather(allp, pa)
    same = all(torch.equal(allp[0], q) for q in allp)
    if rank == 0:
        print(f"{tag}: dp loss {loss.item():.8f} ref loss {lref.item():.8f} max|dp-ref| params {err:.2e} replicas identical {same}", flush=True)
    assert err < 2e-6 and same and abs(loss.item() - lref.item()) <= 2e-6 * abs(lref.item()), (tag, err, same)


# a "dataset" of 2*B + (B//2+1) + 1 frames: two full batches, one uneven partial batch, one single-frame batch (empty shards)
perm = dp.epoch_permutation(n_frames, seed=0)[:2 * B + B // 2 + 1 + 1]
bounds = [(0, B), (B, 2 * B), (2 * B, 2 * B + B // 2 + 1), (2 * B + B // 2 + 1, 2 * B + B // 2 + 2)]
for step, (lo, hi) in enumerate(bounds):
    n_global = hi - lo
    base, rem = divmod(n_global, world)
    start = lo + rank * base + min(rank, rem)
    idx = perm[start:start + base + (1 if rank < rem else 0)]
    ia = perm[lo:hi]
    if step % 2 == 0:
        loss = tr.step(dp.gather_frames(stream, idx, T), dp.gather_frames(target, idx, T), global_count=2 * n_global * T)
    else:   # on-device framing: the kernels read the raw stream through frame starts
        loss = tr.step_indexed(stream, target, dp.frame_starts(idx, dev), T, global_count=2 * n_global * T)
    lref = trr.step(dp.gather_frames(stream, ia, T), dp.gather_frames(target, ia, T))
    check(f"step {step} (global batch {n_global}, this shard {idx.numel()})", loss, lref)
tr._check_exchange(wait=True)
dist.barrier()
if tr.px is not None:
    tr.px.close()
dist.destroy_process_group()
if rank == 0:
    print("dp_check ok")
