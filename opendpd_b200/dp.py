"""Data-parallel host logic: one process per GPU, parameters replicated, each rank takes a contiguous
slice of the SAME seeded permutation the reference's DataLoader(shuffle=True) would draw (project.py:236, seeded by
project.py:111), the loss uses the GLOBAL element count, and one SUM all-reduce runs on the flat [grad | loss] buffer."""
import torch


def epoch_permutation(n_frames, seed, epoch=0):
    """The frame order of one epoch — same construction as torch's RandomSampler with a seeded generator."""
    g = torch.Generator()
    g.manual_seed(int(seed) + int(epoch))
    return torch.randperm(n_frames, generator=g)


def shard_batch_indices(perm, step, global_batch, rank, world):
    """Indices (into the frame list) of this rank's share of global batch `step`.  The union over ranks is exactly the
    single-process batch perm[step*B:(step+1)*B]; the last partial batch is split as evenly as possible (sizes differ by <=1)."""
    lo = step * global_batch
    hi = min(lo + global_batch, perm.numel())
    n = max(hi - lo, 0)
    base, rem = divmod(n, world)
    start = lo + rank * base + min(rank, rem)
    return perm[start:start + base + (1 if rank < rem else 0)], n


def gather_frames(stream, starts, frame_length):
    """On-device framing (replaces IQFrameDataset's materialised frames, data_collector.py:233-252): stream (N,2) resident
    on the device, frame k = rows [k, k+T)."""
    idx = starts.to(stream.device).view(-1, 1) + torch.arange(frame_length, device=stream.device).view(1, -1)
    return stream[idx]


def allreduce_flat_(buf, group=None):
    """SUM all-reduce of the flat [grad | loss] buffer (2-14 KB) — NCCL over NVLink on GPUs, gloo in the CPU tests."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(buf, op=torch.distributed.ReduceOp.SUM, group=group)
    return buf
