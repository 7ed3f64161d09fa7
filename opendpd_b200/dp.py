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


def frame_starts(indices, device=None, out=None):
    """This rank's frame indices (shard_batch_indices) as the int32 vector `NativeTrainStep.step_indexed` / `OdpdDims.x_starts`
    take: with stride-1 framing (data_collector.py:240-247) frame k starts at sample k, so the indices ARE the start offsets into
    the raw stream that every rank keeps resident — only these B/N integers differ between ranks and steps.
    `out` (int32, pinned host or device) is filled in place and returned; without it a new tensor is made on `device` (None = host).
    Either way `step_indexed` copies the starts into its own persistent device buffer, so no CUDA graph is keyed on this tensor."""
    idx = torch.as_tensor(indices, dtype=torch.int32)
    if out is not None:
        out.copy_(idx, non_blocking=True)
        return out
    return idx.contiguous() if device is None else idx.to(device).contiguous()


def allreduce_flat_(buf, group=None):
    """SUM all-reduce of the flat [grad | loss] buffer (2-14 KB) — NCCL over NVLink on GPUs, gloo in the CPU tests."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(buf, op=torch.distributed.ReduceOp.SUM, group=group)
    return buf


class PeerExchange:
    """Symmetric NVLink-mapped receive buffers for the fused push all-reduce + clip + AdamW kernel (include/odpd.h, csrc/dp.cu).
    Each rank owns one cudaMalloc'ed buffer of {value, step tag} words [2 parities][world sources][stride]; the 64-byte IPC handles
    are all-gathered once through torch.distributed and every peer buffer is mapped into this process."""

    def __init__(self, n_params, device, group, world, rank):
        import ctypes
        from . import _ffi
        self.L, self.n, self.world, self.rank, self.device = _ffi.lib(), int(n_params), int(world), int(rank), device
        if world > 8:
            raise _ffi.OdpdError("the fused peer exchange covers one NVSwitch node (world <= 8); use ODPD_DP_P2P=0 (NCCL) beyond that")
        nbytes = int(self.L.odpd_dp_buffer_bytes(self.n))
        own = ctypes.c_void_p()
        _ffi.check(self.L.odpd_dp_alloc(nbytes, ctypes.byref(own)))
        self.own = own.value
        hbuf = ctypes.create_string_buffer(64)
        _ffi.check(self.L.odpd_dp_ipc_handle(ctypes.c_void_p(self.own), hbuf))
        mine = torch.tensor(list(hbuf.raw), dtype=torch.uint8, device=device)
        allh = [torch.empty_like(mine) for _ in range(world)]
        torch.distributed.all_gather(allh, mine, group=group)
        self.ptrs = (ctypes.c_void_p * world)()
        self._opened = []
        for r in range(world):
            if r == rank:
                self.ptrs[r] = self.own
            else:
                peer = ctypes.c_void_p()
                _ffi.check(self.L.odpd_dp_ipc_open(bytes(allh[r].cpu().tolist()), ctypes.byref(peer)))
                self.ptrs[r] = peer.value
                self._opened.append(peer.value)
        torch.distributed.barrier(group=group)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.loss_out = torch.zeros(1, dtype=torch.float32, device=device)

    def close(self):
        import ctypes
        for p in self._opened:
            self.L.odpd_dp_ipc_close(ctypes.c_void_p(p))
        self._opened = []
        if self.own:
            self.L.odpd_dp_free(ctypes.c_void_p(self.own))
            self.own = None
