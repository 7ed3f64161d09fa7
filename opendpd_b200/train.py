"""Train-step surface: the body of the reference's `net_train` loop (modules/train_funcs.py:28-48) on the native path.

`net_train(...)` keeps the reference signature and semantics (zero_grad -> forward -> MSE -> backward -> clip -> step
-> loss.item()) and works with any torch optimizer.  `NativeTrainStep` is the fused fast path for the configuration
every script of record uses (nn.MSELoss + clip_grad_norm_ + AdamW, project.py:262-297): forward kernels with the I/Q
MSE fused, backward kernels writing one flat gradient, optional data-parallel all-reduce of that flat buffer
(SURVEY.md §8e), and clip+AdamW in a single launch (odpd_clip_adamw)."""
import ctypes
import os
import numpy as np
import torch
from torch import nn

from . import _ffi
from .functional import backbone_forward_raw, backbone_backward_raw, IqStream, _ptr, _stream
from .models import CoreModel, CascadedModel
from .dp import allreduce_flat_, PeerExchange


def net_train(log, net, dataloader, optimizer, criterion, grad_clip_val, device):
    """Drop-in for modules/train_funcs.py:16-54 (same arguments, same log side effect)."""
    net = net.train()
    losses = []
    fused = isinstance(criterion, nn.MSELoss) and criterion.reduction == "mean" and hasattr(net, "forward_mse")
    for features, targets in dataloader:
        features = features.to(device)
        targets = targets.to(device)
        optimizer.zero_grad()
        if fused:
            _, loss = net.forward_mse(features, targets)
        else:
            loss = criterion(net(features), targets)
        loss.backward()
        if grad_clip_val != 0:
            nn.utils.clip_grad_norm_(net.parameters(), grad_clip_val)
        optimizer.step()
        losses.append(loss.item())
    log["loss"] = np.mean(losses)
    return net


def net_eval(log, net, dataloader, criterion, device):
    """Drop-in for modules/train_funcs.py:57-90: forward-only pass over whole segments (B<=256, T=nperseg up to 19 662),
    returns (net, prediction, ground_truth) as numpy like the reference."""
    net = net.eval()
    fused = isinstance(criterion, nn.MSELoss) and criterion.reduction == "mean" and hasattr(net, "forward_mse")
    with torch.no_grad():
        losses, prediction, ground_truth = [], [], []
        for features, targets in dataloader:
            features = features.to(device)
            targets = targets.to(device)
            if fused:
                outputs, loss = net.forward_mse(features, targets)
            else:
                outputs = net(features)
                loss = criterion(outputs, targets)
            prediction.append(outputs.cpu())
            ground_truth.append(targets.cpu())
            losses.append(loss.item())
    log["loss"] = np.mean(losses)
    return net, torch.cat(prediction, dim=0).numpy(), torch.cat(ground_truth, dim=0).numpy()


def warmup_policy(st, wu, failed, ratio):
    """One decision of the chunk controller for one (backbone, direction).  `st` carries `clean` (consecutive clean checks), `floor`
    (shortest warm-up still allowed) and `base` (the cell's default); `wu` is the warm-up in use, `failed` whether the verify pass had
    to re-run sequences since the last check, `ratio` the worst boundary mismatch in units of the tolerance.  Returns the new
    warm-up, or None to keep the current one.  (Pure: tests/test_host_logic.py drives it without a GPU.)
      grow   x2 after a failure, or — below the default — when the mismatch exceeds half the tolerance (and never shrink below it again)
      shrink /2 after three consecutive checks under a quarter of the tolerance, not below `floor`"""
    if failed or (ratio > 0.5 and wu < st["base"]):
        st["clean"] = 0
        st["floor"] = max(st["floor"], 2 * wu)
        return 2 * wu
    if ratio < 0.25:
        st["clean"] += 1
        if st["clean"] >= 3 and wu // 2 >= st["floor"]:
            st["clean"] = 0
            return wu // 2
        return None
    st["clean"] = 0
    return None


class NativeTrainStep:
    """One fused train step for a CoreModel (train_pa, steps/train_pa.py:24-29) or a CascadedModel with frozen PA
    (train_dpd, steps/train_dpd.py:60-63).

    step(features, targets)  -> device float64 loss (no host sync); a fresh tensor every call (safe to collect and read later)
    step_host(features, targets) -> python float; inputs are (pinned) HOST tensors, H2D copies and the loss D2H read
    happen inside the call (the reference's per-batch `.to(device)` ... `loss.item()`, train_funcs.py:30-48).

    Data parallel: pass `process_group` (torch.distributed); each rank feeds its shard of the global batch, the loss
    scale uses the GLOBAL element count and one SUM all-reduce runs on [flat_grad, loss] per step."""

    def __init__(self, net, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_clip_val=200.0,
                 process_group=None, world_size=1, peer_exchange=None):
        """peer_exchange: None = auto (fused NVLink exchange when world_size>1 and env ODPD_DP_P2P != 0, else NCCL all-reduce)."""
        self.net = net
        if isinstance(net, CascadedModel):
            self.dpd, self.pa = net.dpd_model.backbone, net.pa_model.backbone
            if any(p.requires_grad for p in self.pa.parameters()):
                raise ValueError("CascadedModel: call freeze_pa_model() first (steps/train_dpd.py:63)")
            self.train_bb = self.dpd
        elif isinstance(net, CoreModel):
            self.dpd, self.pa = None, net.backbone
            self.train_bb = self.pa
        else:
            raise TypeError("NativeTrainStep wants a native CoreModel or CascadedModel")
        flat, layout = self.train_bb._flat_sync()
        if self.dpd is not None:
            self.pa._flat_sync()
        self.n = sum(n for _, n, _ in layout)
        dev = flat.device
        self.device = dev
        # [flat grad | loss] share one fp32 buffer so data-parallel needs a single all-reduce
        self.gbuf = torch.zeros(flat.numel() + 4, dtype=torch.float32, device=dev)
        self.gflat = self.gbuf[:flat.numel()]
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self.gnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.betas, self.eps, self.wd, self.clip = betas, eps, weight_decay, float(grad_clip_val)
        self.pg, self.world = process_group, int(world_size)
        self.px = None
        use_p2p = peer_exchange if peer_exchange is not None else (os.environ.get("ODPD_DP_P2P", "1") != "0")
        if self.world > 1 and self.pg is not None and use_p2p and self.n + 1 <= 4096:
            self.px = PeerExchange(self.n, dev, self.pg, self.world, torch.distributed.get_rank(self.pg))
            self._px_pin, self._px_ev = torch.zeros(1, dtype=torch.int32).pin_memory(), None
        self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        self.fuse_optimizer = os.environ.get("ODPD_FUSE_ADAMW", "1") != "0"     # clip + AdamW inside the gradient-reduction kernel (single GPU)
        self._host_step = 0
        self._stage = None
        self._bufs = [dict() for _ in range(4)]
        # self-tuning of the time-chunked kernels (include/odpd.h "Time-chunked execution"): every `chunk_check_every` steps (8) the
        # re-run counters of the verify passes are read back (asynchronously, one check late); a backbone whose chunk boundaries
        # stopped meeting — its weights moved towards a longer memory — gets its warm-up doubled, and past 1024 steps falls back
        # to the plain serial kernels.  Correctness never depends on this: failing sequences are always re-run serially.
        self.chunk_check_every = int(os.environ.get("ODPD_CHUNK_CHECK_EVERY", "8"))
        self.chunk_events = []          # [(step, cell, 'fwd'|'bwd', action)]
        self._ctl = {}
        # CUDA graphs: one step is 6-10 small launches issued through ctypes, which costs more host time (~0.2 ms) than the kernels
        # take on the device.  Steps are therefore captured once per (features, targets) buffer pair and replayed: every scalar the
        # step needs (step counter, learning rate, loss, the parity of the peer-exchange slot) lives in device memory, so a replay is
        # exact.  (The NCCL fallback of the data-parallel exchange, ODPD_DP_P2P=0, stays un-captured.)
        self.use_graphs = os.environ.get("ODPD_GRAPHS", "1") != "0" and (self.world == 1 or self.px is not None)
        self._graphs, self._graph_warm = {}, set()
        self._starts_stage = {}

    def set_lr(self, lr):
        """ReduceLROnPlateau hook (project.py:288-297) — host-side scalar write, no kernel change."""
        self.lr_dev.fill_(float(lr))

    def step(self, features, targets, global_count=None, loss_out=None):
        """One train step; returns the device float64 loss (no host sync).
        global_count: number of scalars of the GLOBAL batch (2*B_global*T); defaults to 2*B*T*world (equal shards).
        loss_out: optional 1-element float64 tensor (device or pinned host) the loss is copied into (stream-ordered, non-blocking)
        and which is then returned; without it the step returns a fresh device tensor.  (The kernels accumulate the loss in a
        buffer that the next step — or the next replay of the captured graph — overwrites, so it is never handed out itself.)"""
        B, T = features.shape[0], features.shape[1]
        if not self.use_graphs:
            loss = self._step_impl(features, targets, global_count)
        else:
            sid = lambda t: t.starts.data_ptr() if isinstance(t, IqStream) else 0
            shape_key = (tuple(features.shape), tuple(targets.shape), features.dtype, targets.dtype, global_count)
            key = (features.data_ptr(), targets.data_ptr(), sid(features), sid(targets)) + shape_key
            g = self._graphs.get(key)
            if g is None and shape_key not in self._graph_warm:
                # first step of a shape runs eagerly: allocates the cached buffers, sets kernel attributes
                self._graph_warm.add(shape_key)
                loss = self._step_impl(features, targets, global_count)
            else:
                if g is None:
                    if len(self._graphs) >= 256:
                        self._graphs.pop(next(iter(self._graphs)))
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                        gl = self._step_impl(features, targets, global_count)
                    g = self._graphs[key] = (graph, gl, features, targets)      # keeps the captured buffers alive
                g[0].replay()
                loss = g[1]
        if loss_out is not None:
            loss_out.copy_(loss, non_blocking=True)
            loss = loss_out
        else:
            loss = loss.clone()
        self._host_step += 1
        if self.chunk_check_every > 0 and self._host_step % self.chunk_check_every == 0:
            self._chunk_control(B, T)
            self._check_exchange()
        return loss

    def step_indexed(self, stream_x, stream_y, starts, T, global_count=None, loss_out=None):
        """One train step on frames addressed inside device-resident raw streams (SURVEY f-2: on-device framing inside the kernels):
        stream_x / stream_y are (N,2) fp32 or bf16 tensors, starts the B frame start indices (int32; device, or pinned host) —
        frame b is stream[starts[b] : starts[b]+T], exactly IQFrameDataset's stride-1 window (data_collector.py:233-252).  No frame
        is gathered or copied: the kernels read the stream in place.  The B indices are copied (stream-ordered) into a persistent
        device buffer owned by the trainer, so every step of a given batch size replays the SAME captured graph whatever tensor
        `starts` lives in."""
        B = int(starts.numel())
        st = self._starts_stage.get(B)
        if st is None:
            st = self._starts_stage[B] = torch.empty(B, dtype=torch.int32, device=self.device)
        if starts.dtype != torch.int32:
            raise _ffi.OdpdError("step_indexed wants int32 frame starts")
        st.copy_(starts.reshape(-1), non_blocking=True)
        return self.step(IqStream(stream_x, st, T), IqStream(stream_y, st, T), global_count, loss_out)

    def steps_indexed(self, stream_x, stream_y, starts, T, global_count=None, losses_out=None):
        """K consecutive train steps in ONE graph replay: `starts` is a (K, B) int32 tensor (device or pinned host), row k = the frame
        starts of step k (same framing as step_indexed).  Returns the K losses (device float64; or `losses_out`, a K-element float64
        tensor on the device or in pinned host memory, filled with one stream-ordered copy).
        Why: a step is ~5 kernels and ~160 us; launched one graph per step, the stream also carries a copy of the starts before and of the
        loss after every step and the GPU idles between consecutive graph launches.  K steps per replay carry one copy in, one copy out
        and no gap between the steps inside — the steps themselves are unchanged and still strictly sequential (step k+1 reads the
        parameters step k wrote).  The chunk controller is consulted once per call."""
        if starts.dim() != 2 or starts.dtype != torch.int32:
            raise _ffi.OdpdError("steps_indexed wants a (K, B) int32 tensor of frame starts")
        K, B = int(starts.shape[0]), int(starts.shape[1])
        st = self._starts_stage.get((K, B))
        if st is None:
            st = self._starts_stage[(K, B)] = (torch.empty(K, B, dtype=torch.int32, device=self.device),
                                               torch.zeros(K, dtype=torch.float64, device=self.device))
        stage, losses = st
        stage.copy_(starts, non_blocking=True)
        items = [(IqStream(stream_x, stage[k], T), IqStream(stream_y, stage[k], T)) for k in range(K)]
        self._run_multi(("idx", stream_x.data_ptr(), stream_y.data_ptr(), K, B, int(T), stream_x.dtype, stream_y.dtype), items, losses, global_count,
                        keep=(stream_x, stream_y))
        if losses_out is not None:
            losses_out.copy_(losses, non_blocking=True)
            out = losses_out
        else:
            out = losses.clone()
        before = self._host_step
        self._host_step += K
        if self.chunk_check_every > 0 and before // self.chunk_check_every != self._host_step // self.chunk_check_every:
            self._chunk_control(B, T)
            self._check_exchange()
        return out

    def _run_multi(self, key, items, losses, global_count, keep=()):
        """len(items) consecutive train steps on (features, targets) pairs whose ADDRESSES are fixed across calls with the same `key`,
        loss k -> losses[k]; eager on the first call of a key shape, captured into one CUDA graph on the second, replayed afterwards."""
        def run_all():
            for k, (f, t) in enumerate(items):
                losses[k:k + 1].copy_(self._step_impl(f, t, global_count))
        if not self.use_graphs:
            return run_all()
        key = ("multi",) + tuple(key) + (global_count,)
        g = self._graphs.get(key)
        if g is None and key not in self._graph_warm:
            self._graph_warm.add(key)
            return run_all()
        if g is None:
            if len(self._graphs) >= 256:
                self._graphs.pop(next(iter(self._graphs)))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                run_all()
            g = self._graphs[key] = (graph, losses, items, keep)       # keeps the captured buffers alive
        g[0].replay()

    def _step_impl(self, features, targets, global_count=None):
        L = _ffi.lib()
        tb = self.train_bb
        flat, _ = tb._flat_sync()
        B, T = features.shape[0], features.shape[1]
        count = float(global_count) if global_count else float(2 * B * T * self.world)   # nn.MSELoss 'mean', GLOBAL batch
        gflat = self.gflat

        single = not (self.pg is not None and self.world > 1)
        fuse_opt = single and self.fuse_optimizer

        def arm_publish(loss):
            # data parallel: the backward that produces the weight gradient publishes it (and the loss) to every rank straight from
            # the epilogue of its gradient reduction (csrc/api.cu reduce_partials_kernel, include/odpd.h odpd_dp_publish_next_bwd);
            # single GPU: the same reduction kernel runs clip + AdamW in its last CTA (odpd_fuse_next_bwd_with_adamw): one launch less
            if self.px is not None:
                _ffi.check(L.odpd_dp_publish_next_bwd(self.px.ptrs, self.world, self.px.rank, ctypes.c_int64(self.n), _ptr(self.step_dev), _ptr(loss)))
            elif fuse_opt:
                _ffi.check(L.odpd_fuse_next_bwd_with_adamw(_ptr(flat), _ptr(self.exp_avg), _ptr(self.exp_avg_sq), _ptr(self.lr_dev), self.betas[0],
                                                           self.betas[1], self.eps, self.wd, self.clip, _ptr(self.step_dev), _ptr(self.gnorm),
                                                           _ptr(self._ticket)))
        if self.dpd is None:
            spec = tb._spec()
            out, loss, saved = backbone_forward_raw(spec, features, flat, targets, 1.0 / count, True, tb._stats_tensor(self.device),
                                                    self._bufs[0])
            arm_publish(loss)
            backbone_backward_raw(spec, features, flat, saved, False, True, out=out, target=targets, gscale=2.0 / count,
                                  gflat=gflat, bufs=self._bufs[1])
        else:
            paflat, _ = self.pa._flat_sync()
            sd, sp = self.dpd._spec(), self.pa._spec()
            mid, _, saved_d = backbone_forward_raw(sd, features, flat, None, 0.0, True, self.dpd._stats_tensor(self.device), self._bufs[0])
            out, loss, saved_p = backbone_forward_raw(sp, mid, paflat, targets, 1.0 / count, True, self.pa._stats_tensor(self.device),
                                                      self._bufs[2])
            gmid, _ = backbone_backward_raw(sp, mid, paflat, saved_p, True, False, out=out, target=targets, gscale=2.0 / count,
                                            bufs=self._bufs[3])
            arm_publish(loss)
            backbone_backward_raw(sd, features, flat, saved_d, False, True, gout=gmid, gflat=gflat, bufs=self._bufs[1])
        if self.px is not None:
            # fused: gather of the pushed gradients from the LOCAL receive buffer + ordered sum + clip + AdamW in ONE kernel (csrc/dp.cu)
            _ffi.check(L.odpd_dp_clip_adamw(_ptr(flat), self.px.ptrs, self.world, self.px.rank, ctypes.c_int64(self.n), None, _ptr(loss),
                                            _ptr(self.exp_avg), _ptr(self.exp_avg_sq), _ptr(self.lr_dev), self.betas[0], self.betas[1],
                                            self.eps, self.wd, self.clip, _ptr(self.step_dev), _ptr(self.gnorm), _ptr(self.px.loss_out),
                                            _ptr(self.px.status), _stream()))
            return self.px.loss_out.to(torch.float64)
        if fuse_opt:
            return loss
        if self.pg is not None and self.world > 1:
            self.gbuf[-4] = loss.to(torch.float32)[0]
            allreduce_flat_(self.gbuf, self.pg)
            loss = self.gbuf[-4:-3].to(torch.float64)
        _ffi.check(L.odpd_clip_adamw(_ptr(flat), _ptr(self.gflat), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                                     ctypes.c_int64(self.n), _ptr(self.lr_dev), self.betas[0], self.betas[1], self.eps, self.wd,
                                     self.clip, _ptr(self.step_dev), _ptr(self.gnorm), 0, _stream()))
        return loss

    def _check_exchange(self, wait=False):
        """The fused peer exchange reports a peer that never published through a device status word and skips that step's update
        (the kernel does not hang).  Read it back asynchronously on the chunk-controller cadence (pinned copy + event, one check
        late) and fail loudly: a silent skip would leave the replicas diverged.  wait=True synchronises (end of a run)."""
        if self.px is None or torch.cuda.is_current_stream_capturing():
            return
        for _ in range(2 if wait else 1):
            if self._px_ev is None:
                self._px_pin.copy_(self.px.status, non_blocking=True)
                self._px_ev = torch.cuda.Event()
                self._px_ev.record()
                if not wait:
                    return
            if wait or self._px_ev.query():
                self._px_ev.synchronize()
                bad, self._px_ev = int(self._px_pin[0]), None
                if bad:
                    raise _ffi.OdpdError(f"NVLink peer exchange: rank {bad - 1} did not publish its gradient within 2 s (around step "
                                         f"{self._host_step}); this rank skipped the update - replicas are no longer in sync. "
                                         "Re-run with ODPD_DP_P2P=0 to use the NCCL all-reduce")

    def chunk_calls(self):
        """The backbone launches of one step: (module, backward?, index into self._bufs, forward saves?, backward needs dW?)."""
        if self.dpd is None:
            return [(self.train_bb, False, 0, True, True), (self.train_bb, True, 1, True, True)]
        return [(self.dpd, False, 0, True, True), (self.pa, False, 2, True, True), (self.pa, True, 3, True, False), (self.dpd, True, 1, True, True)]

    def _chunk_control(self, B, T):
        """Self-tuning of the warm-up of the time-chunked kernels, per backbone and direction:
          grow   x2 when the verify pass had to re-run sequences, or — below the cell's default — when the worst boundary mismatch
                 exceeds half the tolerance; past 1024 steps that direction falls back to the plain serial kernel
          shrink /2 (not below 64, never below a length that once failed) after three consecutive checks whose worst mismatch
                 stayed under a quarter of the tolerance
        Telemetry is read back asynchronously one check late; results never depend on it (failing sequences are re-run serially)."""
        for mod, backward, bi, save, need_dw in self.chunk_calls():
            buf = self._bufs[bi].get("ws" if backward else "saved")
            if buf is None:
                continue
            spec = mod._spec()
            plan = spec.chunk_plan(B, T, backward, save, need_dw)
            if plan[0] <= 1 or plan[3] < 0:
                continue
            k, name = (1, "bwd") if backward else (0, "fwd")
            st = self._ctl.setdefault((id(mod), backward), dict(pin=torch.zeros(2, dtype=torch.int32).pin_memory(), ev=None, seen=0, ptr=0,
                                                                  clean=0, floor=64, base=plan[2], rebase=False))
            if st["ptr"] != buf.data_ptr():                       # buffer re-allocated (shape or chunk count changed): counters restart
                st.update(ptr=buf.data_ptr(), ev=None, seen=0, rebase=False)
            new, failed = None, False
            if st["ev"] is not None:
                st["ev"].synchronize()                            # copy enqueued a whole check interval ago
                cnt, ratio = int(st["pin"][0]), float(st["pin"][1:2].view(torch.float32)[0])
                failed = cnt > st["seen"] and not st["rebase"]    # re-runs counted before the last plan change do not count
                st["seen"], st["rebase"] = cnt, False
                new = warmup_policy(st, plan[2], failed, ratio)
            if new is not None:
                st["clean"] = 0
                self._graphs.clear()                              # captured launches carry the old plan
                if new <= 1024 and new <= T // 2:
                    tw = list(spec.twarm)
                    tw[k] = new
                    mod.time_warmup = tuple(tw)
                    self.chunk_events.append((self._host_step, mod.cell, name, f"warm-up {plan[2]} -> {new}" + (" (boundary failed)" if failed else "")))
                else:
                    tc = list(spec.tchunks)
                    tc[k] = 1
                    mod.time_chunks = tuple(tc)
                    self.chunk_events.append((self._host_step, mod.cell, name, "serial"))
                buf[plan[3] + 1:plan[3] + 2].zero_()              # restart the worst-mismatch telemetry under the new plan
                st.update(ev=None, rebase=True)
                continue
            st["pin"].copy_(buf[plan[3]:plan[3] + 2].view(torch.int32), non_blocking=True)
            st["ev"] = torch.cuda.Event()
            st["ev"].record()

    def step_host(self, features_cpu, targets_cpu):
        if self._stage is None or self._stage[0].shape != features_cpu.shape or self._stage[0].dtype != features_cpu.dtype:
            self._stage = (torch.empty(features_cpu.shape, dtype=features_cpu.dtype, device=self.device),     # fp32 or bf16 storage
                           torch.empty(targets_cpu.shape, dtype=targets_cpu.dtype, device=self.device))
        fx, fy = self._stage
        fx.copy_(features_cpu, non_blocking=True)
        fy.copy_(targets_cpu, non_blocking=True)
        return float(self.step(fx, fy).item())

    def run_host_batches(self, batches, steps_per_replay=None):
        """Train on an iterable of HOST (pinned) (features, targets) batches — the reference's `for features, targets in dataloader`
        loop (train_funcs.py:28-48) as a 2-deep pipeline of BLOCKS of `steps_per_replay` (default 8, env ODPD_STEPS_PER_REPLAY) steps:
        block j+1 is copied host->device on a side stream while block j computes as ONE CUDA-graph replay of its steps (block sizes ramp
        1, 2, 4 ... so that only one batch copy is ever exposed), and the losses
        of block j are read back (async D2H, pinned) while block j+1 is already queued.  Every step still has its own H2D copy of its
        batch and its own loss read-back inside the region; they no longer serialise with the kernels, and there is no launch gap
        between the steps of a block.  Returns the list of losses."""
        K = int(steps_per_replay or os.environ.get("ODPD_STEPS_PER_REPLAY", "8"))
        cur = torch.cuda.current_stream()
        if not hasattr(self, "_cs"):
            self._cs = torch.cuda.Stream(device=self.device)
            self._pipe = None
        cs = self._cs
        it = iter(batches)

        ramp = [1]       # block sizes 1, 2, 4, ... K: the first block's copy is the only one nothing overlaps, so it is kept to one batch

        def take_block():
            size = min(ramp[0], K)
            ramp[0] = min(2 * ramp[0], K)
            blk = []
            for b_ in it:
                blk.append(b_)
                if len(blk) == size:
                    break
            return blk
        nxt = take_block()
        if not nxt:
            return []
        shp = (tuple(nxt[0][0].shape), tuple(nxt[0][1].shape), nxt[0][0].dtype, nxt[0][1].dtype, K)
        if self._pipe is None or self._pipe["shape"] != shp:
            mk = lambda s_, dt: torch.empty(s_, dtype=dt, device=self.device)
            self._pipe = dict(shape=shp, stage=[[(mk(shp[0], shp[2]), mk(shp[1], shp[3])) for _ in range(K)] for _ in range(2)],
                              dloss=[torch.zeros(K, dtype=torch.float64, device=self.device) for _ in range(2)],
                              loss=[torch.zeros(K, dtype=torch.float64).pin_memory() for _ in range(2)],
                              ev_copy=[torch.cuda.Event() for _ in range(2)], ev_free=[torch.cuda.Event() for _ in range(2)],
                              ev_loss=[torch.cuda.Event() for _ in range(2)])
        P = self._pipe
        used = [False, False]

        def enqueue_copy(slot, blk):
            with torch.cuda.stream(cs):
                if used[slot]:
                    cs.wait_event(P["ev_free"][slot])
                else:
                    cs.wait_stream(cur)
                for k, (f, t) in enumerate(blk):
                    P["stage"][slot][k][0].copy_(f, non_blocking=True)
                    P["stage"][slot][k][1].copy_(t, non_blocking=True)
                P["ev_copy"][slot].record(cs)

        enqueue_copy(0, nxt)
        losses, pending, i = [], None, 0
        while nxt:
            slot, blk = i & 1, nxt
            nxt = take_block()
            if nxt:
                enqueue_copy(slot ^ 1, nxt)
            cur.wait_event(P["ev_copy"][slot])
            n = len(blk)
            B, T = blk[0][0].shape[0], blk[0][0].shape[1]
            self._run_multi(("host", slot, n) + shp[:4], P["stage"][slot][:n], P["dloss"][slot], None)
            P["ev_free"][slot].record(cur)
            used[slot] = True
            P["loss"][slot][:n].copy_(P["dloss"][slot][:n], non_blocking=True)      # async D2H of the block's losses
            P["ev_loss"][slot].record(cur)
            before = self._host_step
            self._host_step += n
            if self.chunk_check_every > 0 and before // self.chunk_check_every != self._host_step // self.chunk_check_every:
                self._chunk_control(B, T)
                self._check_exchange()
            if pending is not None:
                P["ev_loss"][pending[0]].synchronize()
                losses += [float(v) for v in P["loss"][pending[0]][:pending[1]]]
            pending = (slot, n)
            i += 1
        P["ev_loss"][pending[0]].synchronize()
        losses += [float(v) for v in P["loss"][pending[0]][:pending[1]]]
        return losses

    def grads_as_param_grads(self):
        """Expose the flat gradient through each parameter's .grad (views), e.g. for inspection or a torch optimizer."""
        _, layout = self.train_bb._flat_sync()
        for (_, p), (off, n, shape) in zip(self.train_bb.named_parameters(), layout):
            p.grad = self.gflat[off:off + n].view(shape)
