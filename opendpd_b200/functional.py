"""Autograd bridge: one torch.autograd.Function that runs a whole backbone forward (optionally + fused I/Q MSE)
and its backward through libodpd.so.  PyTorch is plumbing here (allocation, streams, autograd graph edges)."""
import ctypes
import os
import torch
from . import _ffi


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class CellSpec:
    """Static description of one native backbone call."""

    def __init__(self, cell, H, K=0, thx=0.0, thh=0.0, tchunks=None, twarm=None):
        """tchunks / twarm: OdpdDims.tchunks / .twarm (include/odpd.h "Time-chunked execution"); None = environment
        ODPD_TCHUNKS[_FWD|_BWD] / ODPD_TWARM, else 0 = the library picks.  Both may be an int or a (forward, backward) pair."""
        self.cell, self.H, self.K, self.thx, self.thh = cell, int(H), int(K or 0), float(thx), float(thh)
        self.cell_id = _ffi.CELLS[cell]
        self.keep_saved = None
        env = os.environ.get
        if tchunks is None:
            both = env("ODPD_TCHUNKS", "0")
            tchunks = (int(env("ODPD_TCHUNKS_FWD", both)), int(env("ODPD_TCHUNKS_BWD", both)))
        elif isinstance(tchunks, int):
            tchunks = (tchunks, tchunks)
        self.tchunks = tuple(int(v) for v in tchunks)
        if twarm is None:
            twarm = int(env("ODPD_TWARM", "0"))
        self.twarm = (int(twarm), int(twarm)) if isinstance(twarm, int) else tuple(int(v) for v in twarm)   # (forward, backward)

    def dims(self, B, T, flags, backward=False, x_starts=None, target_starts=None):
        k = 1 if backward else 0
        twarm = self.twarm if isinstance(self.twarm, int) else self.twarm[k]
        return _ffi.OdpdDims(self.cell_id, int(B), int(T), self.H, self.K, int(flags), self.thx, self.thh, self.tchunks[k], twarm,
                             None if x_starts is None else x_starts.data_ptr(), None if target_starts is None else target_starts.data_ptr())

    def chunk_plan(self, B, T, backward=False, save=True, need_dw=True):
        """(chunks, steps per chunk, warm-up steps, index of the re-run counter) the library will use for this call shape."""
        flags = (_ffi.F_NEED_DW if need_dw else 0) if backward else (_ffi.F_SAVE if save else 0)
        d = self.dims(B, T, flags, backward)
        out = (ctypes.c_int32 * 4)()
        _ffi.check(_ffi.lib().odpd_chunk_plan(ctypes.byref(d), 1 if backward else 0, out))
        return tuple(int(v) for v in out)


class IqStream:
    """On-device framing (OdpdDims.x_starts / .target_starts): instead of a framed (B,T,2) tensor, hand the kernels the raw (N,2) IQ
    stream (fp32 or bf16) plus the B frame start indices (device int32) — the reference's IQFrameDataset materialises every stride-1
    frame (data_collector.py:233-252) and copies B*T*2 floats per step; this copies B indices."""

    def __init__(self, stream, starts, T):
        if stream.dim() != 2 or stream.size(-1) != 2 or not stream.is_contiguous():
            raise _ffi.OdpdError("IqStream wants a contiguous (N,2) stream")
        if starts.dtype != torch.int32 or starts.dim() != 1 or not starts.is_contiguous() or starts.device != stream.device:
            raise _ffi.OdpdError("IqStream wants contiguous int32 frame starts on the stream's device")
        self.stream, self.starts, self.T = stream, starts, int(T)
        self.shape = (starts.numel(), self.T, 2)
        self.device, self.dtype, self.is_cuda = stream.device, stream.dtype, stream.is_cuda

    def data_ptr(self):
        return self.stream.data_ptr()

    def frames(self):
        """Materialised (B,T,2) frames (tests / fallbacks)."""
        idx = self.starts.long()[:, None] + torch.arange(self.T, device=self.device)[None, :]
        return self.stream[idx]


def _iq(t):
    """(tensor whose data_ptr the kernel reads, bf16 flag, starts tensor or None) for a framed tensor or an IqStream."""
    if t is None:
        return None, False, None
    if isinstance(t, IqStream):
        return t.stream, t.stream.dtype == torch.bfloat16, t.starts
    return t, t.dtype == torch.bfloat16, None


def _check_x(x):
    if not x.is_cuda:
        raise _ffi.OdpdError("native backbones run on CUDA tensors only (no CPU fallback); got a CPU tensor")
    if isinstance(x, IqStream):
        return x
    if x.dtype not in (torch.float32, torch.bfloat16) or x.dim() != 3 or x.size(-1) != 2:
        raise _ffi.OdpdError(f"expected a float32 / bfloat16 (B,T,2) tensor, got {tuple(x.shape)} {x.dtype}")
    return x.contiguous()


def backbone_forward_raw(spec, x, flat, target=None, loss_scale=0.0, save=True, stats=None, bufs=None):
    """Launch the forward kernel.  Returns (out, loss_double_or_None, saved_or_None).
    `bufs` (dict) caches out/saved/loss allocations across calls of identical shape (NativeTrainStep)."""
    L = _ffi.lib()
    B, T = x.shape[0], x.shape[1]
    xt, xbf, xst = _iq(x)
    tt, tbf, tst = _iq(target)
    d = spec.dims(B, T, (_ffi.F_SAVE if save else 0) | _ffi.F_ZERO_LOSS | (_ffi.F_X_BF16 if xbf else 0) | (_ffi.F_TARGET_BF16 if tbf else 0),
                  x_starts=xst, target_starts=tst)
    key = (B, T, bool(save), target is not None, spec.tchunks[0])
    if bufs is not None and bufs.get("key") == key:
        out, saved, loss = bufs["out"], bufs["saved"], bufs["loss"]
    else:
        out = torch.empty((B, T, 2), dtype=torch.float32, device=x.device)
        saved = None
        nbytes = L.odpd_saved_bytes(ctypes.byref(d))     # activations (when saving) + chunk scratch
        if nbytes < 0:
            _ffi.check(-1)
        if save or nbytes > 0:
            saved = torch.empty(max(int(nbytes) // 4, 1), dtype=torch.float32, device=x.device)
            idx = spec.chunk_plan(B, T, False, save)[3]
            if idx >= 0:
                saved[idx:idx + 2].zero_()               # re-run counter + worst boundary mismatch of the verify pass
        loss = torch.empty(1, dtype=torch.float64, device=x.device) if target is not None else None
        if bufs is not None:
            bufs.update(key=key, out=out, saved=saved, loss=loss)
    _ffi.check(L.odpd_backbone_fwd(ctypes.byref(d), _ptr(xt), _ptr(tt), _ptr(flat), _ptr(out), _ptr(loss),
                                   ctypes.c_double(loss_scale), _ptr(saved), _ptr(stats), _stream()))
    return out, loss, saved


def backbone_backward_raw(spec, x, flat, saved, need_dx, need_dw, gout=None, out=None, target=None, gscale=0.0,
                          gscale_dev=None, gflat=None, bufs=None):
    """Launch the backward kernel (+ ordered partial reduction).  A caller-supplied gflat is OVERWRITTEN with the
    parameter gradient (ODPD_F_OVERWRITE_DW).  Returns (gx, gflat)."""
    L = _ffi.lib()
    B, T = x.shape[0], x.shape[1]
    xt, xbf, xst = _iq(x)
    tt, tbf, tst = _iq(target)
    flags = ((_ffi.F_NEED_DX if need_dx else 0) | (_ffi.F_NEED_DW if need_dw else 0) | _ffi.F_OVERWRITE_DW |
             (_ffi.F_X_BF16 if xbf else 0) | (_ffi.F_TARGET_BF16 if tbf else 0))
    d = spec.dims(B, T, flags, backward=True, x_starts=xst, target_starts=tst)
    key = (B, T, bool(need_dx), bool(need_dw), spec.tchunks[1])
    if bufs is not None and bufs.get("key") == key:
        gx, ws = bufs["gx"], bufs["ws"]
    else:
        gx = torch.empty((B, T, 2), dtype=torch.float32, device=x.device) if need_dx else None
        ws = torch.empty(int(L.odpd_bwd_workspace_bytes(ctypes.byref(d))) // 4, dtype=torch.float32, device=x.device)
        idx = spec.chunk_plan(B, T, True, True, need_dw)[3]
        if idx >= 0:
            ws[idx:idx + 2].zero_()                      # re-run counter + worst boundary mismatch of the verify pass
        if bufs is not None:
            bufs.update(key=key, gx=gx, ws=ws)
    if need_dw and gflat is None:
        gflat = torch.empty_like(flat)
    _ffi.check(L.odpd_backbone_bwd(ctypes.byref(d), _ptr(xt), _ptr(flat), _ptr(saved), _ptr(gout), _ptr(out), _ptr(tt),
                                   ctypes.c_double(gscale), _ptr(gscale_dev), _ptr(gx), _ptr(gflat), _ptr(ws), _stream()))
    return gx, gflat


def chunk_reruns(spec, buf, B, T, backward=False, save=True, need_dw=True):
    """Number of sequences the verify pass re-ran serially since `buf` (saved / workspace) was allocated (host sync)."""
    idx = spec.chunk_plan(B, T, backward, save, need_dw)[3]
    if idx < 0 or buf is None:
        return 0
    return int(buf[idx:idx + 1].view(torch.int32).item())


def chunk_worst_mismatch(spec, buf, B, T, backward=False, save=True, need_dw=True):
    """Largest chunk-boundary mismatch the verify pass has seen since `buf` was allocated, in units of its tolerance
    (<= 1 passes; host sync)."""
    idx = spec.chunk_plan(B, T, backward, save, need_dw)[3]
    if idx < 0 or buf is None:
        return 0.0
    return float(buf[idx + 1:idx + 2].item())


class BackboneFn(torch.autograd.Function):
    """out, loss = BackboneFn.apply(spec, layout, flat, stats, target, loss_count, x, *params)

    `params` are the module's nn.Parameters (views into `flat`, see FlatParams) — passed so autograd routes their
    gradients; the kernels read `flat`.  With `target` the I/Q MSE (nn.MSELoss 'mean' over loss_count scalars) is
    fused into the forward and its gradient into the backward."""

    @staticmethod
    def forward(ctx, spec, layout, flat, stats, target, loss_count, x, *params):
        x = _check_x(x)
        if target is not None:
            target = _check_x(target)
        if isinstance(x, IqStream) or isinstance(target, IqStream):
            raise _ffi.OdpdError("IqStream inputs go through NativeTrainStep.step / the raw wrappers, not through autograd")
        nig = ctx.needs_input_grad  # (spec, layout, flat, stats, target, loss_count, x, *params)
        need_dx = bool(nig[6])
        need_dw = any(nig[7:])
        save = need_dx or need_dw
        lc = float(loss_count) if loss_count else float(x.numel())
        out, loss_d, saved = backbone_forward_raw(spec, x, flat, target, 1.0 / lc if target is not None else 0.0, save, stats)
        if spec.keep_saved is not None:       # debugging hook (delta masks for the parity tests)
            spec.keep_saved._last_saved = (saved, x.shape[0], x.shape[1])
        ctx.spec, ctx.layout, ctx.flat, ctx.saved, ctx.lc = spec, layout, flat, saved, lc
        ctx.need = (need_dx, need_dw, [bool(v) for v in nig[7:]])
        ctx.save_for_backward(x, out, target)
        ctx.set_materialize_grads(False)
        loss = loss_d.to(torch.float32).reshape(()) if loss_d is not None else None
        return out, loss

    @staticmethod
    def backward(ctx, g_out, g_loss):
        x, out, target = ctx.saved_tensors
        need_dx, need_dw, pmask = ctx.need
        if ctx.saved is None:
            raise _ffi.OdpdError("backward called but the forward ran without saving activations")
        gout, gscale, gscale_dev = None, 0.0, None
        if g_loss is not None and target is not None:
            gscale = 2.0 / ctx.lc
            gscale_dev = g_loss.to(torch.float32).contiguous()
            if g_out is not None:  # rare: both heads used downstream -> materialise the sum
                gout = (g_out + gscale_dev * gscale * (out - target)).contiguous()
                gscale_dev = None
        elif g_out is not None:
            gout = g_out.contiguous()
        else:
            return (None,) * (7 + len(pmask))
        gx, gflat = backbone_backward_raw(ctx.spec, x, ctx.flat, ctx.saved, need_dx, need_dw, gout=gout,
                                          out=out if gout is None else None, target=target if gout is None else None,
                                          gscale=gscale, gscale_dev=gscale_dev)
        if gx is not None and gx.dtype != x.dtype:
            gx = gx.to(x.dtype)
        gparams = [None] * len(pmask)
        if need_dw:
            for i, (off, n, shape) in enumerate(ctx.layout):
                if pmask[i]:
                    gparams[i] = gflat[off:off + n].view(shape)
        return (None, None, None, None, None, None, gx, *gparams)
