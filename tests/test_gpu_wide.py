"""GPU parity of the layered RNN path (csrc/wide.cu): hidden sizes 33..64 and / or num_layers > 1 for GRU, LSTM, DGRU, QGRU and
QGRU_AMP1 — the part of the reference's command line (arguments.py:51,60 -> nn.GRU / nn.LSTM num_layers, any hidden size) that
the fused one-layer kernels do not cover.  Checked against
  (1) goldens of the unmodified reference (tests/golden/wide_*.npz, oracle/make_golden.py), and
  (2) on seeded inputs at larger and ragged sizes, the PyTorch-ATen restatement oracle/torch_port.forward_layers in fp64 (itself
      pinned to those goldens by tests/test_torch_port.py), with its fp32 run giving the conditioning of the case."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, rel_err, tol_for, assert_close, note_achieved, wide_cases
from tests.test_gpu_parity import grads_flat

pytestmark = pytest.mark.gpu


def build_native(g, device="cuda"):
    from opendpd_b200 import models
    net = models.CoreModel(2, g["H"], g["L"], g["kind"])
    assert [n for n, _ in net.backbone.named_parameters()] == [n for n, _ in g["param_index"]], "parameter names/order differ from the reference"
    off = 0
    with torch.no_grad():
        for (_, p), (_, shape) in zip(net.backbone.named_parameters(), g["param_index"]):
            n = int(np.prod(shape))
            assert list(p.shape) == shape
            p.copy_(torch.from_numpy(g["params"][off:off + n]).view(shape))
            off += n
    return net.to(device)


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", wide_cases())
def test_wide_golden_parity(name, fused):
    g = load_golden(name)
    net = build_native(g)
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    y = torch.from_numpy(g["y"]).cuda()
    if fused:
        out, loss = net.forward_mse(x, y)
    else:
        out = net(x)
        loss = torch.nn.MSELoss()(out, y)
    loss.backward()
    torch.cuda.synchronize()
    errs = dict(out=rel_err(out.detach().cpu().numpy(), g["out"]), gx=rel_err(x.grad.cpu().numpy(), g["gx"]),
                gparams=rel_err(grads_flat(net), g["gparams"]), loss=abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])))
    note_achieved(name, **errs, tol_out=tol_for(g, "out"), tol_gx=tol_for(g, "gx"), tol_gparams=tol_for(g, "gparams"))
    assert errs["out"] < tol_for(g, "out"), errs
    assert errs["loss"] <= 1e-5, errs
    assert errs["gx"] < tol_for(g, "gx"), errs
    assert errs["gparams"] < tol_for(g, "gparams"), errs


def _seeded(kind, H, L, B, T, seed=7):
    from opendpd_b200 import models
    torch.manual_seed(1234 + H + 100 * L)
    net = models.CoreModel(2, H, L, kind).cuda()
    gen = torch.Generator().manual_seed(seed)
    xc = (0.2 * torch.randn(B, T, 2, generator=gen)).clamp(-0.7, 0.7)
    yc = xc * (1 - 0.2 * (xc ** 2).sum(-1, keepdim=True))
    return net, xc, yc


def _port(kind, H, L, xc, yc, params, dtype):
    from oracle import torch_port
    torch.set_num_threads(8)
    flat = torch.tensor(params, dtype=dtype, requires_grad=True)
    x = xc.to(dtype).clone().requires_grad_(True)
    out = torch_port.forward_layers(kind, x, flat, H, L)
    loss = torch.nn.MSELoss()(out, yc.to(dtype))
    loss.backward()
    return dict(out=out.detach().numpy(), gx=x.grad.numpy(), gparams=flat.grad.numpy(), loss=float(loss.detach()))


def _q_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    e = np.abs(a - b) / (np.abs(b).max() + 1e-300)
    return float(np.quantile(e, 0.9999)) if e.size >= 10000 else float(e.max())


@pytest.mark.parametrize("kind,H,L,B,T", [("gru", 64, 1, 5, 300), ("gru", 33, 2, 4, 97), ("dgru", 48, 1, 6, 257), ("dgru", 13, 2, 9, 130),
                                          ("lstm", 64, 2, 3, 161), ("lstm", 40, 1, 7, 64), ("qgru", 50, 1, 4, 50), ("qgru_amp1", 10, 3, 5, 75),
                                          ("gru", 8, 8, 2, 40), ("dgru", 64, 2, 64, 512)])
def test_wide_seeded_parity(kind, H, L, B, T):
    net, xc, yc = _seeded(kind, H, L, B, T)
    x = xc.cuda().requires_grad_(True)
    out, loss = net.forward_mse(x, yc.cuda())
    loss.backward()
    torch.cuda.synchronize()
    params = np.concatenate([p.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])
    r64 = _port(kind, H, L, xc, yc, params, torch.float64)
    r32 = _port(kind, H, L, xc, yc, params, torch.float32)
    for key, mine in (("out", out.detach().cpu().numpy()), ("gx", x.grad.cpu().numpy()), ("gparams", grads_flat(net))):
        tol = max(1e-5, 3 * _q_err(r32[key], r64[key]))
        assert_close(mine, r64[key], tol, f"{kind} H{H} L{L} {key}")
    assert abs(loss.item() - r64["loss"]) <= 1e-5 * abs(r64["loss"])


@pytest.mark.parametrize("kind,H,L", [("dgru", 40, 2), ("lstm", 48, 1)])
def test_wide_dx_only_and_inference(kind, H, L):
    """Frozen parameters (the PA of train_dpd, models.py:169-171): the dX-only backward equals the full one; a forward under no_grad
    (no saved activations: only h per layer is kept) gives the same output bit for bit."""
    net, xc, yc = _seeded(kind, H, L, 4, 150)
    x = xc.cuda().requires_grad_(True)
    out = net(x)
    loss = torch.nn.MSELoss()(out, yc.cuda())
    loss.backward()
    gx_full = x.grad.clone()
    with torch.no_grad():
        out_inf = net(xc.cuda())
    assert torch.equal(out_inf, out.detach())
    for p in net.parameters():
        p.requires_grad_(False)
    x2 = xc.cuda().requires_grad_(True)
    loss2 = torch.nn.MSELoss()(net(x2), yc.cuda())
    loss2.backward()
    assert torch.equal(x2.grad, gx_full)


def test_wide_on_device_framing_and_bf16_storage():
    """The IQ storage options of the C ABI (frame starts into the raw stream, bf16 pairs) reach the layered kernels too."""
    from opendpd_b200.functional import IqStream, backbone_forward_raw, backbone_backward_raw
    net, _, _ = _seeded("dgru", 36, 2, 1, 32)
    bb = net.backbone
    gen = torch.Generator().manual_seed(3)
    N, B, T = 4000, 6, 200
    stream = (0.25 * torch.randn(N, 2, generator=gen)).clamp(-0.7, 0.7).to(torch.bfloat16)
    tgt = (stream.float() * 0.9).to(torch.bfloat16)
    starts = torch.randint(0, N - T, (B,), generator=gen, dtype=torch.int32)
    xs, ys = IqStream(stream.cuda(), starts.cuda(), T), IqStream(tgt.cuda(), starts.cuda(), T)
    flat, _ = bb._flat_sync()
    spec = bb._spec()
    scale = 1.0 / (B * T * 2)
    out_i, loss_i, saved_i = backbone_forward_raw(spec, xs, flat, ys, scale, True)
    gx_i, gw_i = backbone_backward_raw(spec, xs, flat, saved_i, True, True, out=out_i, target=ys, gscale=2 * scale)
    xf, yf = xs.frames().float().contiguous(), ys.frames().float().contiguous()
    out_f, loss_f, saved_f = backbone_forward_raw(spec, xf, flat, yf, scale, True)
    gx_f, gw_f = backbone_backward_raw(spec, xf, flat, saved_f, True, True, out=out_f, target=yf, gscale=2 * scale)
    torch.cuda.synchronize()
    n = sum(p.numel() for p in bb.parameters())      # the flat buffers are padded to a multiple of 4 floats
    assert torch.equal(out_i, out_f) and torch.equal(gx_i, gx_f) and torch.equal(gw_i[:n], gw_f[:n])
    assert abs(float(loss_i) - float(loss_f)) <= 1e-12 * abs(float(loss_f))      # double atomics: the order of the per-CTA sums is free


@pytest.mark.parametrize("kind,H,L", [("gru", 48, 1), ("dgru", 20, 2), ("lstm", 33, 2)])
def test_wide_fused_train_step_equals_stock_loop(kind, H, L):
    """NativeTrainStep (graph-replayed fwd / bwd / clip / AdamW) on a layered backbone == autograd + clip_grad_norm_ + torch AdamW."""
    import copy
    from opendpd_b200 import models
    from opendpd_b200.train import NativeTrainStep
    from tests.test_gpu_train import _data, _stock_steps, _flat
    torch.manual_seed(0)
    a = models.CoreModel(2, H, L, kind).cuda()
    b = copy.deepcopy(a)
    batches = [_data(6, 70, s) for s in range(4)]
    la = _stock_steps(a, batches)
    tr = NativeTrainStep(b, lr=5e-4, grad_clip_val=200.0)
    lb = [float(tr.step(x, y).item()) for x, y in batches]
    assert np.allclose(la, lb, rtol=2e-6, atol=0)
    pa, pb = _flat(a), _flat(b)
    assert np.abs(pa - pb).max() <= 2e-6 * max(1.0, np.abs(pa).max())
