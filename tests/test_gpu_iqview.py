"""Storage variants of the IQ tensors at the C ABI (include/odpd.h): bf16 pairs instead of fp32 pairs (ODPD_F_X_BF16 /
ODPD_F_TARGET_BF16; BASELINE config 3's "bf16") and on-device framing (OdpdDims.x_starts / .target_starts: frame b = samples
starts[b] .. starts[b]+T-1 of the raw stream, the stride-1 windows of modules/data_collector.py:233-252).  Both are address/width
conversions only, so the results must equal the fp32 / materialised-frame path BIT FOR BIT."""
import copy
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

KINDS = [("gru", 16), ("dgru", 13), ("dgru", 23), ("qgru", 10), ("lstm", 9), ("deltagru", 15), ("deltagru_tcnskip", 15), ("pgjanet", 15),
         ("dvrjanet", 15), ("gmp", 1), ("vdlstm", 9), ("rvtdcnn", 6), ("gru", 40), ("bojanet", 10), ("tcnn", 8), ("neuraltx", 8), ("apnrru", 8), ("mcldnn", 8), ("deltajanet", 10)]


def _net(kind, H):
    from opendpd_b200 import models
    torch.manual_seed(11)
    return models.CoreModel(2, H, 1, kind, num_dvr_units=3, thx=0.01, thh=0.05).cuda()


def _run(net, x, y):
    from opendpd_b200.functional import backbone_forward_raw, backbone_backward_raw
    bb = net.backbone
    flat, _ = bb._flat_sync()
    spec = bb._spec()
    B, T = x.shape[0], x.shape[1]
    count = float(2 * B * T)
    out, loss, saved = backbone_forward_raw(spec, x, flat, y, 1.0 / count, True, bb._stats_tensor(flat.device))
    gx, gflat = backbone_backward_raw(spec, x, flat, saved, True, True, out=out, target=y, gscale=2.0 / count)
    torch.cuda.synchronize()
    return out.clone(), loss.clone(), gx.clone(), gflat[:bb.flat_layout()[1]].clone()      # (the flat buffer is padded to 4 floats)


def _same(a, b):
    """out, gx, gparams bit for bit; the loss is a double accumulated with atomics over CTAs (order-dependent last bits)."""
    for u, v, name in zip(a, b, ("out", "loss", "gx", "gparams")):
        if name == "loss":
            assert abs(u.item() - v.item()) <= 1e-13 * abs(v.item()), name
        else:
            assert torch.equal(u, v), name


def _stream(N, seed):
    g = torch.Generator().manual_seed(seed)
    s = (0.2 * torch.randn(N, 2, generator=g)).clamp(-0.7, 0.7)
    return s, s * (1 - 0.2 * (s ** 2).sum(-1, keepdim=True))


@pytest.mark.parametrize("kind,H", KINDS)
def test_bf16_storage_is_an_exact_widening(kind, H):
    net = _net(kind, H)
    sx, sy = _stream(6 * 320, 3)
    xb, yb = sx.view(6, 320, 2).cuda().bfloat16(), sy.view(6, 320, 2).cuda().bfloat16()
    a = _run(net, xb, yb)
    b = _run(net, xb.float(), yb.float())
    _same(a, b)
    assert a[0].dtype == torch.float32 and a[2].dtype == torch.float32


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("kind,H", KINDS)
def test_frame_starts_equal_materialised_frames(kind, H, bf16):
    from opendpd_b200.functional import IqStream
    net = _net(kind, H)
    sx, sy = _stream(5000, 5)
    sx, sy = sx.cuda(), sy.cuda()
    if bf16:
        sx, sy = sx.bfloat16(), sy.bfloat16()
    T = 288
    starts = torch.tensor([0, 4711, 17, 5000 - T, 1234, 1235, 1233], dtype=torch.int32, device="cuda")   # overlapping windows, both ends
    vx, vy = IqStream(sx, starts, T), IqStream(sy, starts, T)
    a = _run(net, vx, vy)
    b = _run(net, vx.frames().contiguous(), vy.frames().contiguous())
    _same(a, b)


def test_chunked_long_frames_through_frame_starts():
    """The time-chunked kernels address warm-up steps relative to the frame start as well."""
    from opendpd_b200.functional import IqStream
    net = _net("dgru", 13)
    sx, sy = _stream(9000, 9)
    sx, sy = sx.cuda().bfloat16(), sy.cuda().bfloat16()
    T = 2048
    starts = torch.arange(0, 64, dtype=torch.int32, device="cuda") * 97
    vx, vy = IqStream(sx, starts, T), IqStream(sy, starts, T)
    a = _run(net, vx, vy)
    b = _run(net, vx.frames().float().contiguous(), vy.frames().float().contiguous())
    _same(a, b)


def test_trainer_step_indexed_equals_step_on_gathered_frames():
    from opendpd_b200.train import NativeTrainStep
    from opendpd_b200.functional import IqStream
    net = _net("dgru", 13)
    ref = copy.deepcopy(net)
    tr, trr = NativeTrainStep(net), NativeTrainStep(ref)
    sx, sy = _stream(6000, 2)
    sx, sy = sx.cuda(), sy.cuda()
    T = 512
    g = torch.Generator().manual_seed(0)
    for i in range(5):
        starts = torch.randint(0, 6000 - T, (16,), generator=g).to(torch.int32).cuda()
        la = tr.step_indexed(sx, sy, starts, T)
        fx, fy = IqStream(sx, starts, T).frames().contiguous(), IqStream(sy, starts, T).frames().contiguous()
        lb = trr.step(fx, fy)
        assert abs(la.item() - lb.item()) <= 1e-13 * abs(lb.item()), i
    pa = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    pb = torch.cat([p.detach().reshape(-1) for p in ref.parameters()])
    assert torch.equal(pa, pb)
