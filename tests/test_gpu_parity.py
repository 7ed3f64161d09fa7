"""GPU parity tests proper: the CUDA path (through the C ABI, via the drop-in backbones) against
(1) the committed golden vectors produced by the unmodified reference and (2) the CPU oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from tests.util import golden_cases, load_golden, rel_err, tol_for, assert_close

pytestmark = pytest.mark.gpu

IMPLEMENTED = ("gru", "dgru", "qgru", "lstm", "deltagru", "tres", "pgjanet", "dvrjanet", "gmp")


def _native_kinds():
    from opendpd_b200 import backbones as bb
    have = set()
    for k, cls in (("gru", "GRU"), ("dgru", "DGRU"), ("qgru", "QGRU"), ("lstm", "LSTM"), ("deltagru", "DeltaGRU"),
                   ("tres", "TResDeltaGRU"), ("pgjanet", "PGJANET"), ("dvrjanet", "DVRJANET"), ("gmp", "GMP")):
        if hasattr(bb, cls):
            have.add(k)
    return have


def build_native(g, device="cuda"):
    from opendpd_b200 import models
    net = models.CoreModel(2, max(g["H"], 1), 1, g["kind"], num_dvr_units=g["K"], thx=g["thx"], thh=g["thh"])
    sd_names = [n for n, _ in net.backbone.named_parameters()]
    assert sd_names == [n for n, _ in g["param_index"]], "parameter names/order differ from the reference"
    off = 0
    with torch.no_grad():
        for (_, p), (_, shape) in zip(net.backbone.named_parameters(), g["param_index"]):
            n = int(np.prod(shape))
            assert list(p.shape) == shape
            p.copy_(torch.from_numpy(g["params"][off:off + n]).view(shape))
            off += n
    return net.to(device)


def grads_flat(net):
    return np.concatenate([p.grad.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])


def _cases():
    return [c for c in golden_cases() if c.split("_")[0] in IMPLEMENTED]


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("name", _cases())
def test_golden_parity(name, fused):
    g = load_golden(name)
    if name.split("_")[0] not in _native_kinds():
        pytest.skip("backbone not built yet")
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
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < tol_for(g, "out")
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert rel_err(x.grad.cpu().numpy(), g["gx"]) < tol_for(g, "gx")
    assert rel_err(grads_flat(net), g["gparams"]) < tol_for(g, "gparams")
    if "mask_x" in g and hasattr(net.backbone, "last_masks"):
        mx, mh = net.backbone.last_masks()
        assert np.array_equal(mx, g["mask_x"])       # delta-x keep mask: bit exact
        assert int((mh != g["mask_h"]).sum()) == 0
        st = net.backbone.raw_statistics()
        assert st == [int(v) for v in g["stats"]]


@pytest.mark.parametrize("kind,H,B,T", [("dgru", 13, 64, 2048), ("gru", 32, 8, 1024), ("dgru", 13, 5, 100), ("gru", 16, 33, 64),
                                        ("qgru", 10, 16, 50), ("qgru_amp1", 10, 16, 50)])
def test_oracle_parity_seeded(kind, H, B, T):
    """Same seeded inputs through the CUDA path and the CPU oracle (fp32 and fp64 arbiter)."""
    from oracle import oracle
    from opendpd_b200 import models
    torch.manual_seed(1234)
    net = models.CoreModel(2, H, 1, kind).cuda()
    gen = torch.Generator().manual_seed(7)
    xc = (0.2 * torch.randn(B, T, 2, generator=gen)).clamp(-0.7, 0.7)
    amp2 = (xc ** 2).sum(-1, keepdim=True)
    yc = xc * (1 - 0.2 * amp2)
    x = xc.cuda().requires_grad_(True)
    out, loss = net.forward_mse(x, yc.cuda())
    loss.backward()
    torch.cuda.synchronize()
    params = np.concatenate([p.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])
    r64 = oracle.run(kind, xc.numpy(), params, target=yc.numpy(), H=H, dtype=np.float64, nthreads=8)
    r32 = oracle.run(kind, xc.numpy(), params, target=yc.numpy(), H=H, dtype=np.float32, nthreads=8)
    for key, mine in (("out", out.detach().cpu().numpy()), ("gx", x.grad.cpu().numpy()), ("gparams", grads_flat(net))):
        tol = max(1e-5, 20 * rel_err(r32[key], r64[key]))
        assert rel_err(mine, r64[key]) < tol, key
    assert abs(loss.item() - r64["loss"]) <= 1e-5 * abs(r64["loss"])


def test_dx_only_matches_full_backward():
    """Frozen-PA mode (models.py:169-171): backward with weights frozen must give the same dX."""
    from opendpd_b200 import models
    torch.manual_seed(3)
    net = models.CoreModel(2, 13, 1, "dgru").cuda()
    x = (0.3 * torch.randn(6, 70, 2)).cuda()
    y = torch.randn(6, 70, 2).cuda()
    xa = x.clone().requires_grad_(True)
    _, la = net.forward_mse(xa, y)
    la.backward()
    for p in net.parameters():
        p.requires_grad = False
    xb = x.clone().requires_grad_(True)
    _, lb = net.forward_mse(xb, y)
    lb.backward()
    assert torch.equal(xa.grad, xb.grad)
    assert all(p.grad is not None for p in net.parameters())  # from the first pass only


def test_bitwise_reproducible():
    from opendpd_b200 import models
    torch.manual_seed(5)
    net = models.CoreModel(2, 13, 1, "dgru").cuda()
    x = (0.3 * torch.randn(32, 96, 2)).cuda()
    y = torch.randn(32, 96, 2).cuda()
    res = []
    for _ in range(2):
        net.zero_grad()
        xa = x.clone().requires_grad_(True)
        out, l = net.forward_mse(xa, y)
        l.backward()
        res.append((out.clone(), xa.grad.clone(), torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()))
    assert all(torch.equal(a, b) for a, b in zip(*res))


def test_cpu_tensor_fails_loudly():
    from opendpd_b200 import models, _ffi
    net = models.CoreModel(2, 8, 1, "gru")
    with pytest.raises(_ffi.OdpdError):
        net(torch.zeros(1, 4, 2))
