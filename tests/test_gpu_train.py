"""GPU tests of the train-step surface: fused NativeTrainStep == stock (autograd + clip_grad_norm_ + torch AdamW) loop,
cascaded DPD->frozen-PA step (steps/train_dpd.py:60-63), and the net_train drop-in."""
import copy
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _data(B, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (0.25 * torch.randn(B, T, 2, generator=g)).clamp(-0.8, 0.8)
    y = x * (1 - 0.2 * (x ** 2).sum(-1, keepdim=True))
    return x.cuda(), y.cuda()


def _stock_steps(net, batches, lr=5e-4, clip=200.0):
    opt = torch.optim.AdamW([p for p in net.parameters() if p.requires_grad], lr=lr)
    losses = []
    for x, y in batches:
        opt.zero_grad()
        loss = nn.MSELoss()(net(x), y)
        loss.backward()
        nn.utils.clip_grad_norm_(net.parameters(), clip)
        opt.step()
        losses.append(loss.item())
    return losses


def _flat(net):
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu().numpy()


@pytest.mark.parametrize("kind,H", [("dgru", 13), ("gru", 32), ("deltagru_tcnskip", 15), ("pgjanet", 10), ("gmp", 1), ("lstm", 9),
                                    ("dvrjanet", 10), ("qgru_qat", 10), ("vdlstm", 9), ("rvtdcnn", 6), ("bojanet", 10), ("tcnn", 8), ("neuraltx", 8), ("apnrru", 8), ("mcldnn", 8), ("deltajanet", 10), ("tres_qat", 15)])
def test_fused_step_equals_stock_loop(kind, H):
    from opendpd_b200 import models
    from opendpd_b200.train import NativeTrainStep
    torch.manual_seed(0)
    if kind in ("qgru_qat", "tres_qat"):
        from opendpd_b200.quant import get_quant_model

        class _Proj:
            quant, n_bits_w, n_bits_a, pretrained_model = True, 16, 16, ""
        base = models.CoreModel(2, H, 1, "qgru") if kind == "qgru_qat" else models.CoreModel(2, H, 1, "deltagru_tcnskip", thx=0.01, thh=0.05)
        a = get_quant_model(_Proj(), base).cuda().train()
    else:
        a = models.CoreModel(2, H, 1, kind, num_dvr_units=3, thx=0.01, thh=0.05).cuda()
    b = copy.deepcopy(a)
    batches = [_data(6, 70, s) for s in range(4)]
    la = _stock_steps(a, batches)
    tr = NativeTrainStep(b, lr=5e-4, grad_clip_val=200.0)
    lb = [float(tr.step(x, y).item()) for x, y in batches]
    assert np.allclose(la, lb, rtol=2e-6, atol=0)
    pa, pb = _flat(a), _flat(b)
    assert np.abs(pa - pb).max() <= 2e-6 * max(1.0, np.abs(pa).max())


def test_clipping_path_matches_torch():
    """max_norm small enough to actually clip."""
    from opendpd_b200 import models
    from opendpd_b200.train import NativeTrainStep
    torch.manual_seed(1)
    a = models.CoreModel(2, 13, 1, "dgru").cuda()
    b = copy.deepcopy(a)
    batches = [_data(4, 50, s) for s in range(3)]
    la = _stock_steps(a, batches, clip=1e-3)
    tr = NativeTrainStep(b, lr=5e-4, grad_clip_val=1e-3)
    lb = [float(tr.step(x, y).item()) for x, y in batches]
    assert np.allclose(la, lb, rtol=2e-6)
    assert np.abs(_flat(a) - _flat(b)).max() < 2e-6


def test_cascaded_step_dpd_into_frozen_pa():
    from opendpd_b200 import models
    from opendpd_b200.train import NativeTrainStep
    from oracle import torch_port
    torch.manual_seed(2)
    dpd = models.CoreModel(2, 13, 1, "dgru").cuda()
    torch.manual_seed(3)
    pa = models.CoreModel(2, 13, 1, "dgru").cuda()
    cas = models.CascadedModel(dpd, pa)
    cas.freeze_pa_model()
    x, y = _data(5, 64, 7)
    # (1) autograd path through two native backbones vs the PyTorch-op restatement on CPU
    cas.zero_grad()
    loss = nn.MSELoss()(cas(x), y)
    loss.backward()
    fd = torch.cat([p.detach().reshape(-1) for p in dpd.backbone.parameters()]).cpu().requires_grad_(True)
    fp = torch.cat([p.detach().reshape(-1) for p in pa.backbone.parameters()]).cpu()
    mid = torch_port.forward("dgru", x.cpu(), fd, 13)
    ref_loss = nn.MSELoss()(torch_port.forward("dgru", mid, fp, 13), y.cpu())
    ref_loss.backward()
    g = torch.cat([p.grad.reshape(-1) for p in dpd.backbone.parameters()]).cpu().numpy()
    assert abs(loss.item() - ref_loss.item()) < 1e-5 * ref_loss.item()
    assert np.abs(g - fd.grad.numpy()).max() < 1e-5 * np.abs(fd.grad.numpy()).max()
    assert all(p.grad is None for p in pa.parameters())
    # (2) fused cascaded train step == stock loop
    cas2 = copy.deepcopy(cas)
    batches = [_data(5, 64, s) for s in range(3)]
    la = _stock_steps(cas, batches)
    tr = NativeTrainStep(cas2)
    lb = [float(tr.step(xb, yb).item()) for xb, yb in batches]
    assert np.allclose(la, lb, rtol=2e-6)
    assert np.abs(_flat(cas.dpd_model) - _flat(cas2.dpd_model)).max() < 2e-6
    assert np.array_equal(_flat(cas.pa_model), _flat(cas2.pa_model))


def test_net_train_dropin_and_host_step():
    from opendpd_b200 import models
    from opendpd_b200.train import net_train, NativeTrainStep
    torch.manual_seed(4)
    net = models.CoreModel(2, 13, 1, "dgru").cuda()
    xs, ys = zip(*[(x.cpu(), y.cpu()) for x, y in [_data(4, 40, s) for s in range(3)]])
    loader = list(zip(xs, ys))
    opt = torch.optim.AdamW(net.parameters(), lr=5e-4)
    log = {}
    net_train(log, net, loader, opt, nn.MSELoss(), 200.0, torch.device("cuda"))
    assert np.isfinite(log["loss"])
    tr = NativeTrainStep(net)
    v = tr.step_host(xs[0].pin_memory(), ys[0].pin_memory())
    assert np.isfinite(v)


def test_state_dict_roundtrip_after_training():
    from opendpd_b200 import models
    from opendpd_b200.train import NativeTrainStep
    torch.manual_seed(5)
    net = models.CoreModel(2, 13, 1, "dgru").cuda()
    tr = NativeTrainStep(net)
    x, y = _data(4, 40)
    tr.step(x, y)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    assert list(sd) == ["backbone.rnn.weight_ih_l0", "backbone.rnn.weight_hh_l0", "backbone.rnn.bias_ih_l0", "backbone.rnn.bias_hh_l0",
                        "backbone.fc_out.weight", "backbone.fc_out.bias", "backbone.fc_hid.weight", "backbone.fc_hid.bias"]
    net2 = models.CoreModel(2, 13, 1, "dgru")
    net2.load_state_dict(sd)
    net2 = net2.cuda()
    assert torch.equal(net2(x), net(x))


@pytest.mark.parametrize("kind,H,B,T", [("dgru", 13, 2, 19662), ("deltagru_tcnskip", 15, 1, 19662), ("gru", 32, 3, 2560), ("lstm", 9, 1, 19662),
                                        ("pgjanet", 15, 1, 19662), ("rvtdcnn", 6, 2, 19662), ("bojanet", 10, 1, 19662), ("tcnn", 8, 2, 19662),
                                        ("neuraltx", 8, 1, 19662), ("apnrru", 8, 2, 19662), ("mcldnn", 8, 1, 19662), ("deltajanet", 10, 1, 19662)])
def test_eval_path_whole_segments(kind, H, B, T):
    """net_eval / run_dpd run forward-only over whole segments (T = nperseg up to 19 662, B = 1..6)."""
    from oracle import oracle
    from opendpd_b200 import models
    from opendpd_b200.train import net_eval
    from tests.util import assert_close
    torch.manual_seed(6)
    net = models.CoreModel(2, H, 1, kind, thx=0.0, thh=0.0).cuda()
    g = torch.Generator().manual_seed(8)
    x = (0.25 * torch.randn(B, T, 2, generator=g)).clamp(-0.8, 0.8)
    y = 0.9 * x
    log = {}
    _, pred, gt = net_eval(log, net, [(x, y)], nn.MSELoss(), torch.device("cuda"))
    params = np.concatenate([p.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])
    r64 = oracle.run(kind, x.numpy(), params, target=y.numpy(), H=H, dtype=np.float64, want_grads=False, nthreads=4)
    r32 = oracle.run(kind, x.numpy(), params, target=y.numpy(), H=H, dtype=np.float32, want_grads=False, nthreads=4)
    tol = max(1e-5, 3 * float(np.quantile(np.abs(r32["out"] - r64["out"]) / np.abs(r64["out"]).max(), 0.9999)))
    assert_close(pred, r64["out"], tol, "eval out")
    assert abs(log["loss"] - r64["loss"]) <= 1e-5 * r64["loss"] and np.array_equal(gt, y.numpy())


def test_pipelined_host_loop_equals_sequential():
    from opendpd_b200 import models
    from opendpd_b200.train import NativeTrainStep
    import copy
    torch.manual_seed(9)
    a = models.CoreModel(2, 13, 1, "dgru").cuda()
    b = copy.deepcopy(a)
    batches = [tuple(t.cpu().pin_memory() for t in _data(4, 64, s)) for s in range(7)]
    ta, tb = NativeTrainStep(a), NativeTrainStep(b)
    la = [ta.step_host(x, y) for x, y in batches]
    lb = tb.run_host_batches(iter(batches))
    assert la == lb
    assert np.array_equal(_flat(a), _flat(b))


@pytest.mark.parametrize("kind,H,pa", [("dgru", 13, None), ("deltagru_tcnskip", 15, ("dgru", 8))])
def test_multi_step_graph_replay_equals_single_steps(kind, H, pa):
    """NativeTrainStep.steps_indexed (K steps per CUDA-graph replay) must be the same training run as K step_indexed calls: same
    losses, bit-identical parameters — also across the eager / capture / replay phases of the graph cache."""
    from opendpd_b200 import models
    from opendpd_b200.train import NativeTrainStep
    torch.manual_seed(4)

    def build():
        torch.manual_seed(4)
        net = models.CoreModel(2, H, 1, kind, thx=0.01, thh=0.05).cuda()
        if pa:
            torch.manual_seed(5)
            net = models.CascadedModel(net, models.CoreModel(2, pa[1], 1, pa[0]).cuda())
            net.freeze_pa_model()
        return net
    a, b = build(), build()
    g = torch.Generator().manual_seed(9)
    N, T, B, K = 3000, 128, 6, 4
    sx = (0.25 * torch.randn(N, 2, generator=g)).cuda()
    sy = (0.9 * sx).contiguous()
    starts = torch.randint(0, N - T, (5 * K, B), generator=g).to(torch.int32)
    ta, tb = NativeTrainStep(a), NativeTrainStep(b)
    la = [float(ta.step_indexed(sx, sy, starts[i].cuda(), T).item()) for i in range(5 * K)]
    lb = []
    for r in range(5):                       # replay 0 eager, 1 captures, 2.. replay; pinned-host starts on odd rounds
        rows = starts[r * K:(r + 1) * K]
        rows = rows.pin_memory() if r & 1 else rows.cuda()
        lb += [float(v) for v in tb.steps_indexed(sx, sy, rows, T).cpu()]
    assert la == lb, (la, lb)
    pa_, pb_ = _flat(a), _flat(b)
    assert np.array_equal(pa_, pb_)
