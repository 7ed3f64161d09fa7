"""Diagnostic (not a test): detailed error statistics of the CUDA path vs the fp64/fp32 CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle
from opendpd_b200 import models

def stats(name, a, b64, b32):
    a = a.astype(np.float64); sc = np.abs(b64).max()
    e = np.abs(a - b64) / sc; e32 = np.abs(b32.astype(np.float64) - b64) / sc
    i = np.unravel_index(np.argmax(e), e.shape)
    print(f"  {name:8s} mine: max {e.max():.2e} p99.9 {np.quantile(e,0.999):.2e} med {np.median(e):.2e} | oracle32: max {e32.max():.2e} p99.9 {np.quantile(e32,0.999):.2e} med {np.median(e32):.2e} | argmax {i}")
    return i

for kind, H, B, T in [("dgru", 13, 64, 2048), ("gru", 32, 8, 1024)]:
    torch.manual_seed(1234)
    net = models.CoreModel(2, H, 1, kind).cuda()
    gen = torch.Generator().manual_seed(7)
    xc = (0.2 * torch.randn(B, T, 2, generator=gen)).clamp(-0.7, 0.7)
    yc = xc * (1 - 0.2 * (xc ** 2).sum(-1, keepdim=True))
    x = xc.cuda().requires_grad_(True)
    out, loss = net.forward_mse(x, yc.cuda()); loss.backward(); torch.cuda.synchronize()
    params = np.concatenate([p.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])
    r64 = oracle.run(kind, xc.numpy(), params, target=yc.numpy(), H=H, dtype=np.float64, nthreads=8)
    r32 = oracle.run(kind, xc.numpy(), params, target=yc.numpy(), H=H, dtype=np.float32, nthreads=8)
    print(kind, H, B, T, "loss", loss.item(), r64["loss"])
    stats("out", out.detach().cpu().numpy(), r64["out"], r32["out"])
    i = stats("gx", x.grad.cpu().numpy(), r64["gx"], r32["gx"])
    amp = np.sqrt((xc.numpy() ** 2).sum(-1))
    print("   amp at argmax", amp[i[0], i[1]], "min amp", amp.min())
    g = np.concatenate([p.grad.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])
    stats("gparams", g, r64["gparams"], r32["gparams"])
