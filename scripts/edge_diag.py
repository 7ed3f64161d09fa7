"""Diagnostic: where do the short-frame gradient differences sit?  (element indices of the largest |native - fp64 oracle|)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle
from opendpd_b200 import models


def run(kind, H, T, B=3, seed=99, xseed=5, skip=0):
    torch.manual_seed(seed)
    thx, thh = (0.01, 0.05) if kind in ("deltagru", "deltagru_tcnskip") else (0.0, 0.0)
    net = models.CoreModel(2, H, 1, kind, num_dvr_units=3, thx=thx, thh=thh).cuda()
    params = np.concatenate([p.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])
    gen = torch.Generator().manual_seed(xseed)
    for _ in range(skip):
        torch.randn(3, 7, 2, generator=gen)
    xc = (0.2 * torch.randn(B, T, 2, generator=gen)).clamp(-0.7, 0.7)
    yc = xc * (1 - 0.2 * (xc ** 2).sum(-1, keepdim=True))
    x = xc.cuda().requires_grad_(True)
    out, loss = net.forward_mse(x, yc.cuda())
    loss.backward()
    torch.cuda.synchronize()
    r64 = oracle.run(kind, xc.numpy(), params, target=yc.numpy(), H=H, thx=thx, thh=thh, dtype=np.float64, nthreads=1)
    r32 = oracle.run(kind, xc.numpy(), params, target=yc.numpy(), H=H, thx=thx, thh=thh, dtype=np.float32, nthreads=1)
    gx = x.grad.cpu().numpy().astype(np.float64)
    gp = np.concatenate([p.grad.cpu().numpy().ravel() for _, p in net.backbone.named_parameters()]).astype(np.float64)
    o = out.detach().cpu().numpy().astype(np.float64)
    res = []
    for key, mine in (("out", o), ("gx", gx), ("gparams", gp)):
        ref = r64[key]
        e = np.abs(mine - ref) / (np.abs(ref).max() + 1e-300)
        e32 = np.abs(r32[key].astype(np.float64) - ref) / (np.abs(ref).max() + 1e-300)
        top = np.argsort(e.ravel())[::-1][:4]
        idx = [tuple(int(v) for v in np.unravel_index(i, e.shape)) for i in top]
        res.append(f"{key}: max {e.max():.2e} (oracle fp32 {e32.max():.2e}) at {idx} vals {[float(f'{e.ravel()[i]:.1e}') for i in top]}")
    print(f"{kind} H={H} T={T} B={B}: " + " | ".join(res), flush=True)


for T in (15, 16, 17, 31, 33, 40, 64, 65, 100):
    run("dvrjanet", 11, T)
for H in (8, 10, 11, 12, 13, 15):
    run("dvrjanet", H, 15)
for T in (32, 33, 34, 40, 63, 65, 97):
    run("deltagru", 15, T)
for H in (12, 13, 15, 16):
    run("deltagru", H, 33)
run("deltagru_tcnskip", 15, 33)
run("pgjanet", 13, 15)
run("pgjanet", 11, 15)
