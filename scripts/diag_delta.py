import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle
from opendpd_b200 import models
def q(e): return f"max {e.max():.2e} p99.99 {np.quantile(e,0.9999):.2e} p99 {np.quantile(e,0.99):.2e} med {np.median(e):.2e}"
for kind,H,B,T,thx,thh in [("deltagru_tcnskip",15,64,2048,0.01,0.05),("deltagru_tcnskip",15,64,2048,0.0,0.0),("dgru",13,64,2048,0,0)]:
    torch.manual_seed(4321 if "delta" in kind else 1234)
    net = models.CoreModel(2, H, 1, kind, thx=thx, thh=thh).cuda()
    if "delta" in kind: net.backbone.keep_masks=True
    gen = torch.Generator().manual_seed(11 if "delta" in kind else 7)
    xc = (0.2 * torch.randn(B, T, 2, generator=gen)).clamp(-0.7, 0.7)
    yc = xc * (1 - 0.2 * (xc ** 2).sum(-1, keepdim=True))
    x = xc.cuda().requires_grad_(True)
    out, loss = net.forward_mse(x, yc.cuda()); loss.backward(); torch.cuda.synchronize()
    params = np.concatenate([p.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])
    kw=dict(H=H, thx=thx, thh=thh, nthreads=8, want_masks=True)
    r64 = oracle.run(kind, xc.numpy(), params, target=yc.numpy(), dtype=np.float64, **kw)
    r32 = oracle.run(kind, xc.numpy(), params, target=yc.numpy(), dtype=np.float32, **kw)
    good=np.arange(B)
    if "delta" in kind:
        mx,mh=net.backbone.last_masks()
        f1=np.unique(np.nonzero(mh!=r64["mask_h"])[0]); f2=np.unique(np.nonzero(r32["mask_h"]!=r64["mask_h"])[0])
        print(kind,thx,thh,"flipped seqs mine-vs-64",len(f1),"oracle32-vs-64",len(f2))
        good=np.setdiff1d(good,np.union1d(f1,f2))
    for key,mine in (("out",out.detach().cpu().numpy()),("gx",x.grad.cpu().numpy())):
        sc=np.abs(r64[key][good]).max()
        print(" ",key,"mine-64:",q(np.abs(mine[good]-r64[key][good])/sc),"| o32-64:",q(np.abs(r32[key][good].astype(np.float64)-r64[key][good])/sc))
    g=np.concatenate([p.grad.detach().cpu().numpy().ravel() for _, p in net.backbone.named_parameters()])
    sc=np.abs(r64["gparams"]).max()
    print("  gparams mine-64:",q(np.abs(g-r64["gparams"])/sc),"| o32-64:",q(np.abs(r32["gparams"].astype(np.float64)-r64["gparams"])/sc), "loss",loss.item(),r64["loss"])
