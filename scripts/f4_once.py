"""One forward + backward (dW + dX) of every row f-4 cell and two layered-path configurations at B=64 x T=2048 (synthetic IQ), for an
ncu launch list:   ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none --csv python scripts/f4_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendpd_b200 import models
from opendpd_b200.functional import backbone_forward_raw, backbone_backward_raw

torch.manual_seed(0)
B, T = 64, 2048
x = (0.2 * torch.randn(B, T, 2)).clamp(-0.7, 0.7).cuda(); y = (0.8 * x).contiguous()
for kind, H, L in (("vdlstm", 9, 1), ("bojanet", 10, 1), ("apnrru", 8, 1), ("deltajanet", 10, 1), ("mcldnn", 8, 1), ("rvtdcnn", 6, 1), ("tcnn", 8, 1),
                   ("neuraltx", 8, 1), ("dgru", 64, 1), ("gru", 32, 2)):
    net = models.CoreModel(2, H, L, kind).cuda()
    bb = net.backbone
    flat, _ = bb._flat_sync()
    spec = bb._spec()
    for _ in range(2):      # the second pass is the one to read (attributes set, caches warm)
        out, loss, saved = backbone_forward_raw(spec, x, flat, y, 1.0 / x.numel(), True, bb._stats_tensor(x.device))
        gx, g = backbone_backward_raw(spec, x, flat, saved, True, True, out=out, target=y, gscale=2.0 / x.numel())
    torch.cuda.synchronize()
    print(kind, H, L, float(loss), flush=True)
