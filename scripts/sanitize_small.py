"""Tiny chunked fwd+bwd for compute-sanitizer (memcheck / racecheck): python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendpd_b200 import models
from opendpd_b200.functional import CellSpec, backbone_forward_raw, backbone_backward_raw, chunk_reruns

torch.manual_seed(0)
for kind, H, B, T, tch, tw in (("dgru", 13, 3, 512, (4, 2), 64), ("gru", 32, 2, 300, (3, 3), 32), ("dgru", 13, 2, 256, (8, 8), 32)):
    net = models.CoreModel(2, H, 1, kind).cuda()
    bb = net.backbone
    flat, _ = bb._flat_sync()
    x = (0.2 * torch.randn(B, T, 2)).cuda(); y = (0.8 * x).contiguous()
    spec = CellSpec(bb.cell, H, tchunks=tch, twarm=tw)
    fb, bbuf = {}, {}
    out, loss, saved = backbone_forward_raw(spec, x, flat, y, 1.0 / x.numel(), True, None, fb)
    gx, g = backbone_backward_raw(spec, x, flat, saved, True, True, out=out, target=y, gscale=2.0 / x.numel(), bufs=bbuf)
    torch.cuda.synchronize()
    print(kind, spec.chunk_plan(B, T, False), spec.chunk_plan(B, T, True), float(loss.item()), float(g.abs().sum()),
          chunk_reruns(spec, saved, B, T, False), chunk_reruns(spec, bbuf["ws"], B, T, True))
print("sanitize ok")
