"""Tiny fwd+bwd of EVERY cell, serial and time-chunked, for compute-sanitizer:
    compute-sanitizer --tool memcheck  --launch-timeout 900 python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck --launch-timeout 900 python scripts/sanitize_small.py
(--launch-timeout: the first `import torch` on a fresh box takes about a minute; the default made round 1's run give up before the
process had created its CUDA context.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from opendpd_b200 import models
from opendpd_b200.functional import CellSpec, backbone_forward_raw, backbone_backward_raw, chunk_reruns

only = set(sys.argv[1:])
torch.manual_seed(0)
CASES = [  # kind, H, B, T, tchunks (fwd,bwd), twarm
    ("dgru", 13, 3, 160, (1, 1), 0), ("dgru", 13, 3, 320, (4, 2), 64), ("gru", 32, 2, 200, (3, 3), 32), ("qgru", 10, 2, 100, (1, 1), 0),
    ("lstm", 9, 2, 130, (1, 1), 0), ("lstm", 9, 2, 256, (2, 2), 64),
    ("deltagru", 15, 2, 100, (1, 1), 0), ("deltagru_tcnskip", 15, 3, 100, (1, 1), 0),
    ("pgjanet", 15, 2, 100, (1, 1), 0), ("pgjanet", 15, 2, 256, (2, 2), 64), ("dvrjanet", 15, 2, 100, (1, 1), 0), ("dvrjanet", 15, 2, 256, (2, 2), 64),
    ("gmp", 1, 2, 60, (1, 1), 0), ("qgru_qat", 10, 2, 40, (1, 1), 0), ("qgru_qat", 20, 5, 70, (1, 1), 0),
    ("vdlstm", 9, 2, 100, (1, 1), 0), ("vdlstm", 9, 2, 256, (2, 2), 64), ("dgru", 23, 2, 130, (1, 1), 0), ("gru", 8, 2, 70, (1, 1), 0),
    ("dgru", 23, 2, 256, (2, 2), 64), ("gru", 32, 2, 100, (1, 1), 0),                      # lane-per-timestep forward helpers (tiers above 16)
    ("gru", 40, 2, 70, (1, 1), 0, 1), ("dgru", 12, 2, 70, (1, 1), 0, 2), ("lstm", 36, 2, 45, (1, 1), 0, 2),   # layered path (7th field = num_layers)
    ("rvtdcnn", 6, 2, 100, (1, 1), 0), ("bojanet", 10, 2, 100, (1, 1), 0), ("tcnn", 8, 2, 150, (1, 1), 0), ("neuraltx", 8, 2, 150, (1, 1), 0), ("apnrru", 8, 2, 100, (1, 1), 0), ("mcldnn", 8, 2, 100, (1, 1), 0), ("deltajanet", 10, 2, 100, (1, 1), 0), ("tres_qat", 15, 3, 70, (1, 1), 0),
]
for case in CASES:
    kind, H, B, T, tch, tw = case[:6]
    layers = case[6] if len(case) > 6 else 1
    if only and kind not in only:
        continue
    if kind == "tres_qat":
        from opendpd_b200.quant import get_quant_model

        class _Proj:
            quant, n_bits_w, n_bits_a, pretrained_model = True, 16, 16, ""
        net = get_quant_model(_Proj(), models.CoreModel(2, H, 1, "deltagru_tcnskip", thx=0.01, thh=0.05)).cuda().train()
    elif kind == "qgru_qat":
        from opendpd_b200.quant import get_quant_model

        class _Proj:
            quant, n_bits_w, n_bits_a, pretrained_model = True, 8, 8, ""
        net = get_quant_model(_Proj(), models.CoreModel(2, H, 1, "qgru")).cuda().train()
    else:
        net = models.CoreModel(2, H, layers, kind, num_dvr_units=3, thx=0.01, thh=0.05).cuda()
    bb = net.backbone
    flat, _ = bb._flat_sync()
    x = (0.2 * torch.randn(B, T, 2)).cuda(); y = (0.8 * x).contiguous()
    spec = bb._spec()
    spec.tchunks, spec.twarm = tch, (tw, tw)
    fb, bbuf = {}, {}
    stats = bb._stats_tensor(x.device)
    out, loss, saved = backbone_forward_raw(spec, x, flat, y, 1.0 / x.numel(), True, stats, fb)
    gx, g = backbone_backward_raw(spec, x, flat, saved, True, True, out=out, target=y, gscale=2.0 / x.numel(), bufs=bbuf)
    torch.cuda.synchronize()
    print(kind, H, (B, T), "plan", spec.chunk_plan(B, T, False), spec.chunk_plan(B, T, True), "loss", float(loss.item()), "|g|", float(g[:bb.flat_layout()[1]].abs().sum()),      # (the flat buffer is padded to 4 floats: the tail is never written)
          "|gx|", float(gx.abs().sum()), flush=True)
print("sanitize ok")
