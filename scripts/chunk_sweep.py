"""GPU tuning helper: time the forward / backward launches of one backbone for a sweep of time-chunk plans.
    python scripts/chunk_sweep.py [kind H B T]            (default: dgru 13 64 2048 = BASELINE configs[1])"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import synth_batches
from opendpd_b200 import models
from opendpd_b200.functional import CellSpec, backbone_forward_raw, backbone_backward_raw, chunk_reruns, chunk_worst_mismatch


def main():
    kind, H, B, T = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else ("dgru", 13, 64, 2048)
    torch.manual_seed(0)
    net = models.CoreModel(2, H, 1, kind, num_dvr_units=3).cuda()
    bb = net.backbone
    flat, _ = bb._flat_sync()
    POOL = 40
    xs, ys = synth_batches(POOL, B, T, 1000)
    xd, yd = xs.cuda(), ys.cuda()
    count = float(2 * B * T)
    plans = [((1, 1), 0), ((4, 4), 128), ((8, 4), 128), ((8, 5), 128), ((8, 6), 128), ((8, 8), 128), ((16, 8), 128), ((8, 6), 64), ((8, 6), 96),
             ((16, 16), 64), ((0, 0), 0)]
    if os.environ.get("SWEEP_PLANS"):
        plans = [tuple(p) for p in json.loads(os.environ["SWEEP_PLANS"])]
        plans = [((p[0], p[1]), p[2]) for p in plans]
    for tch, tw in plans:
        spec = bb._spec()
        spec.tchunks, spec.twarm = tuple(tch), tw
        fb, bbuf = {}, {}
        n = 60
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
        gflat = torch.empty_like(flat)
        for i in range(-5, n):
            xb, yb = xd[(i * 7 + 3) % POOL], yd[(i * 7 + 3) % POOL]
            e = ev[max(i, 0)]
            e[0].record()
            out, loss, saved = backbone_forward_raw(spec, xb, flat, yb, 1.0 / count, True, None, fb)
            e[1].record()
            backbone_backward_raw(spec, xb, flat, saved, False, True, out=out, target=yb, gscale=2.0 / count, gflat=gflat, bufs=bbuf)
            e[2].record()
        torch.cuda.synchronize()
        f = float(np.median([e[0].elapsed_time(e[1]) for e in ev])); b = float(np.median([e[1].elapsed_time(e[2]) for e in ev]))
        print(json.dumps({"tchunks": tch, "twarm": tw, "plan_f": spec.chunk_plan(B, T, False), "plan_b": spec.chunk_plan(B, T, True),
                          "fwd_ms": round(f, 4), "bwd_ms": round(b, 4), "reruns_f": chunk_reruns(spec, saved, B, T, False),
                          "reruns_b": chunk_reruns(spec, bbuf["ws"], B, T, True),
                          "worst_f": chunk_worst_mismatch(spec, saved, B, T, False), "worst_b": chunk_worst_mismatch(spec, bbuf["ws"], B, T, True),
                          "loss": float(loss.item())}), flush=True)


if __name__ == "__main__":
    main()
