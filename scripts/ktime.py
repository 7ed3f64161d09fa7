"""GPU tuning helper: device time of the forward / backward launches of ONE backbone call on real frames, measured like bench.py's
kernel_ms (CUDA events around a CUDA-graph replay of NK back-to-back launches; distinct cold batches for the forward).
    python scripts/ktime.py kind H B T [plan ...]        plan = fwdchunks,bwdchunks,warmup   (default: 0,0,0 = library picks; 1,1,0 = serial)
Environment knobs of the library (ODPD_*) apply; prints one JSON line per plan."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from opendpd_b200 import models
from opendpd_b200.functional import backbone_forward_raw, backbone_backward_raw, chunk_reruns


def main():
    kind, H, B, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    plans = [tuple(int(v) for v in p.split(",")) for p in sys.argv[5:]] or [(0, 0, 0)]
    torch.manual_seed(0)
    dev = torch.device("cuda", 0)
    net = models.CoreModel(2, H, int(os.environ.get("KT_LAYERS", "1")), kind, num_dvr_units=3, thx=0.01, thh=0.05).to(dev)   # KT_LAYERS: stacked layers (layered path)
    bb = net.backbone
    flat, _ = bb._flat_sync()
    wl = dict(kind=kind, H=H, B=B, T=T, dataset="APA_200MHz")
    feed = bench.Feed(wl, None, 0, 1, B, B)
    POOL = int(max(4, min(40, (160 << 20) // (2 * B * T * 8) + 1)))
    xs, ys = feed.frames(feed.table(0, POOL))
    xd, yd = xs.to(dev), ys.to(dev)
    count = float(2 * B * T)
    NK = min(POOL, 32)
    for fc, bc, tw in plans:
        spec = bb._spec()
        spec.tchunks, spec.twarm = (fc, bc), (tw, tw)
        gflat = torch.empty_like(flat)
        kb0, kb1, kb2 = {}, {}, {}
        stats = bb._stats_tensor(dev)
        out, loss, saved = backbone_forward_raw(spec, xd[0], flat, yd[0], 1.0 / count, True, stats, kb0)
        backbone_backward_raw(spec, xd[0], flat, saved, True, True, out=out, target=yd[0], gscale=2.0 / count, gflat=gflat, bufs=kb1)
        backbone_backward_raw(spec, xd[0], flat, saved, True, False, out=out, target=yd[0], gscale=2.0 / count, bufs=kb2)
        torch.cuda.synchronize()

        def replay_ms(fn):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                for i in range(NK):
                    fn(i)
            g.replay(); torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); g.replay(); b_.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b_) / NK)
            return float(np.median(ts))
        f = replay_ms(lambda i: backbone_forward_raw(spec, xd[(i * 7 + 3) % POOL], flat, yd[(i * 7 + 3) % POOL], 1.0 / count, True, stats, kb0))
        last = ((NK - 1) * 7 + 3) % POOL
        bdw = replay_ms(lambda i: backbone_backward_raw(spec, xd[last], flat, saved, False, True, out=out, target=yd[last], gscale=2.0 / count, gflat=gflat, bufs=kb1))
        bdx = replay_ms(lambda i: backbone_backward_raw(spec, xd[last], flat, saved, True, False, out=out, target=yd[last], gscale=2.0 / count, bufs=kb2))
        print(json.dumps({"kind": kind, "H": H, "B": B, "T": T, "plan_f": spec.chunk_plan(B, T, False), "plan_b": spec.chunk_plan(B, T, True),
                          "fwd_us": round(f * 1e3, 2), "bwd_dw_us": round(bdw * 1e3, 2), "bwd_dx_us": round(bdx * 1e3, 2),
                          "reruns_f": chunk_reruns(spec, saved, B, T, False), "reruns_b": chunk_reruns(spec, kb1["ws"], B, T, True),
                          "env": {k: v for k, v in os.environ.items() if k.startswith("ODPD_")}}), flush=True)


if __name__ == "__main__":
    main()
