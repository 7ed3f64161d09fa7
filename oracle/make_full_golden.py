#!/usr/bin/env python
"""Full-size golden vectors from the UNMODIFIED reference (test infrastructure; authoring container only).

    python oracle/make_full_golden.py          -> tests/golden/full_c2a.npz, tests/golden/full_c3.npz

The small fixtures of make_golden.py top out at B8.T256; these two pin the kernels at the sizes BASELINE.json quotes, on real frames:
  full_c2a  DGRU H=13 (1041 p), B=64 x T=2048 frames of APA_200MHz (configs[1]), target = measured PA output (train_pa)
  full_c3   TRes-DeltaGRU H=15 (999 p, thx .01, thh .05), B=256 x T=2048 frames of APA_200MHz (configs[2]), target = gain*x (train_dpd)
Frames are IQFrameDataset's stride-1 windows at the first B indices of torch.randperm(n_frames, manual_seed(0)) — the first batch
bench.py times.  Inputs are NOT stored: the tests rebuild them bit for bit from tests/golden/iq_streams.npz + the stored start indices.
Stored: start indices, parameters (reference init, seed 0), fp32 reference out / gx / gparams / loss for the whole batch, the fp64
arbiter for gparams / loss and for the first N64 sequences of out / gx, and for full_c3 the delta keep-masks and sparsity counters."""
import os, sys, json
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import make_golden as mg          # imports the reference (sys.path -> /root/reference) and its helpers

N64 = 32


def main():
    z = np.load(os.path.join(mg.OUT, "iq_streams.npz"))
    torch.set_num_threads(8)
    for name, kind, H, B, T, thx, thh, tkey in (("full_c2a", "dgru", 13, 64, 2048, 0.0, 0.0, "APA_200MHz.y"),
                                                ("full_c3", "deltagru_tcnskip", 15, 256, 2048, 0.01, 0.05, "APA_200MHz.dpd_target")):
        X, Y = z["APA_200MHz.x"], z[tkey]
        n = X.shape[0] - T + 1
        starts = torch.randperm(n, generator=torch.Generator().manual_seed(0))[:B].numpy().astype(np.int32)
        x = np.stack([X[k:k + T] for k in starts]); y = np.stack([Y[k:k + T] for k in starts])
        net = mg.build(kind, H, 0, thx, thh)
        tap_cls = None
        if kind == "deltagru_tcnskip":
            from backbones.deltagru_tcnskip import DeltaGRULayer as tap_cls
        params = mg.flat_params(net)
        r32 = mg.run(net, x, y, torch.float32, tap_cls)
        r64 = mg.run(net, x, y, torch.float64, tap_cls)
        net.float()
        rec = dict(starts=starts, params=params.astype(np.float32), kind=np.array(kind), H=np.array(H), K=np.array(3),
                   thx=np.array(thx, dtype=np.float64), thh=np.array(thh, dtype=np.float64),
                   param_index=np.array(json.dumps(mg.param_index(net))), target_key=np.array(tkey),
                   out=r32["out"], gx=r32["gx"], gparams=r32["gparams"], loss=r32["loss"],
                   out64=r64["out"][:N64], gx64=r64["gx"][:N64], gparams64=r64["gparams"], loss64=r64["loss"])
        if tap_cls is not None:
            rec["mask_x"] = r32["mask_x"].astype(np.uint8); rec["mask_h"] = r32["mask_h"].astype(np.uint16)
            rec["mask_x64"] = r64["mask_x"].astype(np.uint8); rec["mask_h64"] = r64["mask_h"].astype(np.uint16)
            rec["stats"] = r32["stats"]; rec["stats64"] = r64["stats"]
        path = os.path.join(mg.OUT, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, "loss", float(r32["loss"]), float(r64["loss"]), "bytes", os.path.getsize(path), flush=True)


if __name__ == "__main__":
    main()
