#!/usr/bin/env python
"""Pin the C oracle to the UNMODIFIED reference at the edges of every backbone's frame-length domain (test infrastructure only;
authoring container only — imports /root/reference through oracle/make_golden.py).

tests/test_gpu_edges.py compares the CUDA path with the oracle at the shortest frame each backbone accepts and around one 32-step
block; this script is what makes the oracle trustworthy there: reference fp64 == oracle fp64 (out, loss, dL/dx, dL/dparams) for
every backbone at those lengths, and the lengths the reference itself rejects are recorded.  Prints one JSON line.
Run by tests/test_oracle_edges.py when /root/reference is present."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import torch
import make_golden as mg          # imports the reference
from oracle import oracle

# kind, H, shortest frame the reference accepts
CELLS = [("gru", 9, 1), ("dgru", 13, 1), ("dgru", 10, 1), ("qgru", 11, 1), ("qgru_amp1", 10, 1), ("lstm", 9, 1), ("deltagru", 15, 1), ("deltagru_tcnskip", 15, 1),
         ("pgjanet", 13, 1), ("dvrjanet", 11, 1), ("gmp", 0, 1), ("tcnn", 7, 1), ("neuraltx", 9, 1), ("deltajanet", 11, 1),
         ("rvtdcnn", 7, 3), ("mcldnn", 7, 4), ("bojanet", 9, 15), ("apnrru", 7, 15)]
LENGTHS = (1, 2, 3, 4, 5, 15, 16, 31, 32, 33, 65)


def main():
    worst, rejected, n = {}, {}, 0
    rng = np.random.default_rng(12)
    for kind, H, tmin in CELLS:
        thx, thh = (0.01, 0.05) if kind in ("deltagru", "deltagru_tcnskip") else (0.0, 0.0)
        net = mg.build(kind, max(H, 1), 7, thx, thh)
        if kind == "apnrru":
            with torch.no_grad():
                net.backbone.rru.Z.normal_(0.0, 0.5)
        params = mg.flat_params(net)
        rej = []
        for T in LENGTHS:
            x = np.clip(0.2 * rng.standard_normal((3, T, 2)), -0.7, 0.7)
            y = x * (1 - 0.2 * (x ** 2).sum(-1, keepdims=True))
            try:
                ref = mg.run(net, x, y, torch.float64)
                assert ref["out"].shape == (3, T, 2)
            except Exception:
                rej.append(T)
                continue
            if T < tmin:                      # accepted by accident of a reshape (BOJANET at T=10 with B=2): not a supported length
                continue
            r = oracle.run(kind, x, params, target=y, H=H, thx=thx, thh=thh, dtype=np.float64, nthreads=1)
            for key in ("out", "gx", "gparams"):
                e = float(np.abs(r[key] - ref[key]).max() / (np.abs(ref[key]).max() + 1e-300))
                worst[kind] = max(worst.get(kind, 0.0), e)
            worst[kind] = max(worst[kind], abs(float(r["loss"]) - float(ref["loss"])) / abs(float(ref["loss"])))
            n += 1
        rejected[kind] = rej
    print(json.dumps({"cases": n, "worst_rel_err": worst, "reference_rejects": rejected}))


if __name__ == "__main__":
    main()
